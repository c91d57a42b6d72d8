"""vistaocr_b200: B200-native (sm_100a) line-recognition hot path of isi-vista/VistaOCR behind the
reference's own Python surface.  See DESIGN.md; the C ABI is include/vistaocr_b200.h."""
from .alphabet import Alphabet  # noqa: F401
from .decoder import ArgmaxDecoder  # noqa: F401
from .warpctc import CTCLoss  # noqa: F401
from .cnnlstm import CnnOcrModel  # noqa: F401
from .optim import ClampAdam, train_step  # noqa: F401
from .ops import get_precision, set_precision  # noqa: F401
from .graphs import GraphedDecoder, GraphedTrainStep  # noqa: F401

__all__ = ["Alphabet", "ArgmaxDecoder", "CTCLoss", "CnnOcrModel", "ClampAdam", "train_step", "set_precision",
           "get_precision", "GraphedTrainStep", "GraphedDecoder"]
