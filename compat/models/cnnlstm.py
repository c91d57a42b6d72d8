from vistaocr_b200.cnnlstm import CnnOcrModel  # noqa: F401
