from vistaocr_b200.alphabet import Alphabet  # noqa: F401
