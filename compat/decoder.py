from vistaocr_b200.decoder import ArgmaxDecoder  # noqa: F401
