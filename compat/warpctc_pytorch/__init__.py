from vistaocr_b200.warpctc import CTCLoss  # noqa: F401
