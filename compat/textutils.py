from vistaocr_b200.textutils import uxxxx_to_utf8, utf8_to_uxxxx  # noqa: F401
