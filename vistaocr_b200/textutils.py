"""The one text helper the hot path needs: `uxxxx_to_utf8` (reference src/textutils.py:216-243).
The reference module cannot be imported outside its authors' cluster (ICU bindings, absolute data paths)."""


def uxxxx_to_utf8(in_str):
    """'u0041 u0062' -> 'Ab'.  The tokens <unk>, <s>, </s> pass through; an empty/blank string maps to ''."""
    if in_str.strip() == "":
        return ""
    out = []
    for tok in in_str.split():
        if tok in ("<unk>", "<s>", "</s>"):
            out.append(tok)
        else:
            out.append(chr(int(tok[1:], 16)))
    return "".join(out)


def utf8_to_uxxxx(in_str, output_array=False):
    """Inverse mapping (reference src/textutils.py:245-255)."""
    arr = ["u%s" % hex(ord(ch))[2:].zfill(4).lower() for ch in in_str]
    return arr if output_array else " ".join(arr)
