"""Alphabet container with the interface the decoder boundary needs (reference src/alphabet.py:1-12):
`len(alphabet)`, `alphabet.idx_to_char[i]`, `alphabet.char_to_idx[c]`.  Index 0 is the CTC blank
(reference src/ocr_dataset.py:109-120)."""


class Alphabet(object):
    def __init__(self, char_array, left_to_right=False):
        chars = list(char_array)
        self.left_to_right = left_to_right
        self.char_array = char_array
        self.char_to_idx = {c: i for i, c in enumerate(chars)}
        self.idx_to_char = {i: c for i, c in enumerate(chars)}

    def __len__(self):
        return len(self.idx_to_char)
