"""Oracle (test infrastructure): edit distance + CER/WER restated in plain Python/NumPy
(reference src/textutils.py:264-351).  Pinned against the reference functions themselves (extracted from the source
file because the module cannot be imported: ICU + absolute data paths) in tests/test_oracle_vs_reference.py."""
import numpy as np


def edit_distance_ref(A, B):
    if len(A) == 0 and len(B) == 0:
        return 0
    if len(A) == 0 or len(B) == 0:
        return len(A) + len(B)
    prev = np.arange(len(B) + 1)
    for i in range(1, len(A) + 1):
        cur = np.empty_like(prev)
        cur[0] = i
        for j in range(1, len(B) + 1):
            cur[j] = prev[j - 1] if A[i - 1] == B[j - 1] else 1 + min(cur[j - 1], prev[j], prev[j - 1])
        prev = cur
    return int(prev[-1])


def compute_cer_wer_ref(hyp, ref, form_tokenized_words):
    hyp_chars, ref_chars = hyp.split(" "), ref.split(" ")
    char_dist = edit_distance_ref(hyp_chars, ref_chars)
    strip = lambda w: _strip(w)
    hyp_words, ref_words = strip(form_tokenized_words(hyp_chars)), strip(form_tokenized_words(ref_chars))
    word_dist = edit_distance_ref(hyp_words, ref_words)
    return float(char_dist) / len(ref_chars), float(word_dist) / len(ref_words)


def _strip(words):
    while len(words) > 0 and words[0] == "u0020":
        words = words[1:]
    while len(words) > 0 and words[-1] == "u0020":
        words = words[:-1]
    return words
