"""GPU parity (exact) of the batched edit distance / CER / WER against the oracle restatement of textutils."""
import numpy as np
import pytest

from oracle.scoring_ref import compute_cer_wer_ref, edit_distance_ref

pytestmark = pytest.mark.gpu


def test_edit_distance_exact(cuda):
    from vistaocr_b200.scoring import edit_distances
    rng = np.random.default_rng(0)
    a, b = [], []
    for n, m in [(0, 0), (0, 5), (4, 0), (1, 1), (7, 3), (50, 60), (200, 180), (333, 1), (2, 400), (129, 128)]:
        x = rng.integers(0, 6, size=n).tolist()
        y = (x[: m // 2] + rng.integers(0, 6, size=max(0, m - m // 2)).tolist())[:m] if m else []
        while len(y) < m:
            y.append(int(rng.integers(0, 6)))
        a.append(x)
        b.append(y)
    got = edit_distances(a, b).tolist()
    assert got == [edit_distance_ref(x, y) for x, y in zip(a, b)]


def test_cer_wer_batch(cuda):
    from vistaocr_b200.scoring import compute_cer_wer_batch, form_tokenized_words
    rng = np.random.default_rng(1)
    alphabet = ["u0020", "u002e", "u0031", "u002c"] + ["u%04x" % (0x61 + i) for i in range(20)]

    def line(n):
        return " ".join(alphabet[int(k)] for k in rng.integers(0, len(alphabet), size=n))

    refs = [line(int(rng.integers(5, 80))) for _ in range(40)]
    hyps = []
    for r in refs:
        toks = r.split(" ")
        for _ in range(int(rng.integers(0, 8))):  # random edits
            k = int(rng.integers(0, len(toks)))
            op = rng.integers(0, 3)
            if op == 0 and len(toks) > 1:
                del toks[k]
            elif op == 1:
                toks.insert(k, alphabet[int(rng.integers(0, len(alphabet)))])
            else:
                toks[k] = alphabet[int(rng.integers(0, len(alphabet)))]
        hyps.append(" ".join(toks))
    hyps[3] = ""  # an empty hypothesis splits into [''] in the reference
    got = compute_cer_wer_batch(hyps, refs)
    want = [compute_cer_wer_ref(h, r, form_tokenized_words) for h, r in zip(hyps, refs)]
    assert got == want
