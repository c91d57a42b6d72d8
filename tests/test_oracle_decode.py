"""CPU: the vectorised decode oracle agrees with its literal frame loop, incl. ties / thresholds / NaN."""
import numpy as np
import pytest

from oracle.decode_ref import collapse, decode_labels, decode_loop


def _adversarial_logits(rng, T, B, A):
    x = rng.normal(size=(T, B, A)).astype(np.float32)
    thr = np.float32(3 / A)
    # quantise so ties are frequent, plant exact-threshold and just-below-threshold maxima, blanks, repeats
    x = np.round(x * 2) / 2
    x[rng.random((T, B)) < 0.2, 0] = 9.0  # blank wins
    m = rng.random((T, B)) < 0.15
    x[m] = np.minimum(x[m], thr)  # max == float32(thresh): not below
    m = rng.random((T, B)) < 0.1
    x[m] = np.minimum(x[m], np.nextafter(thr, np.float32(-1)))
    return x.astype(np.float32)


@pytest.mark.parametrize("A", [2, 5, 80, 97, 120, 121, 167, 200])
def test_vectorised_matches_loop(A):
    rng = np.random.default_rng(A)
    T, B = 37, 6
    x = _adversarial_logits(rng, T, B, A)
    lens = np.array([37, 36, 20, 1, 0, 40], np.int32)
    idx_to_char = {i: "u%04x" % (0x40 + i) for i in range(A)}
    want = decode_loop(x, lens, idx_to_char, uxxxx=True)
    labs, path = decode_labels(x, lens, A)
    got = [" ".join(idx_to_char[k] for k in l) for l in labs]
    assert got == want
    assert path.shape == (B, T) and (path[4] == -1).all()


def test_nan_is_maximal_like_numpy():
    x = np.zeros((3, 1, 4), np.float32)
    x[1, 0, 2] = np.nan
    x[0, 0, 1] = 5
    x[2, 0, 1] = 5
    idx_to_char = {i: "u%04x" % (0x61 + i) for i in range(4)}
    want = decode_loop(x, [3], idx_to_char)
    labs, _ = decode_labels(x, [3], 4)
    assert [" ".join(idx_to_char[k] for k in labs[0])] == want == ["u0062 u0063 u0062"]


def test_canon_collapses_duplicate_strings():
    assert collapse(np.array([1, 2, 0, 2, 3, -1]), canon=np.array([0, 1, 1, 3])) == [1, 2, 3]
