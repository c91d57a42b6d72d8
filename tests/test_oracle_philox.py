"""CPU: the dropout mask generator's restatement (oracle/philox_ref.py) against the published Philox4x32-10
known-answer vectors (Random123 kat_vectors: zero, all-ones and pi-digit counters / keys) and its own mask contract."""
import numpy as np

from oracle.philox_ref import dropout_ref, keep_mask, philox4x32_10

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers():
    for ctr, key, want in KAT:
        got = philox4x32_10([np.array([c], np.uint32) for c in ctr], key)
        assert tuple(int(g[0]) for g in got) == want


def test_keep_mask_contract():
    m = keep_mask(1_000_003, 0.5, seed=1234, offset=77)
    assert m.dtype == np.uint8 and m.shape == (1_000_003,) and abs(m.mean() - 0.5) < 2e-3
    assert np.array_equal(m[:1000], keep_mask(1000, 0.5, 1234, 77))  # a prefix of the same stream
    assert not np.array_equal(m[:1000], keep_mask(1000, 0.5, 1234, 78))  # another offset, another stream
    assert not np.array_equal(m[:1000], keep_mask(1000, 0.5, 1235, 77))
    assert keep_mask(4096, 0.0, 1, 1).all() and abs(keep_mask(400_000, 0.25, 9, 0).mean() - 0.75) < 3e-3
    # offsets beyond 32 bits reach the upper counter word
    assert not np.array_equal(keep_mask(64, 0.5, 3, 1), keep_mask(64, 0.5, 3, 1 + (1 << 32)))
    x = np.arange(-4, 4, dtype=np.float32)
    y = dropout_ref(x, 0.5, np.array([1, 0, 1, 1, 0, 0, 1, 1], np.uint8))
    assert y.tolist() == [-8.0, 0.0, -4.0, -2.0, 0.0, 0.0, 4.0, 6.0]
