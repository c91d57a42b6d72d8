"""Oracle (test infrastructure): the keep-mask generator of the inter-layer LSTM dropout, restated in numpy.

The reference gets its dropout mask from cuDNN's RNN dropout state (nn.LSTM(dropout=p), src/models/cnnlstm.py:148-149;
p = 0.5 at src/train_cnn_lstm.py:331) - a stream no other implementation can reproduce bit for bit, and that the
reference's own results do not depend on beyond "i.i.d. Bernoulli(1-p) keep, kept values scaled by 1/(1-p)".  The
product therefore (a) takes an injected mask, so both sides of a parity test use the SAME mask, and (b) otherwise draws
it from Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11 - the generator behind
cuRAND / torch CUDA dropout).  This file restates (b) from the published algorithm; tests pin it to the paper's
known-answer vectors and the kernel (csrc/dropout.cu) to it.

Element i uses word (i & 3) of Philox4x32-10(counter = {lo32(i>>2), hi32(i>>2), lo32(offset), hi32(offset)},
key = {lo32(seed), hi32(seed)}) and is kept iff word >= floor(p * 2^32).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: 4 arrays of uint32 (same shape), key: 2 uint32 scalars/arrays -> 4 arrays of uint32."""
    c0, c1, c2, c3 = (np.asarray(c, np.uint32) for c in ctr)
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def keep_mask(n, p, seed, offset):
    """uint8[n]: 1 = element kept."""
    groups = (n + 3) // 4
    g = np.arange(groups, dtype=np.uint64)
    seed, offset = int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1)
    ctr = ((g & MASK32).astype(np.uint32), (g >> np.uint64(32)).astype(np.uint32),
           np.full(groups, offset & 0xFFFFFFFF, np.uint32), np.full(groups, offset >> 32, np.uint32))
    words = np.stack(philox4x32_10(ctr, (seed & 0xFFFFFFFF, seed >> 32)), axis=1).reshape(-1)[:n]
    t = float(p) * 4294967296.0
    thresh = np.uint32(0xFFFFFFFF if t >= 4294967295.0 else int(t))
    return (words >= thresh).astype(np.uint8)


def dropout_ref(x, p, mask):
    """x * keep / (1-p) with the scale formed in float32 like the kernel (and torch)."""
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return np.where(mask.reshape(x.shape) != 0, x * scale, 0).astype(x.dtype)
