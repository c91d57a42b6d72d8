"""GPU parity: csrc/decode.cu through the C ABI vs the decode oracle - bit-exact strings, labels and frame paths."""
import numpy as np
import pytest
import torch

from oracle.decode_ref import decode_labels, decode_loop
from tests.test_oracle_decode import _adversarial_logits

pytestmark = pytest.mark.gpu


def _alphabet(A):
    from vistaocr_b200 import Alphabet
    return Alphabet(["<ctc-blank>"] + ["u%04x" % (0x4e00 + i) for i in range(1, A)])


@pytest.mark.parametrize("T,B,A", [(37, 6, 5), (1, 1, 2), (50, 64, 120), (33, 3, 121), (64, 5, 97), (40, 7, 200),
                                   (3, 130, 80), (129, 2, 167), (10, 4, 1000), (6, 2, 5000), (4, 1, 20000)])
def test_decode_matches_oracle(cuda, T, B, A):
    from vistaocr_b200 import ArgmaxDecoder
    rng = np.random.default_rng(T * 7 + B * 3 + A)
    x = _adversarial_logits(rng, T, B, A)
    lens = rng.integers(0, T + 2, size=B).astype(np.int32)
    lens[0] = T
    alpha = _alphabet(A)
    dec = ArgmaxDecoder(alpha)
    got = dec.decode(torch.from_numpy(x).to(cuda), torch.from_numpy(lens), uxxxx=True)
    want = decode_loop(x, lens, alpha.idx_to_char, uxxxx=True)
    assert got == want
    path = dec.decode_alignment(torch.from_numpy(x).to(cuda), torch.from_numpy(lens)).cpu().numpy()
    _, want_path = decode_labels(x, np.minimum(lens, T), A)
    np.testing.assert_array_equal(path, want_path)
    # utf-8 output and CPU-tensor input (decode_testset.py ships CPU logits to the decoder)
    got8 = dec.decode(torch.from_numpy(x), torch.from_numpy(lens), uxxxx=False)
    assert got8 == decode_loop(x, lens, alpha.idx_to_char, uxxxx=False)


def test_nan_ties_and_empty(cuda):
    from vistaocr_b200 import ArgmaxDecoder
    A = 7
    alpha = _alphabet(A)
    x = np.zeros((5, 3, A), np.float32)  # all ties -> argmax 0 -> blank everywhere
    x[1, 0, 3] = np.nan
    x[2, 0, 3] = np.nan
    x[3, 1, 2] = 0.5
    x[3, 1, 4] = 0.5  # tie: first index (2) wins
    lens = np.array([5, 5, 0], np.int32)
    got = ArgmaxDecoder(alpha).decode(torch.from_numpy(x).to(cuda), torch.from_numpy(lens), uxxxx=True)
    assert got == decode_loop(x, lens, alpha.idx_to_char, uxxxx=True)
    assert got[2] == ""
    # T == 0
    e = ArgmaxDecoder(alpha).decode(torch.zeros((0, 2, A), device=cuda), torch.zeros(2, dtype=torch.int32))
    assert e == ["", ""]


def test_duplicate_symbol_strings_collapse(cuda):
    from vistaocr_b200 import Alphabet, ArgmaxDecoder
    alpha = Alphabet(["<ctc-blank>", "u0061", "u0061", "u0062"])
    x = np.full((4, 1, 4), -5, np.float32)
    for t, k in enumerate([1, 2, 3, 3]):
        x[t, 0, k] = 5
    got = ArgmaxDecoder(alpha).decode(torch.from_numpy(x).to(cuda), torch.tensor([4], dtype=torch.int32), uxxxx=True)
    assert got == decode_loop(x, [4], alpha.idx_to_char, uxxxx=True) == ["u0061 u0062"]


def test_misaligned_view(cuda):
    """A logits view whose base is not 16-B aligned takes the unstaged path; results must not change."""
    from vistaocr_b200.decoder import greedy_decode_labels
    rng = np.random.default_rng(5)
    T, B, A = 20, 3, 121
    x = _adversarial_logits(rng, T, B, A)
    buf = torch.zeros(T * B * A + 1, device=cuda)
    buf[1:] = torch.from_numpy(x).to(cuda).flatten()
    v = buf[1:].view(T, B, A)
    assert v.data_ptr() % 16 != 0
    lens = torch.full((B,), T, dtype=torch.int32)
    labels, counts, path = greedy_decode_labels(v, lens, 3 / A)
    want, want_path = decode_labels(x, lens.numpy(), A)
    np.testing.assert_array_equal(path.cpu().numpy(), want_path)
    for b in range(B):
        assert labels[b, :counts[b]].cpu().tolist() == want[b]


def test_full_size_properties(cuda):
    """BASELINE cfg5-sized batch (512 lines x 392 frames x 120 symbols): size-independent properties -
    idempotence (decoding the one-hot re-encoding of the frame path reproduces it), no blanks / no immediate
    repeats in the output, counts bounded by lens."""
    from vistaocr_b200.decoder import greedy_decode_labels
    g = torch.Generator(device="cuda").manual_seed(7)
    T, B, A = 392, 512, 120
    x = torch.randn((T, B, A), device=cuda, generator=g)
    lens = torch.randint(1, T + 1, (B,), device=cuda, generator=g, dtype=torch.int32)
    labels, counts, path = greedy_decode_labels(x, lens, 3 / A)
    assert (counts <= lens).all() and (counts >= 0).all()
    tt = torch.arange(T, device=cuda)[None, :]
    valid = tt < counts[:, None]
    assert (labels[valid] > 0).all()
    rep = (labels[:, 1:] == labels[:, :-1]) & valid[:, 1:]
    # a repeat in the collapsed output needs a blank between the two frames; check against the path
    am = x.argmax(2).T
    mx = x.max(2).values.T
    want_path = torch.where((am == 0) | (mx < np.float32(3 / A)), 0, am)
    want_path = torch.where(tt < lens[:, None], want_path, -1).int()
    assert torch.equal(path, want_path)
    onehot = torch.nn.functional.one_hot(path.clamp(min=0).long().T, A).float() * 10  # [T,B,A]
    labels2, counts2, path2 = greedy_decode_labels(onehot.contiguous(), lens, 3 / A)
    assert torch.equal(path2, path) and torch.equal(counts2, counts)
    assert torch.equal(torch.where(valid, labels, 0), torch.where(valid, labels2, 0))
    assert rep.sum() >= 0


@pytest.mark.parametrize("T,B,K,A", [(37, 6, 64, 5), (50, 64, 1024, 120), (33, 3, 256, 121), (9, 130, 1024, 128),
                                     (64, 5, 512, 97), (1, 1, 8, 2)])
def test_prob_layer_with_argmax_epilogue(cuda, T, B, K, A):
    """The fused inference tail (ops.linear_argmax + collapse_labels: the prob-layer GEMM reduces every row to its frame
    label in the epilogue and never writes the logits) gives exactly the path / labels / counts of the unfused sequence
    prob layer -> logits -> vocr_greedy_decode_f32, and both match the decode oracle run on those logits."""
    from vistaocr_b200 import ops
    from vistaocr_b200.decoder import collapse_labels, greedy_decode_labels
    rng = np.random.default_rng(T + 3 * B + K + A)
    x = torch.from_numpy(rng.standard_normal((T * B, K)).astype(np.float32)).to(cuda)
    w = torch.from_numpy((rng.standard_normal((A, K)) / K ** 0.5).astype(np.float32)).to(cuda)
    w[min(2, A - 1)] = w[min(1, A - 1)]  # exact ties between two symbols: the lower index must win
    b = torch.from_numpy(rng.standard_normal(A).astype(np.float32) * 0.1).to(cuda)
    b[min(2, A - 1)] = b[min(1, A - 1)]
    lens = rng.integers(0, T + 1, size=B).astype(np.int32)
    lens[0] = T
    lens_dev = torch.from_numpy(lens).to(cuda)
    assert ops.linear_argmax_supported(K, A)
    for thresh in (-1e30, 0.3, 3 * 1 / A):
        with torch.no_grad():
            logits = ops.linear(x, w, b).view(T, B, A)
            labels, counts, path = greedy_decode_labels(logits, lens_dev, thresh)
            fpath = ops.linear_argmax(x, w, b, lens_dev, T, B, thresh)
            flabels, fcounts = collapse_labels(fpath, lens_dev)
        assert torch.equal(fpath, path)
        assert torch.equal(fcounts, counts)
        for i in range(B):
            n = int(counts[i])
            assert torch.equal(flabels[i, :n], labels[i, :n])
        if thresh == 3 * 1 / A:  # the reference's threshold (decoder.py:125): the oracle's frame path
            _, want_path = decode_labels(logits.cpu().numpy(), lens, A)
            np.testing.assert_array_equal(fpath.cpu().numpy(), want_path)
