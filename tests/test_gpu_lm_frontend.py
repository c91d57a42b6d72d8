"""GPU parity of the LM-decode front end against the oracle restatement of LmDecoder.decode's host half.
Fill values and the sparsity pattern are exact; log-softmax values agree to 2e-6 (fp32 exp/log implementations)."""
import numpy as np
import pytest
import torch

from oracle.lm_frontend_ref import lm_remap_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,B,A,U", [(37, 5, 40, 30), (64, 9, 121, 200), (10, 3, 7, 3)])
def test_lm_frontend_matches_oracle(cuda, T, B, A, U):
    from vistaocr_b200 import Alphabet
    from vistaocr_b200.lm_frontend import FILL, LmFrontend
    rng = np.random.default_rng(T + A)
    chars = ["<ctc-blank>"] + ["u%04x" % (0x61 + i) for i in range(A - 1)]
    alpha = Alphabet(chars)
    lm_units = [chars[i] for i in rng.permutation(np.arange(1, A))[: min(U, A - 1) // 2]] + \
               ["u%04x" % (0x4e00 + i) for i in range(U - min(U, A - 1) // 2)]   # half known, half unknown to the model
    x = torch.from_numpy((rng.normal(size=(T, B, A)) * 3).astype(np.float32))
    lens = rng.integers(0, T + 1, size=B)
    lens[0] = T
    want = lm_remap_ref(x, lens, alpha.idx_to_char, lm_units)
    got = LmFrontend(alpha, lm_units).log_probs_for_lm(x.to(cuda), torch.from_numpy(lens))
    assert len(got) == B
    for b in range(B):
        assert got[b].dtype == np.float64 and got[b].shape == want[b].shape
        fill = want[b] == np.log(1e-10)
        assert np.array_equal(got[b] == FILL, fill)
        assert np.abs(got[b] - want[b]).max(initial=0.0) <= 2e-6
