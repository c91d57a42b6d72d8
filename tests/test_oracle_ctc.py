"""CPU: pin the CTC oracle (oracle/ctc_ref.c) against the warp-ctc known-answer vector, brute force and torch."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.ctc_ref import ctc_brute_force, ctc_ref


def test_warpctc_known_answer():
    # warp-ctc tests/test_cpu.cpp small_test / pytorch_binding tests: T=2, A=5, labels {1,2}
    acts = np.array([[[0.1, 0.6, 0.1, 0.1, 0.1]], [[0.1, 0.1, 0.6, 0.1, 0.1]]], np.float32)
    costs, grads = ctc_ref(acts, [1, 2], [2], [2])
    assert abs(costs[0] - 2.4628584384918) < 1e-6
    want = np.array([[0.177031, -0.708125, 0.177031, 0.177031, 0.177031],
                     [0.177031, 0.177031, -0.708125, 0.177031, 0.177031]], np.float32)
    np.testing.assert_allclose(grads[:, 0], want, atol=1e-6)


@pytest.mark.parametrize("T,A,label", [(1, 3, [1]), (3, 3, [1]), (4, 3, [1, 1]), (5, 4, [2, 2, 1]), (6, 3, []),
                                        (5, 3, [1, 2, 1]), (2, 3, [1, 1])])
def test_brute_force(T, A, label):
    rng = np.random.default_rng(T * 100 + A)
    acts = rng.normal(size=(T, 1, A)).astype(np.float32)
    costs, _ = ctc_ref(acts, label, [T], [len(label)])
    bf = ctc_brute_force(acts[:, 0], label)
    if np.isinf(bf):  # infeasible: adopt warp-ctc's cost 0 / grad 0
        assert costs[0] == 0.0
    else:
        assert abs(costs[0] - bf) < 1e-9 * max(1.0, abs(bf))


def test_against_torch_ctc():
    rng = np.random.default_rng(3)
    T, B, A = 60, 9, 17
    acts = rng.normal(size=(T, B, A)).astype(np.float32) * 2
    ll = np.array([5, 0, 12, 3, 29, 1, 7, 40, 2], np.int32)
    al = np.array([60, 40, 30, 20, 59, 1, 14, 39, 3], np.int32)  # utterance 7 is infeasible (40 labels, 39 frames)
    labels = rng.integers(1, A, size=int(ll.sum())).astype(np.int32)
    labels[1] = labels[0]
    costs, grads = ctc_ref(acts, labels, al, ll)
    x = torch.tensor(acts, dtype=torch.float64, requires_grad=True)
    loss = F.ctc_loss(x.log_softmax(2), torch.tensor(labels, dtype=torch.long), torch.tensor(al, dtype=torch.long),
                      torch.tensor(ll, dtype=torch.long), blank=0, reduction="none", zero_infinity=True)
    loss.sum().backward()
    np.testing.assert_allclose(costs, loss.detach().numpy(), rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(grads, x.grad.numpy(), atol=2e-7)
    assert costs[7] == 0.0 and not grads[:, 7].any()
    for b in range(B):
        assert not grads[al[b]:, b].any()
