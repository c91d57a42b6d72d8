"""Body of __graft_entry__.smoke(): the oracle is imported here only as the checker."""


def run(dev, np):
    import torch
    from oracle.ctc_ref import ctc_ref
    from oracle.decode_ref import decode_loop
    from . import Alphabet, ArgmaxDecoder
    from .warpctc import ctc_costs_and_grads

    rng = np.random.default_rng(0)
    T, B, A = 48, 8, 40
    x = rng.normal(size=(T, B, A)).astype(np.float32)
    lens = np.array([48, 48, 40, 33, 20, 9, 1, 0], np.int32)
    alpha = Alphabet(["<ctc-blank>"] + ["u%04x" % (0x61 + i) for i in range(A - 1)])
    got = ArgmaxDecoder(alpha).decode(torch.from_numpy(x).to(dev), torch.from_numpy(lens), uxxxx=True)
    want = decode_loop(x, lens, alpha.idx_to_char, uxxxx=True)
    assert got == want, "greedy decode differs from the oracle"
    label_lens = np.array([10, 3, 0, 7, 5, 2, 1, 0], np.int32)
    labels = rng.integers(1, A, size=int(label_lens.sum())).astype(np.int32)
    costs, grads = ctc_costs_and_grads(torch.from_numpy(x).to(dev), torch.from_numpy(labels),
                                       torch.from_numpy(lens), torch.from_numpy(label_lens))
    wc, wg = ctc_ref(x, labels, lens, label_lens)
    assert np.allclose(costs.cpu().numpy(), wc, rtol=1e-5, atol=1e-5), "CTC cost differs from the oracle"
    assert np.abs(grads.cpu().numpy() - wg).max() <= 5e-5, "CTC gradient differs from the oracle"
    print("smoke ok: decode + ctc match the oracle")
