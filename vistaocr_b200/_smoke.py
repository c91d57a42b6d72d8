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
    # one tiny training step of the whole path (CNN -> BiLSTM -> CTC -> backward -> clamp+Adam) against the oracle
    from oracle import model_ref as M
    from . import ClampAdam, CnnOcrModel, CTCLoss, train_step
    hp = dict(input_line_height=60, rds_line_height=30, lstm_input_dim=32, num_lstm_layers=2,
              num_lstm_hidden_units=32, p_lstm_dropout=0.0)
    A2 = 17
    sd = M.make_state_dict(hp, A2, seed=3)
    model = CnnOcrModel(alphabet=Alphabet(["<ctc-blank>"] + ["u%04x" % (0x61 + i) for i in range(A2 - 1)]),
                        verbose=False, **hp)
    model.load_state_dict(sd, strict=True)
    model.train()
    xb, widths, lab, lab_lens = M.synth_batch(rng, 3, 60, 70, 160, A2, 2, 6, n_rds=1)
    u1 = torch.from_numpy(rng.random((3, 64, 2)).astype(np.float32))
    u2 = torch.from_numpy(rng.random((3, 128, 2)).astype(np.float32))
    model.cnn[6]._random_samples, model.cnn[13]._random_samples = u1, u2
    opt = ClampAdam(model.parameters(), lr=1e-3)
    batch = (torch.from_numpy(xb), torch.from_numpy(lab), torch.from_numpy(widths), torch.from_numpy(lab_lens), {})
    loss = train_step(batch, model, CTCLoss(host_cost=False), opt)
    want, wlens = M.forward_ref(sd, torch.from_numpy(xb), widths, hp, (u1, u2), training=True, bn_updates={})
    wloss = M.ctc_sum_ref(want, lab, wlens, lab_lens).item()
    assert abs(loss[0].item() - wloss) <= 1e-4 * abs(wloss), ("training loss differs from the oracle", loss, wloss)
    model.eval()
    with torch.no_grad():
        logits, lens2 = model(torch.from_numpy(xb).to(dev), torch.from_numpy(widths))
    hyp = model.decode_without_lm(logits, lens2, uxxxx=True)
    assert hyp == decode_loop(logits.cpu().numpy(), lens2.numpy(), model.alphabet.idx_to_char, uxxxx=True)
    # raw uint8 line images -> padded float batch (cv2-exact resize, inversion, /255) against the oracle, bit for bit
    from oracle.preproc_ref import preprocess_line
    from .imagetransforms import LineBatchPreprocessor
    raw = [rng.integers(0, 256, size=(int(h), int(w)), dtype=np.uint8) for h, w in ((47, 301), (60, 400), (33, 20))]
    pb, pw, po = LineBatchPreprocessor(30, invert=True, device=dev)(raw)
    for k, i in enumerate(po.tolist()):
        ref = preprocess_line(raw[i], 30, invert=True, min_width=15)
        assert np.array_equal(pb[k, :, :, :ref.shape[2]].cpu().numpy(), ref), "pre-processing differs from the oracle"
    print("smoke ok: pre-processing + decode + ctc + one training step of the full path match the oracle (loss %.4f)" % wloss)
