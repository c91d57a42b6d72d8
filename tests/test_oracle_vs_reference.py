"""CPU, authoring container only: the oracle restatements against the UNMODIFIED reference imported read-only from
/root/reference/src (skipped on the GPU box, where the reference does not exist; the committed fixtures in
tests/golden/ carry the same comparison there)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    from oracle.decode_ref import uxxxx_to_utf8
    stub = types.ModuleType("textutils")  # the real one needs ICU + absolute data paths (textutils.py:3,9-35)
    stub.uxxxx_to_utf8 = uxxxx_to_utf8
    saved = {k: sys.modules.get(k) for k in ("textutils", "alphabet", "decoder", "models", "models.cnnlstm")}
    sys.modules["textutils"] = stub
    for k in ("alphabet", "decoder", "models", "models.cnnlstm"):
        sys.modules.pop(k, None)
    sys.path.insert(0, REF)
    import alphabet
    import decoder
    from models import cnnlstm
    yield types.SimpleNamespace(Alphabet=alphabet.Alphabet, ArgmaxDecoder=decoder.ArgmaxDecoder,
                                CnnOcrModel=cnnlstm.CnnOcrModel)
    sys.path.remove(REF)
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def test_decode_oracle_equals_reference_decoder(ref):
    from oracle.decode_ref import decode_loop
    from tests.test_oracle_decode import _adversarial_logits
    for A in (3, 80, 97, 121, 200):
        rng = np.random.default_rng(A)
        x = _adversarial_logits(rng, 29, 4, A)
        lens = np.array([29, 20, 3, 0], np.int32)
        alpha = ref.Alphabet(["<ctc-blank>"] + ["u%04x" % (0x61 + i) for i in range(A - 1)])
        want = ref.ArgmaxDecoder(alpha).decode(torch.from_numpy(x), torch.from_numpy(lens), uxxxx=True)
        assert decode_loop(x, lens, alpha.idx_to_char, uxxxx=True) == want


def test_model_oracle_equals_reference_model(ref):
    from oracle import model_ref as M
    hp = dict(input_line_height=60, rds_line_height=30, lstm_input_dim=16, num_lstm_layers=2,
              num_lstm_hidden_units=24, p_lstm_dropout=0.0)
    A = 13
    alpha = ref.Alphabet(["<ctc-blank>"] + ["u%04x" % (0x61 + i) for i in range(A - 1)])
    model = ref.CnnOcrModel(alphabet=alpha, gpu=False, multigpu=False, verbose=False, **hp)
    sd = M.make_state_dict(hp, A, seed=1)
    assert set(model.state_dict().keys()) == set(sd.keys())
    model.load_state_dict(sd, strict=True)
    rng = np.random.default_rng(0)
    x, widths, _, _ = M.synth_batch(rng, 3, 60, 60, 140, A, n_rds=1)
    u1 = torch.from_numpy(rng.random((3, 64, 2)).astype(np.float32))
    u2 = torch.from_numpy(rng.random((3, 128, 2)).astype(np.float32))
    model.cnn[6]._random_samples, model.cnn[13]._random_samples = u1, u2
    for training in (False, True):
        model.train(training)
        with torch.no_grad():
            want, lens = model(torch.from_numpy(x), torch.from_numpy(widths))
        got, glens = M.forward_ref(sd, torch.from_numpy(x), widths, hp, (u1, u2), training=training,
                                   bn_updates={} if training else None)
        assert glens.tolist() == lens.tolist()
        assert (got - want).abs().max().item() <= 1e-6
        got2, _ = M.forward_ref(sd, torch.from_numpy(x), widths, hp, (u1, u2), training=training,
                                bn_updates={} if training else None, use_nn_lstm=False)
        assert (got2 - want).abs().max().item() <= 2e-6


def test_our_model_class_mirrors_reference_state_dict_and_init(ref):
    from vistaocr_b200 import CnnOcrModel
    for hp in (dict(input_line_height=30, rds_line_height=30), dict(input_line_height=120, rds_line_height=30)):
        hp.update(lstm_input_dim=8, num_lstm_layers=2, num_lstm_hidden_units=12, p_lstm_dropout=0.5)
        alpha = ref.Alphabet(list(range(17)))
        torch.manual_seed(7)
        r = ref.CnnOcrModel(alphabet=alpha, gpu=False, multigpu=False, verbose=False, **hp)
        torch.manual_seed(7)
        m = CnnOcrModel(alphabet=alpha, gpu=False, multigpu=False, verbose=False, **hp)
        a, b = r.state_dict(), m.state_dict()
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k
        assert [n for n, _ in r.named_parameters()] == [n for n, _ in m.named_parameters()]
        for w in (15, 200, 350, 801):
            assert m.cnn_input_size_to_output_size((hp["input_line_height"], w)) == \
                r.cnn_input_size_to_output_size((hp["input_line_height"], w))
        with pytest.raises(Exception):
            CnnOcrModel(30)
        with pytest.raises(Exception):
            CnnOcrModel(alphabet=alpha, **dict(hp, input_line_height=45, rds_line_height=30))


def test_collate_oracle_equals_reference_collater(ref):
    import importlib
    import numpy as np
    from oracle.collate_ref import collate_ref
    sys.path.insert(0, REF)
    try:
        datautils = importlib.import_module("datautils")
    finally:
        sys.path.remove(REF)
    rng = np.random.default_rng(3)
    batch = []
    for i in range(23):
        w = int(rng.integers(15, 200)) if i % 4 else 77  # repeated keys exercise the stable sort
        batch.append((rng.random((1, 30, w), dtype=np.float32), rng.integers(1, 50, size=int(rng.integers(0, 9))).tolist(),
                      {"width": w, "utt-id": "u%d" % i}))
    want = datautils.SortByWidthCollater([(torch.from_numpy(i), t, dict(m)) for i, t, m in batch])
    got = collate_ref(batch)
    assert torch.equal(want[0], torch.from_numpy(got[0]))
    assert want[1].tolist() == got[1].tolist() and want[2].tolist() == got[2].tolist()
    assert want[3].tolist() == got[3].tolist()
    assert want[4]["utt-ids"] == [batch[i][2]["utt-id"] for i in got[4]]
    sys.modules.pop("datautils", None)


def test_lm_frontend_oracle_equals_reference_lmdecoder(ref):
    """LmDecoder.__init__ needs the external eesen binding; build the object without it, fill in the attributes its
    decode() reads (decoder.py:36-59) and capture what it submits to the executor."""
    import importlib
    import numpy as np
    from oracle.lm_frontend_ref import lm_remap_ref
    sys.path.insert(0, REF)
    try:
        decoder = importlib.import_module("decoder")
    finally:
        sys.path.remove(REF)
    A = 17
    chars = ["<ctc-blank>"] + ["u%04x" % (0x61 + i) for i in range(A - 1)]
    alpha = ref.Alphabet(chars)
    lm_units = [chars[3], chars[9], "u4e00", chars[1], "u4e01", chars[16]]
    units = ["<ctc-blank>"] + lm_units
    d = decoder.LmDecoder.__new__(decoder.LmDecoder)
    d.alphabet = alpha
    d.lmidx_to_char = units
    d.lmchar_to_idx = dict(zip(units, range(len(units))))
    d.lm_swap_idxs_modelidx = [m for m in range(A) if chars[m] in d.lmchar_to_idx]
    d.lm_swap_idxs_lmidx = [d.lmchar_to_idx[chars[m]] for m in d.lm_swap_idxs_modelidx]

    class Capture:
        def __init__(self):
            self.got = []

        def submit(self, fn, probs, uttid, uxxxx):
            self.got.append(probs)

    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.normal(size=(11, 3, A)).astype(np.float32))
    lens = [11, 7, 0]
    cap = Capture()
    d.decode(cap, x, lens, ["a", "b", "c"])
    want = lm_remap_ref(x, lens, alpha.idx_to_char, lm_units)
    for g, w in zip(cap.got, want):
        assert g.dtype == np.float64 and np.array_equal(g, w)


def test_scoring_oracle_equals_reference_textutils():
    """textutils.py cannot be imported (ICU, absolute paths): lift the three pure functions out of its source."""
    import ast
    import numpy as np
    from oracle.scoring_ref import compute_cer_wer_ref, edit_distance_ref
    from vistaocr_b200.scoring import form_tokenized_words as ours_ftw
    src = open(os.path.join(REF, "textutils.py")).read()
    tree = ast.parse(src)
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("edit_distance", "form_tokenized_words", "compute_cer_wer"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), "textutils.py", "exec"), ns)
    rng = np.random.default_rng(5)
    alphabet = ["u0020", "u002e", "u0031", "u002c", "u0661"] + ["u%04x" % (0x61 + i) for i in range(12)]
    for _ in range(30):
        ref = " ".join(alphabet[int(k)] for k in rng.integers(0, len(alphabet), size=int(rng.integers(3, 40))))
        hyp = " ".join(alphabet[int(k)] for k in rng.integers(0, len(alphabet), size=int(rng.integers(0, 40))))
        assert ns["edit_distance"](hyp.split(" "), ref.split(" ")) == edit_distance_ref(hyp.split(" "), ref.split(" "))
        assert ns["form_tokenized_words"](ref.split(" ")) == ours_ftw(ref.split(" "))
        assert ns["form_tokenized_words"](ref.split(" "), with_spaces=True) == ours_ftw(ref.split(" "), with_spaces=True)
        try:
            want = ns["compute_cer_wer"](hyp, ref)
        except ZeroDivisionError:
            continue
        assert want == compute_cer_wer_ref(hyp, ref, ours_ftw)
