"""CPU: the import-redirect boundary (`compat/`) and the checkpoint contract (reference src/models/cnnlstm.py:40-71,
src/train_cnn_lstm.py:427-438, src/utils/decode.py:58-71).  No kernel runs here: construction, state_dict and
snapshot loading are host logic."""
import importlib
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"
HP = dict(input_line_height=60, rds_line_height=30, lstm_input_dim=8, num_lstm_layers=2, num_lstm_hidden_units=12,
          p_lstm_dropout=0.5)


def test_compat_import_lines_resolve_to_this_package():
    """The reference's own import lines (train_cnn_lstm.py:12,25, decode_testset.py:6,14), with compat/ first on the
    path, in a fresh interpreter so that no test-session module shadows them."""
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from models.cnnlstm import CnnOcrModel\n"
        "from warpctc_pytorch import CTCLoss\n"
        "from decoder import ArgmaxDecoder\n"
        "from alphabet import Alphabet\n"
        "from textutils import uxxxx_to_utf8\n"
        "import vistaocr_b200 as v\n"
        "assert CnnOcrModel is v.CnnOcrModel and CTCLoss is v.CTCLoss and ArgmaxDecoder is v.ArgmaxDecoder\n"
        "assert Alphabet is v.Alphabet and uxxxx_to_utf8('u0041 u0020 u00e9') == 'A \\xe9'\n"
        "a = Alphabet(['<ctc-blank>', 'u0041']); assert len(a) == 2 and a.idx_to_char[1] == 'u0041'\n"
        "import torch.nn as nn; assert isinstance(CTCLoss(), nn.Module)\n"
        "print('ok')\n" % (ROOT, os.path.join(ROOT, "compat")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr


def _snapshot(model, path, module_prefix=False):
    """What train_cnn_lstm.py:427-438 writes."""
    sd = model.state_dict()
    if module_prefix:  # a model trained under nn.DataParallel(cnn): cnn.N.* -> cnn.module.N.*
        sd = {(("cnn.module." + k[4:]) if k.startswith("cnn.") else k): v for k, v in sd.items()}
    torch.save({"iteration": 12, "state_dict": sd, "optimizer": {}, "model_hyper_params": model.get_hyper_params(),
                "rtl": False, "cur_lr": 1e-3, "val_loss": 1.0, "val_cer": 0.5, "val_wer": 0.75, "line_height": 60}, path)


@pytest.mark.parametrize("module_prefix", [False, True])
def test_from_saved_weights_round_trip(tmp_path, module_prefix):
    from vistaocr_b200 import Alphabet, CnnOcrModel
    alpha = Alphabet(["<ctc-blank>"] + ["u%04x" % (0x61 + i) for i in range(10)])
    torch.manual_seed(3)
    m = CnnOcrModel(alphabet=alpha, gpu=False, multigpu=True, verbose=False, **HP)
    assert m.multigpu is True and m.get_hyper_params()["multigpu"] is False  # never DataParallel: keys are cnn.N.*
    assert not any(".module." in k for k in m.state_dict())
    path = str(tmp_path / "snap.pth")
    _snapshot(m, path, module_prefix)
    m2 = CnnOcrModel.FromSavedWeights(path, verbose=False, gpu=False)
    assert m2.rtl is False and m2.input_line_height == 60 and len(m2.alphabet) == 11
    a, b = m.state_dict(), m2.state_dict()
    assert list(a) == list(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    # strict loading: a missing / unexpected key must raise like the reference's load_state_dict(strict=True)
    bad = {k: v for k, v in a.items() if k != "prob_layer.0.bias"}
    with pytest.raises(RuntimeError):
        m2.load_state_dict(bad, strict=True)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_snapshots_cross_load_with_the_reference(tmp_path):
    """Both directions against the UNMODIFIED reference class: it loads a snapshot written from our model
    (FromSavedWeights, strict=True), and we load one written from it - with and without the DataParallel prefix."""
    import types
    from oracle.decode_ref import uxxxx_to_utf8
    from vistaocr_b200 import CnnOcrModel
    saved = {k: sys.modules.get(k) for k in ("textutils", "alphabet", "decoder", "models", "models.cnnlstm")}
    stub = types.ModuleType("textutils")
    stub.uxxxx_to_utf8 = uxxxx_to_utf8
    sys.modules["textutils"] = stub
    for k in ("alphabet", "decoder", "models", "models.cnnlstm"):
        sys.modules.pop(k, None)
    sys.path.insert(0, REF)
    try:
        ref_alphabet = importlib.import_module("alphabet")
        ref_cnnlstm = importlib.import_module("models.cnnlstm")
        alpha = ref_alphabet.Alphabet(["<ctc-blank>"] + ["u%04x" % (0x61 + i) for i in range(10)])
        torch.serialization.add_safe_globals([ref_alphabet.Alphabet])
        torch.manual_seed(5)
        ours = CnnOcrModel(alphabet=alpha, gpu=False, multigpu=True, verbose=False, **HP)
        p1 = str(tmp_path / "ours.pth")
        _snapshot(ours, p1)
        theirs = ref_cnnlstm.CnnOcrModel.FromSavedWeights(p1, verbose=False, gpu=False)
        for k, v in ours.state_dict().items():
            assert torch.equal(theirs.state_dict()[k], v), k
        for prefix in (False, True):
            p2 = str(tmp_path / ("theirs%d.pth" % prefix))
            _snapshot(theirs, p2, module_prefix=prefix)
            back = CnnOcrModel.FromSavedWeights(p2, verbose=False, gpu=False)
            for k, v in theirs.state_dict().items():
                assert torch.equal(back.state_dict()[k], v), k
    finally:
        sys.path.remove(REF)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_ctc_infeasible_count_is_host_arithmetic():
    from vistaocr_b200.warpctc import count_infeasible
    lab = torch.tensor([1, 1, 2, 3, 4, 4, 4, 5, 6], dtype=torch.int32)
    ll = torch.tensor([3, 1, 3, 0, 2], dtype=torch.int32)  # needs 4, 1, 5, 0, 2 frames
    assert count_infeasible(lab, torch.tensor([4, 1, 5, 0, 2], dtype=torch.int32), ll) == 0
    assert count_infeasible(lab, torch.tensor([3, 1, 4, 0, 1], dtype=torch.int32), ll) == 3
    assert count_infeasible(lab[:0], torch.tensor([0, 2], dtype=torch.int32), torch.tensor([0, 0], dtype=torch.int32)) == 0
