"""Oracle (test infrastructure): runs the reference's OWN classes on the CPU - `CnnOcrModel` (src/models/cnnlstm.py) and
`ArgmaxDecoder` (src/decoder.py), unmodified, imported from /root/reference/src in the authoring container or from the
copy staged by oracle/build_oracle.py under oracle/_ref/ on the GPU box - through the reference's call sequences
(train(): src/train_cnn_lstm.py:131-150; decode: src/decode_testset.py:92-101,166).  Used by bench.py's CPU arm and by
the tests as the checker; never by the product.

Accommodations, none of which edits the reference: a stub `textutils` exporting uxxxx_to_utf8 (the real
module needs ICU and absolute data paths), gpu=False / multigpu=False, and warp-ctc (not installable) replaced by
torch.nn.functional.ctc_loss on log_softmax (same cost and gradient, see oracle/ctc_ref.c).
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE = {}


def reference_dir():
    for d in ("/root/reference/src", os.path.join(HERE, "_ref")):
        if os.path.exists(os.path.join(d, "models", "cnnlstm.py")):
            return d
    return None


def load():
    """-> namespace(CnnOcrModel, ArgmaxDecoder, Alphabet, where) of the unmodified reference, or None."""
    if "ns" in _CACHE:
        return _CACHE["ns"]
    d = reference_dir()
    if d is None:
        _CACHE["ns"] = None
        return None
    from oracle.decode_ref import uxxxx_to_utf8
    saved = {k: sys.modules.get(k) for k in ("textutils", "alphabet", "decoder", "models", "models.cnnlstm")}
    stub = types.ModuleType("textutils")
    stub.uxxxx_to_utf8 = uxxxx_to_utf8
    sys.modules["textutils"] = stub
    for k in ("alphabet", "decoder", "models", "models.cnnlstm"):
        sys.modules.pop(k, None)
    sys.path.insert(0, d)
    try:
        import importlib
        alphabet = importlib.import_module("alphabet")
        decoder = importlib.import_module("decoder")
        cnnlstm = importlib.import_module("models.cnnlstm")
        ns = types.SimpleNamespace(CnnOcrModel=cnnlstm.CnnOcrModel, ArgmaxDecoder=decoder.ArgmaxDecoder,
                                   Alphabet=alphabet.Alphabet, where=d)
    finally:
        sys.path.remove(d)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _CACHE["ns"] = ns
    return ns


def make_model(ns, hp, chars, state_dict=None, seed=7):
    torch.manual_seed(seed)
    model = ns.CnnOcrModel(alphabet=ns.Alphabet(chars), gpu=False, multigpu=False, verbose=False, **hp)
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    return model


def train_step(model, optimizer, batch):
    """The reference's train() (train_cnn_lstm.py:131-150) on the CPU: zero_grad, forward, CTC, backward, per-tensor
    clamp to [-5, 5], Adam step.  Returns the loss."""
    import torch.nn.functional as F
    x, target, widths, target_widths = batch[:4]
    optimizer.zero_grad()
    out, out_lens = model(x, widths)
    loss = F.ctc_loss(out.log_softmax(2), target.long(), out_lens.long(), target_widths.long(), blank=0,
                      reduction="sum", zero_infinity=True)
    loss.backward()
    for p in model.parameters():
        if p.grad is not None:
            p.grad.data.clamp_(min=-5, max=5)
    optimizer.step()
    return float(loss.item())


def decode_batch(model, decoder, x, widths, uxxxx=True):
    """decode_testset.py:92-101,166: eval forward under no_grad + ArgmaxDecoder.decode."""
    with torch.no_grad():
        out, lens = model(x, widths)
    return decoder.decode(out, lens, uxxxx=uxxxx)
