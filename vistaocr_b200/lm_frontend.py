"""Front end of LM decoding (reference src/decoder.py:11-101, `LmDecoder.__init__` unit bookkeeping and the first
half of `LmDecoder.decode`): log-softmax the model output, remap the model alphabet onto the LM's units (units the
model does not know get log(1e-10)), slice every line to its valid frames - one fused CUDA kernel, one compact D2H
copy, instead of a full-logit D2H plus a NumPy scatter per line.  The lattice decoder that consumes these matrices is
the external EESEN binding and stays out of scope; `LmFrontend.log_probs_for_lm` returns exactly the float64
`probs_remapped` arrays the reference submits to it.
"""
import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream

FILL = float(np.log(1e-10))


class LmFrontend:
    def __init__(self, alphabet, lm_units):
        """lm_units: iterable of unit strings in LM order WITHOUT the leading '<ctc-blank>' (the reference reads them
        from the `units.txt` of the LM, decoder.py:31-34) - or a path to that file."""
        self.alphabet = alphabet
        if isinstance(lm_units, str):
            with open(lm_units, "r") as fh:
                lm_units = [line.strip().split(" ")[0] for line in fh]
        self.lmidx_to_char = ["<ctc-blank>"] + list(lm_units)
        self.lmchar_to_idx = dict(zip(self.lmidx_to_char, range(len(self.lmidx_to_char))))
        self.add_to_blank_char = []
        inv = np.full(len(self.lmidx_to_char), -1, np.int32)
        for model_idx in range(len(alphabet.idx_to_char)):
            ch = alphabet.idx_to_char[model_idx]
            if ch not in self.lmchar_to_idx:
                self.add_to_blank_char.append(ch)
                continue
            inv[self.lmchar_to_idx[ch]] = model_idx  # numpy fancy assignment: the last model index wins
        self.inv = inv
        self._inv_dev = None

    def log_probs_device(self, model_output, batch_actual_timesteps):
        """-> (out float64 CUDA [sum len, U], row_offsets list, lens list)."""
        _lib.require_cuda(model_output, "model_output", torch.float32)
        T, B, A = model_output.shape
        dev = model_output.device
        lens = [max(0, min(int(v), T)) for v in torch.as_tensor(batch_actual_timesteps).tolist()]
        offs = np.zeros(B + 1, np.int64)
        offs[1:] = np.cumsum(lens)
        if self._inv_dev is None or self._inv_dev.device != dev:
            self._inv_dev = torch.from_numpy(self.inv).to(dev)
        U = len(self.lmidx_to_char)
        out = torch.empty((int(offs[-1]), U), dtype=torch.float64, device=dev)
        d_lens = torch.tensor(lens, dtype=torch.int32).to(dev, non_blocking=True)
        d_offs = torch.from_numpy(offs[:-1].copy()).to(dev, non_blocking=True)
        st = lib().vocr_lm_frontend_f32(ptr(model_output), T, B, A, ptr(d_lens), ptr(d_offs), ptr(self._inv_dev), U,
                                        FILL, ptr(out), stream())
        check(st, "vocr_lm_frontend_f32")
        return out, offs, lens

    def log_probs_for_lm(self, model_output, batch_actual_timesteps):
        """The list of per-line float64 arrays [len_b, |units|] that LmDecoder.decode submits (decoder.py:99-106)."""
        if not model_output.is_cuda:
            model_output = model_output.cuda(non_blocking=True)
        out, offs, _ = self.log_probs_device(model_output.detach().float().contiguous(), batch_actual_timesteps)
        host = out.cpu().numpy()
        return [host[offs[b]:offs[b + 1]] for b in range(len(offs) - 1)]
