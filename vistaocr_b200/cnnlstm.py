"""`CnnOcrModel` with the reference's surface (reference src/models/cnnlstm.py:36-296,479-541) on top of the
hand-written CUDA kernels in csrc/.

Same keyword-only constructor, attributes, `forward(x[B,C,H,W], widths[B]) -> (logits[T',B,A], lens int32 CPU)`,
`state_dict` names/shapes/dtypes, `FromSavedWeights`, `get_hyper_params`, `cnn_input_size_to_output_size`,
`decode_without_lm`.  The torch.nn sub-modules are kept ONLY as parameter/buffer containers (so checkpoints,
`.parameters()`, `.train()/.eval()` and `cnn[6]._random_samples` behave exactly like the reference); `forward` never
calls them - every stage runs through vistaocr_b200.ops.  Differences, all deliberate:
  * the model always runs on the GPU (there is no CPU code path) and the whole model is one replica per process:
    `multigpu` is accepted but nn.DataParallel is never applied (data parallelism is one process per GPU + NCCL,
    vistaocr_b200/optim.py::FlatGradReducer + sharding.py); checkpoints written by a DataParallel model (`cnn.module.N.*` keys) still load;
  * LM decoding (`init_lm`, `decode_with_lm*`, needs the external EESEN decoder) is out of scope and raises.
"""
import logging

import torch
import torch.nn as nn

from . import ops
from .decoder import ArgmaxDecoder

logger = logging.getLogger("root")

_CONV_PLAN = ((64, False), (64, True), (128, False), (128, True), (256, False), (256, False), (256, False))
# (out channels, followed by FractionalMaxPool2d?)  -> module indices 0..22 of reference `cnn`


class CnnOcrModel(nn.Module):
    def get_hyper_params(self):
        return self.hyper_params

    @classmethod
    def FromSavedWeights(cls, weight_file, verbose=True, gpu=None):
        # reference cnnlstm.py:40-71; snapshots pickle the Alphabet object -> weights_only=False
        weights = torch.load(weight_file, map_location=lambda storage, loc: storage, weights_only=False)
        if verbose:
            logger.info("Loading model from: %s" % weight_file)
            logger.info("\tFrom iteration: %d" % weights["iteration"])
            logger.info("\tWithout LM: Val CER: %.2f\tWER: %.2f" % (100 * weights["val_cer"],
                                                                      100 * weights["val_wer"]))
            logger.info("\tModel Hyperparams = %s" % str(weights["model_hyper_params"]))
        hp = weights["model_hyper_params"]
        if gpu is not None:
            hp["gpu"] = gpu
        hp["verbose"] = verbose
        model = cls(**hp)
        model.rtl = weights["rtl"] if "rtl" in weights else True
        model.load_state_dict(weights["state_dict"], strict=True)
        return model

    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) > 0:
            raise Exception("Only keyword arguments allowed in CnnOcrModel")
        self.hyper_params = kwargs.copy()
        # this model never wraps `cnn` in nn.DataParallel, so its state_dict carries `cnn.N.*` keys: record that in the
        # hyper-parameters a snapshot stores, or the reference's FromSavedWeights (strict=True) would rebuild a
        # DataParallel model expecting `cnn.module.N.*` (reference cnnlstm.py:40-71,198-199)
        self.hyper_params["multigpu"] = False
        self.input_line_height = kwargs["input_line_height"]
        self.rds_line_height = kwargs["rds_line_height"]
        self.alphabet = kwargs["alphabet"]
        self.lstm_input_dim = kwargs["lstm_input_dim"]
        self.num_lstm_layers = kwargs["num_lstm_layers"]
        self.num_lstm_hidden_units = kwargs["num_lstm_hidden_units"]
        self.p_lstm_dropout = kwargs["p_lstm_dropout"]
        self.num_in_channels = kwargs.get("num_in_channels", 1)
        self.gpu = kwargs.get("gpu", True)
        self.multigpu = kwargs.get("multigpu", True)
        self.verbose = kwargs.get("verbose", True)
        self.lattice_decoder = None

        if self.rds_line_height > self.input_line_height:
            raise Exception("rapid-downsample line height must be less than or equal to input line height")
        if self.input_line_height % self.rds_line_height != 0:
            raise Exception("rapid-downsample line height must evenly divide input line height by a power of 2")
        self.num_rds_layers = 0
        lh = self.input_line_height
        while lh > self.rds_line_height:
            if lh % 2 != 0:
                raise Exception("rapid-downsample line height must evenly divide input line height by a power of 2")
            self.num_rds_layers += 1
            lh //= 2
        if lh != self.rds_line_height:
            raise Exception("rapid-downsample line height must evenly divide input line height by a power of 2")

        # ---- parameter containers, same names / order as the reference so state_dict and init RNG order match ----
        self.rapid_ds = nn.Sequential()
        c_in = self.num_in_channels
        for i in range(self.num_rds_layers):
            self.rapid_ds.add_module("%02d-conv" % i, nn.Conv2d(c_in, 16, kernel_size=3, padding=1))
            self.rapid_ds.add_module("%02d-relu" % i, nn.ReLU(inplace=True))
            self.rapid_ds.add_module("%02d-pool" % i, nn.MaxPool2d(2, stride=2))
            c_in = 16
        layers = []
        self._conv_idx, self._pool_idx = [], []
        for c_out, pooled in _CONV_PLAN:
            self._conv_idx.append(len(layers))
            layers += [nn.Conv2d(c_in, c_out, kernel_size=3, padding=1), nn.BatchNorm2d(c_out), nn.ReLU(inplace=True)]
            if pooled:
                self._pool_idx.append(len(layers))
                layers.append(nn.FractionalMaxPool2d(2, output_ratio=(0.5, 0.7)))
            c_in = c_out
        self.cnn = nn.Sequential(*layers)
        cnn_out_h, _ = self.cnn_input_size_to_output_size((self.input_line_height, 20))
        self.cnn_out_h, self.cnn_out_c = cnn_out_h, c_in
        self.bridge_layer = nn.Sequential(nn.Linear(c_in * cnn_out_h, self.lstm_input_dim), nn.ReLU(inplace=True))
        self.lstm = nn.LSTM(self.lstm_input_dim, self.num_lstm_hidden_units, num_layers=self.num_lstm_layers,
                            dropout=self.p_lstm_dropout, bidirectional=True)
        self.prob_layer = nn.Sequential(nn.Linear(2 * self.num_lstm_hidden_units, len(self.alphabet)))

        for param in self.parameters():  # reference cnnlstm.py:158-159: EVERY parameter, BN gamma/beta included
            torch.nn.init.uniform_(param, -0.08, 0.08)

        if self.verbose:
            total = sum(p.numel() for p in self.parameters())
            logger.info("Total Model Params = %d" % total)
            logger.info("\tCNN Params = %d" % sum(p.numel() for p in self.cnn.parameters()))
            logger.info("\tLSTM Params = %d" % sum(p.numel() for p in self.lstm.parameters()))
        if self.gpu and torch.cuda.is_available():
            self.cuda()
        self._decoder = ArgmaxDecoder(self.alphabet)

    # ---- size bookkeeping (reference cnnlstm.py:211-260) -------------------------------------------------------
    def cnn_output_num_channels(self):
        return _CONV_PLAN[-1][0]

    def cnn_input_size_to_output_size(self, in_size):
        return ops.out_hw(in_size[0], in_size[1], self.num_rds_layers)

    # ---- checkpoints written under nn.DataParallel carry `cnn.module.` (reference utils/decode.py:58-71) -------
    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        for k in [k for k in state_dict if k.startswith(prefix + "cnn.module.")]:
            state_dict[prefix + "cnn." + k[len(prefix + "cnn.module."):]] = state_dict.pop(k)
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    # ---- forward (reference cnnlstm.py:268-296) ------------------------------------------------------------------
    def _pool_samples(self, module, n, c, device):
        s = getattr(module, "_random_samples", None)
        if s is None:  # F.fractional_max_pool2d draws rand(N, C, 2) per call, in train AND eval
            return torch.rand((n, c, 2), dtype=torch.float32, device=device)
        cache = getattr(module, "_vocr_samples_dev", None)  # injected samples: uploaded once, not per call
        if cache is None or cache[0] is not s or cache[1] != s._version or cache[2].device != device:
            cache = module._vocr_samples_dev = (s, s._version, s.to(device=device, dtype=torch.float32))
        return cache[2]

    def forward(self, x, actual_minibatch_widths):
        if not x.is_cuda:
            raise ops._lib.VocrError("CnnOcrModel.forward needs a CUDA input: there is no CPU path")
        b, c, h, w = x.shape
        x = x.float()
        # NCHW -> NHWC (free for one input channel)
        feat = x.reshape(b, h, w, 1) if c == 1 else x.permute(0, 2, 3, 1).contiguous()
        for i in range(self.num_rds_layers):
            conv = self.rapid_ds[3 * i]
            feat = ops.rapid_ds(feat, conv.weight, conv.bias)
        n_blocks = len(self._conv_idx)
        for k, ci in enumerate(self._conv_idx):
            conv, bn = self.cnn[ci], self.cnn[ci + 1]
            training = self.training and bn.training
            last = k == n_blocks - 1
            feat = ops.conv_bn_relu(feat, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean,
                                    bn.running_var, training, bn.momentum, bn.eps, seq_layout=last,
                                    planes=not last and not _CONV_PLAN[k][1],
                                    allow_planes_only=not last and not _CONV_PLAN[k][1] and _CONV_PLAN[k + 1][0] % 64 == 0)
            if training:
                bn.num_batches_tracked += 1
            if _CONV_PLAN[k][1]:
                pool = self.cnn[ci + 3]
                feat = ops.fracpool(feat, self._pool_samples(pool, b, feat.shape[3], feat.device))
        # feat: [W', B, h'*C] with feature index y*C + c; the reference orders features c*h' + y (cnnlstm.py:275-278)
        wf = feat.shape[0]
        hh, cc = self.cnn_out_h, self.cnn_out_c
        lin = self.bridge_layer[0]
        w_bridge = lin.weight.view(-1, cc, hh).permute(0, 2, 1).reshape(-1, hh * cc)
        seq = ops.linear(feat.view(wf * b, hh * cc), w_bridge, lin.bias, relu=True,
                         x_bound=getattr(feat, "_vocr_bound", None)).view(wf, b, -1)

        widths = actual_minibatch_widths.tolist() if torch.is_tensor(actual_minibatch_widths) \
            else list(actual_minibatch_widths)
        lens = [self.cnn_input_size_to_output_size((self.input_line_height, int(wd)))[1] for wd in widths]
        tmax = max(lens) if lens else 0
        if tmax > wf or min(lens) < 0:
            raise ops._lib.VocrError("actual_minibatch_widths exceed the padded batch width")
        if any(lens[i] < lens[i + 1] for i in range(len(lens) - 1)):
            # pack_padded_sequence(enforce_sorted=True) raises in the reference as well
            raise RuntimeError("`actual_minibatch_widths` must be sorted in decreasing order")
        lens_cpu = torch.tensor(lens, dtype=torch.int32)
        # (graphs.py replays this forward with other lengths of the same maximum: it supplies the device copy itself)
        lens_dev = getattr(self, "_lens_dev_override", None)
        if lens_dev is None:
            lens_dev = lens_cpu.to(x.device, non_blocking=True)
        seq = seq[:tmax]
        hid = self.num_lstm_hidden_units
        dropping = self.training and self.p_lstm_dropout > 0
        masks = getattr(self, "_dropout_masks", None)  # injected keep masks [L-1] x uint8 [T',B,2H] (parity tests)
        bound = None
        for l in range(self.num_lstm_layers):
            g = lambda n: getattr(self.lstm, n % l)
            w_ih = torch.cat([g("weight_ih_l%d"), g("weight_ih_l%d_reverse")], 0)
            w_hh = torch.stack([g("weight_hh_l%d"), g("weight_hh_l%d_reverse")], 0)
            bias = torch.cat([g("bias_ih_l%d") + g("bias_hh_l%d"), g("bias_ih_l%d_reverse") + g("bias_hh_l%d_reverse")])
            seq = ops.bilstm_layer(seq, w_ih, w_hh, bias, lens_dev, tmax, save=torch.is_grad_enabled(), x_bound=bound)
            bound = ops.const_scalar(x.device, 1.0)  # |h| = |o tanh(c)| < 1
            if dropping and l < self.num_lstm_layers - 1:
                # nn.LSTM(dropout=p) (reference cnnlstm.py:148-149): between layers, training only
                if masks is not None:
                    seq = ops.dropout(seq, self.p_lstm_dropout, mask=masks[l])
                else:
                    seq = ops.dropout(seq, self.p_lstm_dropout, rng=self._rng_state(x.device), site=l)
                bound = ops.const_scalar(x.device, 1.0 / (1.0 - self.p_lstm_dropout))
        if dropping and masks is None and self.num_lstm_layers > 1:
            ops.rng_advance(self._rng_state(x.device), self.num_lstm_layers - 1)
        prob = self.prob_layer[0]
        fd = getattr(self, "_fused_decode", None)  # graphs.GraphedDecoder: frame labels straight from the GEMM epilogue
        if fd is not None and not self.training and ops.linear_argmax_supported(2 * hid, prob.weight.shape[0]):
            fd["path"] = ops.linear_argmax(seq.reshape(tmax * b, 2 * hid), prob.weight, prob.bias, lens_dev, tmax, b,
                                           fd["thresh"], x_bound=bound)
            return None, lens_cpu
        logits = ops.linear(seq.reshape(tmax * b, 2 * hid), prob.weight, prob.bias, x_bound=bound).view(tmax, b, -1)
        return logits, lens_cpu

    # ---- inter-layer dropout stream ---------------------------------------------------------------------------------
    def _rng_state(self, device):
        st = getattr(self, "_dropout_rng", None)
        if st is None or st.device != device:
            st = self._dropout_rng = ops.new_rng_state(device)
        return st

    def set_dropout_seed(self, seed, offset=0):
        """Pin the Philox stream of the inter-layer dropout: the next training forward uses offsets `offset + l` for the
        output of LSTM layer l (ops.dropout_keep_mask reproduces the masks), and advances the offset by L-1."""
        dev = next(self.parameters()).device
        self._dropout_rng = ops.new_rng_state(dev, seed, offset)

    # ---- greedy decode (reference cnnlstm.py:479-541 == decoder.py:116-185) --------------------------------------
    def decode_without_lm(self, model_output, batch_actual_timesteps, uxxxx=False):
        return self._decoder.decode(model_output, batch_actual_timesteps, uxxxx=uxxxx)

    def init_lm(self, *args, **kwargs):
        raise NotImplementedError("LM decoding needs the external EESEN lattice decoder: out of scope (DESIGN.md)")

    decode_with_lm = decode_with_lm_mt = init_lm
