"""Drop-in for the reference's baseline ``networks.decoder_deepsdf.Decoder``
(networks/decoder_deepsdf.py:6-106): the vanilla DeepSDF auto-decoder -- ONE net shaped like
DeepJoin's fnet (latent re-injection at one layer), ReLU instead of LeakyReLU (:64), one output,
final tanh (:103-104) -- evaluated by the same fused sm_100a kernels: the pack is built with
LeakyReLU slope 0, the single head row padded to the five rows the kernels' head tile carries, and
a zero stand-in for the (unused, never evaluated) second net; ``use_net=0`` selects tanh(head[0]).

Only the shipped configuration (experiments/mugs_deepsdf/specs.json: weight_norm, one latent_in
layer, hidden width 512, no xyz_in_all / use_tanh / latent_dropout / LayerNorm) has a kernel."""
import os

import numpy as np
import torch
import torch.nn as nn

from deepjoin_b200 import _lib
from .decoder_z_lb_both_leaky_n_sdf import _Linear, _WNLinear


class Decoder(nn.Module):
    def __init__(self, latent_size, dims, dropout=None, dropout_prob=0.0, norm_layers=(), latent_in=(), weight_norm=False,
                 xyz_in_all=None, use_tanh=False, latent_dropout=False):
        super().__init__()
        bad = []
        if xyz_in_all:
            bad.append("xyz_in_all")
        if use_tanh:
            bad.append("use_tanh")
        if latent_dropout:
            bad.append("latent_dropout")
        if len(tuple(latent_in)) != 1:
            bad.append("latent_in=%r" % (latent_in,))
        if any(d != 512 for d in dims):
            bad.append("hidden width != 512")
        if not weight_norm and len(tuple(norm_layers or ())):
            bad.append("LayerNorm (weight_norm=False with norm_layers)")
        if bad:
            raise NotImplementedError("deepjoin_b200 DeepSDF Decoder: unsupported configuration: " + ", ".join(bad))
        self.latent_size = latent_size
        self.tool_latent_size = 1  # stand-in for the second net the kernels' pack carries (never evaluated)
        self.num_dims = 3
        self.norm_layers = norm_layers
        self.latent_in = latent_in
        self.latent_dropout = latent_dropout
        self.xyz_in_all = xyz_in_all
        self.weight_norm = weight_norm
        self.use_tanh = use_tanh
        self.dropout = dropout
        self.dropout_prob = dropout_prob
        self._return_occ = False
        dims = [latent_size + 3] + list(dims) + [1]
        self.num_layers = len(dims)
        for layer in range(self.num_layers - 1):
            out_dim = dims[layer + 1] - dims[0] if (layer + 1) in latent_in else dims[layer + 1]
            wn = weight_norm and (layer in (norm_layers or ()))
            setattr(self, "lin%d" % layer, (_WNLinear if wn else _Linear)(dims[layer], out_dim))
        self.precision = os.environ.get("DEEPJOIN_B200_PRECISION", "fp16x2")
        # latent optimisation (reconstruct_deepsdf.reconstruct): the baseline's objective is a sign-gated L1, so a
        # forward rounding error flips whole per-sample gradients; IEEE-half forward operands (8x finer than bf16)
        # keep the tensor-core gradient within ~1e-2 of fp32 even at small batches
        self.latent_precision = None  # None: "fp16" when precision is "bf16", else precision (latent_prec())
        self._pack = None
        self._pack_key = None

    def pack(self, device=None):
        if device is None:
            device = next(self.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("deepjoin_b200 Decoder runs on CUDA (sm_100a) only; call .cuda() first")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        key = (device.index,) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._pack is None or self._pack_key != key:
            f = [getattr(self, "lin%d" % l).folded() for l in range(self.num_layers - 1)]
            w, b = f[-1]  # [1, 512], [1] -> the kernels' 5-row head, rows 1..4 zero
            w, b = w.cpu(), b.cpu()
            f[-1] = (torch.cat([w, torch.zeros(4, w.shape[1])], 0), torch.cat([b, torch.zeros(4)], 0))
            hshape = [(512, 4)] + [(512, 512)] * 4 + [(5, 512)]
            cpu = lambda t: t.cpu().contiguous().numpy()
            desc = _lib.NetDesc(self.latent_size, 1, 512, len(f), len(hshape), int(tuple(self.latent_in)[0]), 0.0)
            if self._pack is not None:
                self._pack.close()
            self._pack = _lib.Pack(desc, [cpu(w) for w, _ in f], [cpu(b) for _, b in f],
                                   [np.zeros(s, np.float32) for s in hshape], [np.zeros(s[0], np.float32) for s in hshape],
                                   device.index)
            self._pack_key = key
        return self._pack

    def _prec(self, precision=None):
        p = precision or self.precision
        if p not in _lib.PREC:
            raise ValueError("precision must be one of %s" % sorted(_lib.PREC))
        return _lib.PREC[p]

    def latent_prec(self):
        return self.latent_precision or ("fp16" if self.precision == "bf16" else self.precision)

    def _check_eval(self):
        if self.training:
            raise NotImplementedError("deepjoin_b200 Decoder is inference-only: call .eval()")

    def code192(self, latent):
        """[L] latent -> the [L + 1] code the pack expects (zero stand-in for the second net's latent)."""
        latent = latent.detach().reshape(-1).to(torch.float32)
        return torch.cat([latent, latent.new_zeros(1)])

    def query(self, code, xyz, use_net=0, return_occ=None, sigmoid=False, precision=None):
        """One latent code for all points xyz [N,3] (CUDA f32) -> tanh(sdf) [N,1].  ``code`` is [L] or [L+1]."""
        self._check_eval()
        if use_net != 0:
            raise RuntimeError("Requested output from non-existent network: {}".format(use_net))
        pack = self.pack(xyz.device)
        code = code.detach().reshape(-1)
        if code.numel() == self.latent_size:
            code = self.code192(code)
        code = code.to(device=xyz.device, dtype=torch.float32).contiguous()
        xyz = xyz.detach().to(torch.float32).contiguous()
        out = torch.empty((xyz.shape[0], 1), device=xyz.device, dtype=torch.float32)
        with torch.cuda.device(xyz.device):
            _lib.check(_lib.lib().dj_query_points(pack.handle, code.data_ptr(), xyz.data_ptr(), xyz.shape[0], 0, 0,
                                                  int(bool(sigmoid)), self._prec(precision), out.data_ptr(), _lib.stream_ptr()))
        return out

    def forward(self, input):
        """input [N, L+3] (decoder_deepsdf.py:70-106), rows sharing one latent code (what decode_sdf feeds it,
        core/utils_deepsdf.py:11-20)."""
        if input.requires_grad:
            raise NotImplementedError("autograd through Decoder.forward is not provided")
        if input.dim() != 2 or input.shape[1] != self.latent_size + 3:
            raise ValueError("input must be [N, %d]" % (self.latent_size + 3))
        codes = input[:, :self.latent_size]
        if input.shape[0] > 1 and not bool((codes == codes[0:1]).all()):
            raise NotImplementedError("rows with different latent codes in one call (training-style batches)")
        return self.query(codes[0], input[:, self.latent_size:])

    def load_state_dict(self, state_dict, strict=True):
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        return super().load_state_dict(sd, strict=strict)
