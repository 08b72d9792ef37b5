"""Drop-in for the reference's ``networks.decoder_z_lb_both_leaky_n_sdf.Decoder``
(networks/decoder_z_lb_both_leaky_n_sdf.py:14-457): same constructor keywords, same
state-dict keys (``[module.]lin{i}.{weight_g,weight_v,bias}``, ``hnet_lin{i}...``), same
``forward(net_input, use_net, return_occ, return_normals, occ_threshold)`` contract in eval
mode -- evaluated by the fused sm_100a kernels behind the C ABI instead of ATen.

Only the shipped configuration has a kernel (weight_norm, one latent re-injection layer,
xyz fed to the sub-net, no tanh layer, no LayerNorm, hidden width 512); anything else raises
NotImplementedError -- there is no slow path."""
import os

import numpy as np
import torch
import torch.nn as nn

from deepjoin_b200 import _lib


class _WNLinear(nn.Module):
    """Parameters of ``nn.utils.weight_norm(nn.Linear(i, o))``: weight_g [o,1], weight_v [o,i], bias [o]."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        lin = nn.Linear(in_dim, out_dim)
        self.weight_v = nn.Parameter(lin.weight.detach().clone())
        self.weight_g = nn.Parameter(lin.weight.detach().norm(dim=1, keepdim=True))
        self.bias = nn.Parameter(lin.bias.detach().clone())

    def folded(self):
        v = self.weight_v.detach().float()
        return (v * (self.weight_g.detach().float() / v.norm(dim=1, keepdim=True))), self.bias.detach().float()


class _Linear(nn.Module):
    def __init__(self, in_dim, out_dim):
        super().__init__()
        lin = nn.Linear(in_dim, out_dim)
        self.weight = nn.Parameter(lin.weight.detach().clone())
        self.bias = nn.Parameter(lin.bias.detach().clone())

    def folded(self):
        return self.weight.detach().float(), self.bias.detach().float()


class Decoder(nn.Module):
    def __init__(self, latent_size, tool_latent_size, dims, num_dims=2, dropout=None, dropout_prob=0.0,
                 norm_layers=(), latent_in=(), latent_in_inflate=False, weight_norm=False, xyz_in_all=False,
                 use_tanh=False, latent_dropout=False, do_code_regularization=False, subnet_dims=None,
                 subnet_dropout=None, subnet_norm=None, subnet_xyz=False, subnet_latent_in=(),
                 subnet_latent_in_inflate=False, return_occ=False):
        super().__init__()
        bad = []
        if num_dims != 3:
            bad.append("num_dims=%r" % (num_dims,))
        if latent_in_inflate or subnet_latent_in_inflate or tuple(subnet_latent_in):
            bad.append("latent_in_inflate / subnet_latent_in")
        if xyz_in_all:
            bad.append("xyz_in_all")
        if use_tanh:
            bad.append("use_tanh")
        if latent_dropout:
            bad.append("latent_dropout")
        if not subnet_xyz:
            bad.append("subnet_xyz=False")
        if len(tuple(latent_in)) != 1:
            bad.append("latent_in=%r" % (latent_in,))
        if any(d != 512 for d in list(dims) + list(subnet_dims or [])):
            bad.append("hidden width != 512")
        if not weight_norm and (len(tuple(norm_layers or ())) or len(tuple(subnet_norm or ()))):
            bad.append("LayerNorm (weight_norm=False with norm_layers)")
        if bad:
            raise NotImplementedError("deepjoin_b200 Decoder: unsupported configuration: " + ", ".join(bad))

        # attributes the reference exposes (decoder :42-57,139)
        self.latent_size = latent_size
        self.tool_latent_size = tool_latent_size
        self.num_dims = num_dims
        self.dropout = dropout
        self.dropout_prob = dropout_prob
        self.norm_layers = norm_layers
        self.latent_in = latent_in
        self.weight_norm = weight_norm
        self.pts_in_all = xyz_in_all
        self.use_tanh = use_tanh
        self.do_code_regularization = do_code_regularization
        self.latent_dropout = latent_dropout
        self.subnet_xyz = subnet_xyz
        self._return_occ = return_occ

        dims = [latent_size + num_dims] + list(dims) + [5]
        hdims = [tool_latent_size + num_dims] + list(subnet_dims) + [5]
        self.num_layers = len(dims)
        self.hnet_num_layers = len(hdims)
        self.hnet_norm_layers = subnet_norm
        self.hnet_dropout = subnet_dropout
        for layer in range(self.hnet_num_layers - 1):
            wn = weight_norm and (layer in (subnet_norm or ()))
            setattr(self, "hnet_lin%d" % layer, (_WNLinear if wn else _Linear)(hdims[layer], hdims[layer + 1]))
        for layer in range(self.num_layers - 1):
            out_dim = dims[layer + 1] - dims[0] if (layer + 1) in latent_in else dims[layer + 1]
            wn = weight_norm and (layer in (norm_layers or ()))
            setattr(self, "lin%d" % layer, (_WNLinear if wn else _Linear)(dims[layer], out_dim))

        self.precision = os.environ.get("DEEPJOIN_B200_PRECISION", "fp16x2")
        self._pack = None
        self._pack_key = None

    # ------------------------------------------------------------------ packing
    def _key(self, device):
        return (device.index,) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def pack(self, device=None):
        """Fold weight-norm once (W = g*v/||v||_row), upload, build the tensor-core weight stream."""
        if device is None:
            device = next(self.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("deepjoin_b200 Decoder runs on CUDA (sm_100a) only; call .cuda() first")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        key = self._key(device)
        if self._pack is None or self._pack_key != key:
            f = [getattr(self, "lin%d" % l).folded() for l in range(self.num_layers - 1)]
            h = [getattr(self, "hnet_lin%d" % l).folded() for l in range(self.hnet_num_layers - 1)]
            desc = _lib.NetDesc(self.latent_size, self.tool_latent_size, 512, len(f), len(h),
                                int(tuple(self.latent_in)[0]), 0.01)
            cpu = lambda t: t.cpu().contiguous().numpy()
            if self._pack is not None:
                self._pack.close()
            self._pack = _lib.Pack(desc, [cpu(w) for w, _ in f], [cpu(b) for _, b in f],
                                   [cpu(w) for w, _ in h], [cpu(b) for _, b in h], device.index)
            self._pack_key = key
        return self._pack

    # ------------------------------------------------------------------ queries
    def _prec(self, precision=None):
        p = precision or self.precision
        if p not in _lib.PREC:
            raise ValueError("precision must be one of %s" % sorted(_lib.PREC))
        return _lib.PREC[p]

    def _check_eval(self):
        if self.training:
            raise NotImplementedError("deepjoin_b200 Decoder is inference-only: call .eval() "
                                      "(the reference's training-mode branches / dropout have no kernel)")

    def query(self, code, xyz, use_net=1, return_occ=None, sigmoid=False, precision=None):
        """One latent code [L+Lb] for all points xyz [N,3] (CUDA f32) -> [N,1]."""
        self._check_eval()
        if return_occ is None:
            return_occ = self._return_occ
        if use_net not in (0, 1, 2, 3):
            raise RuntimeError("Requested output from non-existent network: {}".format(use_net))
        pack = self.pack(xyz.device)
        code = code.detach().reshape(-1).to(device=xyz.device, dtype=torch.float32).contiguous()
        xyz = xyz.detach().to(torch.float32).contiguous()
        out = torch.empty((xyz.shape[0], 1), device=xyz.device, dtype=torch.float32)
        with torch.cuda.device(xyz.device):
            _lib.check(_lib.lib().dj_query_points(pack.handle, code.data_ptr(), xyz.data_ptr(), xyz.shape[0],
                                                  int(use_net), int(bool(return_occ)), int(bool(sigmoid)),
                                                  self._prec(precision), out.data_ptr(), _lib.stream_ptr()))
        return out

    def query_all(self, code, xyz, precision=None):
        """use_net=None: the 12-tuple of decoder :331-344."""
        self._check_eval()
        pack = self.pack(xyz.device)
        code = code.detach().reshape(-1).to(device=xyz.device, dtype=torch.float32).contiguous()
        xyz = xyz.detach().to(torch.float32).contiguous()
        out = torch.empty((xyz.shape[0], 20), device=xyz.device, dtype=torch.float32)
        with torch.cuda.device(xyz.device):
            _lib.check(_lib.lib().dj_query_all(pack.handle, code.data_ptr(), xyz.data_ptr(), xyz.shape[0],
                                               self._prec(precision), out.data_ptr(), _lib.stream_ptr()))
        o = out
        return (o[:, 0:1], o[:, 1:2], o[:, 2:3], o[:, 3:4], o[:, 4:5], o[:, 5:6], o[:, 6:7], o[:, 7:8],
                o[:, 8:11], o[:, 11:14], o[:, 14:17], o[:, 17:20])

    def forward(self, net_input, use_net=None, return_occ=None, return_normals=False, occ_threshold=0.5):
        if return_normals:
            raise NotImplementedError
        if occ_threshold != 0.5:
            raise NotImplementedError("occ_threshold other than 0.5")
        if use_net is not None and use_net not in (0, 1, 2, 3):
            raise RuntimeError("Requested output from non-existent network: {}".format(use_net))
        if net_input.requires_grad:
            raise NotImplementedError("autograd through Decoder.forward is not provided: latent optimisation "
                                      "runs in deepjoin_b200.reconstruct.reconstruct (fused forward+backward)")
        n_code = self.latent_size + self.tool_latent_size
        if net_input.dim() != 2 or net_input.shape[1] != n_code + self.num_dims:
            raise ValueError("net_input must be [N, %d]" % (n_code + self.num_dims))
        codes = net_input[:, :n_code]
        if net_input.shape[0] > 1 and not bool((codes == codes[0:1]).all()):
            raise NotImplementedError("rows with different latent codes in one call (training-style batches)")
        xyz = net_input[:, n_code:]
        if use_net is None:
            return self.query_all(codes[0], xyz)
        return self.query(codes[0], xyz, use_net=use_net, return_occ=return_occ)

    def load_state_dict(self, state_dict, strict=True):
        """Accepts the DataParallel ``module.`` prefix the shipped checkpoints carry
        (core/utils_multiprocessing.py:23-28)."""
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        return super().load_state_dict(sd, strict=strict)
