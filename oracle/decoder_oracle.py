"""CPU restatement (torch, fp32 or fp64) of the DeepJoin decoder hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PINNED against the live
reference through ``tests/golden`` (see tests/test_oracle_pinned.py).

Follows, for the configuration the reference ships (experiments/mugs/specs.json:
weight_norm, latent_in=[4], subnet_xyz, no tanh layer, dropout inactive in eval):

* networks/decoder_z_lb_both_leaky_n_sdf.py:174-457   Decoder.forward
* core/reconstruct.py:25-30,45-132                    grid points / decode_samples
* reconstruct.py:19-226                               latent optimisation loop
* core/data.py:15-73,122-176                          per-iteration sampler
"""
import math

import numpy as np
import torch

SLOPE = 0.01  # nn.LeakyReLU() default, decoder :135


def _leaky(x):
    return torch.where(x > 0, x, x * SLOPE)


def fold(sd, dtype=torch.float32, latent_size=128, tool_latent_size=64, hidden=512,
         f_layers=9, h_layers=6, latent_in=4):
    """weight_norm fold W = g * v / ||v||_row  (nn.utils.weight_norm, dim=0; decoder :82-87,117-122).
    Accepts the DataParallel 'module.' prefix (core/utils_multiprocessing.py:23-28)."""
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}

    def lin(name):
        if name + ".weight_g" in sd:
            g = sd[name + ".weight_g"].to(dtype)
            v = sd[name + ".weight_v"].to(dtype)
            w = v * (g / v.norm(dim=1, keepdim=True))
        else:
            w = sd[name + ".weight"].to(dtype)
        return w, sd[name + ".bias"].to(dtype)

    return dict(
        f=[lin("lin%d" % l) for l in range(f_layers)],
        h=[lin("hnet_lin%d" % l) for l in range(h_layers)],
        L=latent_size, Lb=tool_latent_size, latent_in=latent_in, dtype=dtype)


def raw_heads(net, code, xyz, h_pre=False):
    """yC [N,5] (no activation on the last fnet layer, decoder :229) and
    yT [N,5] (LeakyReLU IS applied to the last hnet layer, decoder :259-270)."""
    dt = net["dtype"]
    xyz = xyz.to(dt)
    n = xyz.shape[0]
    code = code.to(dt).reshape(-1)
    z = code[: net["L"]].expand(n, -1)
    zb = code[net["L"]: net["L"] + net["Lb"]].expand(n, -1)
    f_in = torch.cat([z, xyz], 1)
    x = f_in
    nf = len(net["f"])
    for l, (w, b) in enumerate(net["f"]):
        if l == net["latent_in"]:
            x = torch.cat([x, f_in], 1)
        x = x @ w.t() + b
        if l < nf - 1:
            x = _leaky(x)
    yc = x
    x = torch.cat([zb, xyz], 1)
    nh = len(net["h"])
    for l, (w, b) in enumerate(net["h"]):
        x = x @ w.t() + b
        if l == nh - 1 and h_pre:
            return yc, x
        x = _leaky(x)
    return yc, x


def forward_all(net, code, xyz):
    """use_net=None branch (decoder :299-344): 12 outputs, select form."""
    yc, yt = raw_heads(net, code, xyz)
    c_sdf, c_occ, c_n = torch.tanh(yc[:, 0:1]), yc[:, 1:2], torch.tanh(yc[:, 2:5])
    t_sdf, t_occ, t_n = torch.tanh(yt[:, 0:1]), yt[:, 1:2], torch.tanh(yt[:, 2:5])
    sc, st = torch.sigmoid(c_occ), torch.sigmoid(t_occ)
    b_occ = sc * (1 - st)
    sel_b = (st > 0.5) | (c_sdf < -t_sdf)
    b_sdf = torch.where(sel_b, -t_sdf, c_sdf)
    b_n = torch.where(sel_b, t_n, c_n)
    r_occ = sc * st
    sel_r = (st <= 0.5) | (c_sdf < t_sdf)
    r_sdf = torch.where(sel_r, t_sdf, c_sdf)
    r_n = torch.where(sel_r, -t_n, c_n)
    return (c_occ, b_occ, r_occ, t_occ, c_sdf, b_sdf, r_sdf, t_sdf, c_n, b_n, r_n, t_n)


def forward(net, code, xyz, use_net=None, return_occ=False):
    """Decoder.forward in eval mode (decoder :345-457)."""
    if use_net is None:
        return forward_all(net, code, xyz)
    if use_net not in (0, 1, 2, 3):
        raise RuntimeError("Requested output from non-existent network: {}".format(use_net))
    yc, yt = raw_heads(net, code, xyz)
    c_sdf, c_occ = torch.tanh(yc[:, 0:1]), yc[:, 1:2]
    t_sdf, t_occ = torch.tanh(yt[:, 0:1]), yt[:, 1:2]
    sc, st = torch.sigmoid(c_occ), torch.sigmoid(t_occ)
    if use_net == 0:
        return c_occ if return_occ else c_sdf
    if use_net == 1:
        return sc * (1 - st) if return_occ else torch.max(c_sdf, -t_sdf)
    if use_net == 2:
        return sc * st if return_occ else torch.max(c_sdf, t_sdf)
    return t_occ if return_occ else t_sdf


def linspace_axis(d, padding):
    """np.linspace(0, 1+2p, d) - (0.5+p) in float64 (core/reconstruct.py:68-70)."""
    return np.linspace(0, 1.0 + (padding * 2), d) - (0.5 + padding)


def get_query_points(dims=(256, 256, 256), padding=0.1):
    """core/reconstruct.py:25-30 — np.meshgrid default indexing='xy'."""
    grid_pts = np.meshgrid(*[linspace_axis(d, padding) for d in dims])
    return np.vstack([p.flatten() for p in grid_pts]).T


def decode_samples(net, vec, pts, use_net=1, batch_size=2 ** 16, sigmoid=True, return_occ=False):
    """core/reconstruct.py:89-132: f64 points -> f32 -> decoder -> [P,1] float64."""
    out = []
    for s in range(0, pts.shape[0], batch_size):
        p = torch.from_numpy(pts[s: s + batch_size]).type(torch.float)
        v = forward(net, vec, p, use_net=use_net, return_occ=return_occ)
        if sigmoid:
            v = torch.sigmoid(v)
        out.append(v.detach().cpu().numpy())
    return np.vstack(out).astype(np.float64)


def decode_shape(net, vec, dims, use_net=1, batch_size=2 ** 16, padding=0.0, reshape=True, sigmoid=True,
                 return_occ=False):
    """core/reconstruct.py:45-86."""
    values = decode_samples(net, vec, get_query_points(dims, padding), use_net, batch_size, sigmoid, return_occ)
    if reshape:
        return values.reshape([d for d in dims]).T
    return values


def sdf_to_occ(x):
    """core/reconstruct.py:33-34."""
    return torch.clamp(torch.sgn(-x), -1, 0) + 1


def loss_terms(net, code, xyz, sdf_gt, n_gt, lambda1=1.0, lambda3=1.0, clamp_dist=0.1,
               code_reg_lambda=1e-4, l2reg=True):
    """reconstruct.py:147-201.  Returns (loss, sdf_loss, occ_loss, norm_loss, reg_loss)."""
    dt = net["dtype"]
    outs = forward_all(net, code, xyz)
    b_occ, b_x, b_n = outs[1], outs[5], outs[9]
    b_x = torch.clamp(b_x, -clamp_dist, clamp_dist)
    sdf_gt = sdf_gt.to(dt).reshape(-1, 1)
    occ_gt = sdf_to_occ(sdf_gt)
    zero = torch.zeros((), dtype=dt)
    sdf_loss = torch.nn.functional.l1_loss(b_x, sdf_gt) * lambda3 if lambda3 != 0 else zero
    occ_loss = torch.nn.functional.binary_cross_entropy(b_occ, occ_gt)
    norm_loss = torch.nn.functional.mse_loss(b_n, n_gt.to(dt)) * lambda1 if lambda1 != 0 else zero
    loss = sdf_loss + occ_loss + norm_loss
    reg = zero
    if l2reg:
        L = net["L"]
        reg = (code.reshape(-1)[:L].pow(2).mean() + code.reshape(-1)[L:].pow(2).mean()) * code_reg_lambda
        loss = loss + reg
    return loss, sdf_loss, occ_loss, norm_loss, reg


def latent_grad(net, code, xyz, sdf_gt, n_gt, **kw):
    code = code.detach().clone().to(net["dtype"]).reshape(-1).requires_grad_(True)
    terms = loss_terms(net, code, xyz, sdf_gt, n_gt, **kw)
    terms[0].backward()
    return code.grad.detach(), [float(t.detach()) for t in terms]


def lr_at(e, lr, num_iterations, decreased_by=10):
    """reconstruct.py:38-46,126."""
    return lr * ((1 / decreased_by) ** (e // int(num_iterations / 2)))


def adam_loop(net, code0, batches, num_iterations, lr=5e-3, betas=(0.9, 0.999), eps=1e-8, log_every=10, **kw):
    """reconstruct.py:58,107-226 with an injected sample schedule:
    ``batches(e) -> (xyz[N,3], sdf[N,1], n[N,3])`` f32 tensors.
    Two parameter tensors (latent, break_latent) share one Adam — elementwise,
    so identical to one [192] vector."""
    dt = net["dtype"]
    code = code0.detach().clone().to(dt).reshape(-1)
    m = torch.zeros_like(code)
    v = torch.zeros_like(code)
    log = {k: [] for k in ("epoch", "loss", "sdf_data_loss", "occ_data_loss", "norm_loss", "reg_loss", "mag", "h_mag")}
    L = net["L"]
    for e in range(num_iterations):
        xyz, sdf, n = batches(e)
        g, t = latent_grad(net, code, xyz, sdf, n, **kw)
        if e % log_every == 0:
            log["epoch"].append(e)
            for k, val in zip(("loss", "sdf_data_loss", "occ_data_loss", "norm_loss", "reg_loss"), t):
                log[k].append(val)
            log["mag"].append(float(code[:L].norm()))
            log["h_mag"].append(float(code[L:].norm()))
        step = e + 1
        m = betas[0] * m + (1 - betas[0]) * g
        v = betas[1] * v + (1 - betas[1]) * g * g
        bc1 = 1 - betas[0] ** step
        bc2 = 1 - betas[1] ** step
        step_size = lr_at(e, lr, num_iterations) / bc1
        denom = v.sqrt() / math.sqrt(bc2) + eps
        code = code - step_size * m / denom
    return log, code.reshape(1, -1)


# ---------------------------------------------------------------- sampler
class NoSamplesError(RuntimeError):
    pass


def select_uniform_ratio_samples(pts, values, uniform_ratio=0.5, randomize=True):
    """core/data.py:122-176 (global numpy RNG, same call order)."""
    num_dims = pts.shape[1]
    data = np.hstack((pts, values))
    if uniform_ratio != 0.5:
        n_pts = data.shape[0]
        max_can_select = int(n_pts / 2)
        surface_ends_at = int(n_pts / 2)
        if uniform_ratio > 0.5:
            k = int((max_can_select * (1 - uniform_ratio)) / uniform_ratio)
            sel = np.random.choice(max_can_select, size=(k), replace=False)
            data = np.vstack((data[sel, :], data[surface_ends_at: surface_ends_at + max_can_select, :]))
        else:
            k = int((max_can_select * uniform_ratio) / (1 - uniform_ratio))
            sel = np.random.choice(max_can_select, size=(k), replace=False) + surface_ends_at
            data = np.vstack((data[:max_can_select, :], data[sel, :]))
    if randomize:
        data = data[np.random.permutation(data.shape[0]), :]
    return data[:, :num_dims], data[:, num_dims:]


def select_samples(pts, values, num_samples=None, uniform_ratio=0.5):
    """core/data.py:15-73 for one value column (the only case reconstruct() uses)."""
    if uniform_ratio is not None:
        pts, values = select_uniform_ratio_samples(pts, values, uniform_ratio)
    assert values.shape[1] == 1
    s = values.shape[0] if num_samples is None else num_samples
    d = values[:, 0]
    pos_inds = np.where(d > 0)[0]
    neg_inds = np.where(d <= 0)[0]
    pos_num, neg_num = (s // 2), (s - (s // 2))
    if len(neg_inds) > neg_num:
        start = np.random.randint(len(neg_inds) - neg_num)
        neg_picked = neg_inds[start: start + neg_num]
    else:
        try:
            neg_picked = np.random.choice(neg_inds, neg_num)
        except ValueError:
            raise NoSamplesError
    if len(pos_inds) > pos_num:
        start = np.random.randint(len(pos_inds) - pos_num)
        pos_picked = pos_inds[start: start + pos_num]
    else:
        try:
            pos_picked = np.random.choice(pos_inds, pos_num)
        except ValueError:
            raise NoSamplesError
    idx = np.concatenate([pos_picked, neg_picked])
    return pts[idx, :], values[idx, :]


# ---------------------------------------------------------------------------------------------
# DeepSDF baseline decoder
# ---------------------------------------------------------------------------------------------
def deepsdf_forward(sd, latent, pts, dtype=torch.float32, latent_in=4):
    """Restatement of networks/decoder_deepsdf.py:70-106 for one latent code: x = [z; p]; lin, ReLU (:64,
    no activation after the last layer :90), input re-injected at ``latent_in`` (:83-84), weight_norm folded
    (:44-48), final tanh (:103-104).  Returns [N,1]."""
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    n_layers = 1 + max(int(k[3:].split(".")[0]) for k in sd if k.startswith("lin"))
    pts = pts.to(dtype)
    inp = torch.cat([latent.to(dtype).reshape(1, -1).expand(pts.shape[0], -1), pts], 1)
    x = inp
    for l in range(n_layers):
        if "lin%d.weight_v" % l in sd:
            v = sd["lin%d.weight_v" % l].to(dtype)
            w = v * (sd["lin%d.weight_g" % l].to(dtype) / v.norm(dim=1, keepdim=True))
        else:
            w = sd["lin%d.weight" % l].to(dtype)
        b = sd["lin%d.bias" % l].to(dtype)
        if l == latent_in:
            x = torch.cat([x, inp], 1)
        x = x @ w.t() + b
        if l < n_layers - 1:
            x = torch.relu(x)
    return torch.tanh(x)


def select_partial_samples(pts, values, mask, num_samples=None, uniform_ratio=0.5):
    """core/data.py:76-119: ratio selection without the shuffle, keep the rows ``mask`` marks (np.intersect1d
    returns them in ascending row order), shuffle, then the fair windows of select_samples."""
    ids = np.arange(pts.shape[0], dtype=np.float64).reshape(-1, 1)
    pts_i, values = select_uniform_ratio_samples(np.hstack((ids, pts)), values, uniform_ratio, randomize=False)
    keep = np.nonzero(np.isin(pts_i[:, 0], np.where(mask)[0]))[0]
    keep = keep[np.argsort(pts_i[keep, 0], kind="stable")]
    pts, values = pts_i[keep, 1:], values[keep, :]
    perm = np.random.permutation(pts.shape[0])
    return select_samples(pts[perm], values[perm], num_samples, uniform_ratio=None)


def deepsdf_loss_terms(sd, latent, xyz, sdf_gt, clamp_dist=0.1, code_reg_lambda=1e-4, l2reg=True, dtype=torch.float32):
    """reconstruct_deepsdf.py:112-134: L1 between the clamped prediction and the clamped target (+ regulariser).
    Returns (loss, data_loss, reg_loss)."""
    pred = torch.clamp(deepsdf_forward(sd, latent, xyz, dtype), -clamp_dist, clamp_dist)
    gt = torch.clamp(sdf_gt.to(dtype).reshape(-1, 1), -clamp_dist, clamp_dist)
    data = torch.nn.functional.l1_loss(pred, gt)
    reg = code_reg_lambda * torch.mean(latent.pow(2)) if l2reg else torch.zeros((), dtype=dtype)
    return data + reg, data, reg


def deepsdf_latent_grad(sd, latent, xyz, sdf_gt, **kw):
    latent = latent.detach().clone().reshape(1, -1).requires_grad_(True)
    terms = deepsdf_loss_terms(sd, latent, xyz, sdf_gt, **kw)
    terms[0].backward()
    return latent.grad.detach().reshape(-1), [float(t.detach()) for t in terms]


def deepsdf_adam_loop(sd, latent0, batches, num_iterations, lr=5e-3, betas=(0.9, 0.999), eps=1e-8, log_every=10, **kw):
    """reconstruct_deepsdf.py:46,98-152 with an injected sample schedule ``batches(e) -> (xyz, sdf)``."""
    code = latent0.detach().clone().reshape(-1)
    m = torch.zeros_like(code)
    v = torch.zeros_like(code)
    log = {k: [] for k in ("epoch", "data_loss", "reg_loss", "mag")}
    for e in range(num_iterations):
        xyz, sdf = batches(e)
        g, t = deepsdf_latent_grad(sd, code, xyz, sdf, **kw)
        step = e + 1
        m = betas[0] * m + (1 - betas[0]) * g
        v = betas[1] * v + (1 - betas[1]) * g * g
        step_size = lr_at(e, lr, num_iterations) / (1 - betas[0] ** step)
        code = code - step_size * m / (v.sqrt() / math.sqrt(1 - betas[1] ** step) + eps)
        if e % log_every == 0:  # the reference logs AFTER optimizer.step(): the norm is the updated latent's (:136-141)
            log["epoch"].append(e)
            log["data_loss"].append(t[1])
            log["reg_loss"].append(t[2])
            log["mag"].append(float(code.norm()))
    return log, code.reshape(1, -1)
