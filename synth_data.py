"""Deterministic synthetic weights, latent codes and sample sets (bench / test INPUT data).

No dependency on the oracle: bench.py's own arm, smoke() and the tests all draw their inputs from here.

The reference ships no trained weights (``.MISSING_LARGE_BLOBS``), so both the
oracle and the CUDA path are fed the same synthetic state dicts:

* ``default``      : PyTorch-default-style init (U(-1/sqrt(in), 1/sqrt(in))),
                     weight_norm g = ||v||_row.  Degenerate field (no iso-surface).
* ``trained_norm`` : the same tensors rescaled so every parameter has the final
                     Frobenius norm recorded in the reference's
                     ``experiments/mugs/Logs.pth["param_magnitude"]`` (committed
                     as ``tests/golden/param_magnitude.json``), then the SDF
                     head biases are shifted so the median pre-tanh SDF over a
                     probe grid is 0 (guarantees an iso-surface).

State-dict keys are exactly the reference's
(networks/decoder_z_lb_both_leaky_n_sdf.py:82-87,117-122).
"""
import json
import math
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden")

MUGS = dict(latent_size=128, tool_latent_size=64, hidden=512, f_layers=9, h_layers=6, latent_in=4)


def layer_shapes(latent_size=128, tool_latent_size=64, hidden=512, f_layers=9, h_layers=6, latent_in=4):
    """[(name, out, in, weight_normed)] as Decoder.__init__ builds them (:60-131)."""
    d0 = latent_size + 3
    out = []
    for l in range(h_layers):
        i = tool_latent_size + 3 if l == 0 else hidden
        o = 5 if l == h_layers - 1 else hidden
        out.append(("hnet_lin%d" % l, o, i, l < h_layers - 1))
    for l in range(f_layers):
        i = d0 if l == 0 else hidden
        if l == f_layers - 1:
            o = 5
        elif l + 1 == latent_in:
            o = hidden - d0
        else:
            o = hidden
        out.append(("lin%d" % l, o, i, l < f_layers - 1))
    return out


def make_state_dict(seed=0, kind="default", prefix="", **arch):
    a = dict(MUGS)
    a.update(arch)
    g = torch.Generator().manual_seed(1000 + seed)
    sd = {}
    for name, o, i, wn in layer_shapes(**a):
        bound = 1.0 / math.sqrt(i)
        w = (torch.rand(o, i, generator=g) * 2 - 1) * bound
        b = (torch.rand(o, generator=g) * 2 - 1) * bound
        if wn:
            sd[name + ".weight_g"] = w.norm(dim=1, keepdim=True).clone()
            sd[name + ".weight_v"] = w
        else:
            sd[name + ".weight"] = w
        sd[name + ".bias"] = b
    if kind == "trained_norm":
        with open(os.path.join(GOLDEN, "param_magnitude.json")) as f:
            mag = json.load(f)
        for k in sd:
            if k in mag:
                sd[k] = sd[k] * (mag[k] / float(sd[k].norm()))
        _center_sdf_heads(sd, a, seed)
    elif kind != "default":
        raise ValueError(kind)
    if prefix:
        sd = {prefix + k: v for k, v in sd.items()}
    return sd


def _center_sdf_heads(sd, a, seed):
    """Shift the two SDF-head biases so that the median pre-tanh SDF over a 24^3 probe grid is 0 at the
    shipped latent row `seed`.  The shifts were computed once with the fp64 oracle
    (oracle/synth.py:compute_head_shifts) and are committed as tests/golden/sdf_head_shifts.json, so
    that generating the weights needs no oracle code."""
    key = "%d" % seed
    with open(os.path.join(GOLDEN, "sdf_head_shifts.json")) as f:
        shifts = json.load(f)
    if a != MUGS or key not in shifts:
        raise KeyError("no committed SDF-head shift for seed %s / this architecture: use oracle.synth.make_state_dict" % key)
    sd["lin%d.bias" % (a["f_layers"] - 1)][0] -= shifts[key][0]
    sd["hnet_lin%d.bias" % (a["h_layers"] - 1)][0] -= shifts[key][1]


def make_deepsdf_state_dict(seed=0, kind="default", prefix=""):
    """State dict of the reference's baseline ``networks.decoder_deepsdf.Decoder`` (mugs_deepsdf specs:
    latent 128, 8 x 512, latent_in [4], weight_norm): the fnet tensors of make_state_dict with the head cut
    to its SDF row (keys lin0..lin7 {weight_g, weight_v, bias}, lin8 {weight [1,512], bias [1]})."""
    sd = make_state_dict(seed, kind)
    out = {k: v.clone() for k, v in sd.items() if k.startswith("lin")}
    out["lin8.weight"] = out["lin8.weight"][:1].clone()
    out["lin8.bias"] = out["lin8.bias"][:1].clone()
    if prefix:
        out = {prefix + k: v for k, v in out.items()}
    return out


def trained_code(seed=0):
    """Row `seed` of the reference's shipped LatentCodes/2000.pth || 2000_tool.pth (first 8 rows are
    committed as tests/golden/latent_rows.npz) -- the code the trained-norm weights are centred on."""
    rows = np.load(os.path.join(GOLDEN, "latent_rows.npz"))
    return torch.from_numpy(np.concatenate([rows["c2000"][seed % 8], rows["c2000_tool"][seed % 8]]))


def make_code(seed=0, latent_size=128, tool_latent_size=64, std=0.0623):
    """Synthetic code with the per-element std of the shipped LatentCodes/2000.pth."""
    g = torch.Generator().manual_seed(7000 + seed)
    return torch.randn(latent_size + tool_latent_size, generator=g) * std


def make_points(n, seed=0, lo=-0.6, hi=0.6):
    g = torch.Generator().manual_seed(9000 + seed)
    return (torch.rand(n, 3, generator=g) * (hi - lo) + lo).float()


def make_sample_set(n=500000, seed=0):
    """Synthetic 'fractured mug' sample set in the reference's on-disk layout
    (fracturing/processor/process_sample.py:82-86,113-128,155; process_sdf2.py:129-133):
    first half near-surface (+N(0,0.01^2)), second half uniform in the 1.1-cube,
    xyz / sdf / n stored as float16.  Solid = hollow cylinder minus a half-space."""
    rng = np.random.default_rng(12345 + seed)
    half = n // 2

    def sdf_fn(p):
        r = np.sqrt(p[:, 0] ** 2 + p[:, 2] ** 2)
        outer = np.maximum(r - 0.30, np.abs(p[:, 1]) - 0.40)
        inner = np.maximum(r - 0.24, np.abs(p[:, 1] - 0.06) - 0.40)
        mug = np.maximum(outer, -inner)
        nrm = np.array([0.35, 0.8, 0.5]) / np.linalg.norm([0.35, 0.8, 0.5])
        cut = p @ nrm - (0.12 + 0.05 * ((seed % 5) - 2))
        return np.maximum(mug, cut)

    def normals(p, eps=1e-3):
        g = np.zeros_like(p)
        for a in range(3):
            d = np.zeros(3)
            d[a] = eps
            g[:, a] = (sdf_fn(p + d) - sdf_fn(p - d)) / (2 * eps)
        g /= np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-9)
        return g

    # surface half: rejection-sample points with |sdf| small, then jitter
    cand = rng.uniform(-0.5, 0.5, size=(half * 12, 3))
    s = sdf_fn(cand)
    keep = np.argsort(np.abs(s))[:half]
    surf = cand[rng.permutation(keep)] + rng.normal(0, 0.01, size=(half, 3))
    uni = rng.uniform(-0.55, 0.55, size=(n - half, 3))
    xyz = np.concatenate([surf, uni], 0).astype(np.float16)
    p = xyz.astype(np.float64)
    sdf = sdf_fn(p)[:, None].astype(np.float16)
    nrm = normals(p).astype(np.float16)
    return xyz, sdf, nrm
