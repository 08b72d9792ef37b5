"""Synthetic inputs for the tests, oracle side.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Everything is defined in the oracle-free
``synth_data`` module at the repository root (so that bench.py's own arm never imports this package);
this module adds the one piece that needs the oracle: computing the SDF-head bias shifts of the
``trained_norm`` weights (committed as tests/golden/sdf_head_shifts.json for the seeds the tests and
the bench use, recomputed here for any other seed / architecture).
"""
import json
import os

import numpy as np
import torch

import synth_data as _S
from synth_data import (GOLDEN, MUGS, layer_shapes, make_code, make_deepsdf_state_dict, make_points,  # noqa: F401
                        make_sample_set, trained_code)


def compute_head_shifts(sd, a, seed):
    """(fnet shift, hnet shift): medians of the pre-tanh SDF (fnet) and of the hnet head pre-activation over a
    24^3 probe grid at the shipped latent row `seed`, evaluated with the fp64 oracle."""
    from . import decoder_oracle as O
    code = trained_code(seed) if (a["latent_size"], a["tool_latent_size"]) == (128, 64) else \
        make_code(seed, a["latent_size"], a["tool_latent_size"])
    lin = np.linspace(-0.55, 0.55, 24)
    pts = torch.tensor(np.stack(np.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(-1, 3), dtype=torch.float64)
    net = O.fold(sd, dtype=torch.float64, **a)
    yc, yt = O.raw_heads(net, code.double(), pts)
    s_f = float(yc[:, 0].median())
    sd = dict(sd)
    sd["lin%d.bias" % (a["f_layers"] - 1)] = sd["lin%d.bias" % (a["f_layers"] - 1)].clone()
    sd["lin%d.bias" % (a["f_layers"] - 1)][0] -= s_f
    net = O.fold(sd, dtype=torch.float64, **a)
    pre_t = O.raw_heads(net, code.double(), pts, h_pre=True)[1]
    return s_f, float(pre_t[:, 0].median())


def _center_with_oracle(sd, a, seed):
    s_f, s_h = compute_head_shifts(sd, a, seed)
    sd["lin%d.bias" % (a["f_layers"] - 1)][0] -= s_f
    sd["hnet_lin%d.bias" % (a["h_layers"] - 1)][0] -= s_h


def make_state_dict(seed=0, kind="default", prefix="", **arch):
    try:
        return _S.make_state_dict(seed, kind, prefix=prefix, **arch)
    except KeyError:
        pass
    # seed / architecture without a committed shift: centre with the oracle
    saved = _S._center_sdf_heads
    _S._center_sdf_heads = _center_with_oracle
    try:
        return _S.make_state_dict(seed, kind, prefix=prefix, **arch)
    finally:
        _S._center_sdf_heads = saved


def write_shift_fixture(seeds=range(8)):
    """Regenerate tests/golden/sdf_head_shifts.json."""
    out = {}
    saved = _S._center_sdf_heads
    for seed in seeds:
        box = {}

        def grab(sd, a, s, box=box):
            box["v"] = compute_head_shifts(sd, a, s)
            sd["lin%d.bias" % (a["f_layers"] - 1)][0] -= box["v"][0]
            sd["hnet_lin%d.bias" % (a["h_layers"] - 1)][0] -= box["v"][1]
        _S._center_sdf_heads = grab
        try:
            _S.make_state_dict(seed, "trained_norm")
        finally:
            _S._center_sdf_heads = saved
        out["%d" % seed] = list(box["v"])
    with open(os.path.join(GOLDEN, "sdf_head_shifts.json"), "w") as f:
        json.dump(out, f, indent=1)
    return out
