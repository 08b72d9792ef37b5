"""Execute the reference's own hot-path files in place (never copied).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Only usable where
``/root/reference`` exists (the build container); the GPU box has no copy, so
everything that must travel is dumped into ``tests/golden`` by
``tests/golden/make_golden.py``.

The shims below are the ones listed in SURVEY.md Appendix C: the reference
hard-codes ``.cuda()`` (networks/decoder_z_lb_both_leaky_n_sdf.py:169-171,
reconstruct.py:49-53,122,132; core/reconstruct.py:122), uses the removed
``np.float`` alias (core/reconstruct.py:130) and imports trimesh / skimage /
pickle5 at module top without needing them for the functions used here.
"""
import importlib
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("DEEPJOIN_REFERENCE", "/root/reference")
REF_PY = os.path.join(REF_ROOT, "deepjoin", "python")
MUGS = os.path.join(REF_ROOT, "deepjoin", "experiments", "mugs")


def available():
    return os.path.isfile(os.path.join(REF_PY, "reconstruct.py"))


_loaded = None


def load(cpu_shims=True):
    """Returns a namespace with .Decoder, .core_reconstruct, .core_data,
    .reconstruct (the function) and .specs, all being the reference's objects."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if REF_PY not in sys.path:
        sys.path.insert(0, REF_PY)
    if not hasattr(np, "float"):
        np.float = float
    if cpu_shims and not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    for m in ("trimesh", "skimage", "skimage.measure", "pickle5"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    saved_core = sys.modules.get("core")
    core = types.ModuleType("core")
    core.__path__ = [os.path.join(REF_PY, "core")]
    sys.modules["core"] = core

    def _load(name):
        spec = importlib.util.spec_from_file_location(
            name, os.path.join(REF_PY, name.replace(".", "/") + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    core.errors = _load("core.errors")
    core.reconstruct = _load("core.reconstruct")
    core.data = _load("core.data")
    rec = _load("reconstruct")
    arch = importlib.import_module("networks.decoder_z_lb_both_leaky_n_sdf")
    with open(os.path.join(MUGS, "specs.json")) as f:
        specs = json.load(f)
    ns = types.SimpleNamespace(
        Decoder=arch.Decoder, core_reconstruct=core.reconstruct,
        core_data=core.data, core_errors=core.errors,
        reconstruct=rec.reconstruct, specs=specs, core=core)
    # leave sys.modules["core"] pointing at the reference: its functions
    # resolve `core.reconstruct...` through the module global at call time.
    _loaded = ns
    return ns


def make_decoder(ns=None, state_dict=None):
    ns = ns or load()
    s = ns.specs
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dec = ns.Decoder(latent_size=s["CodeLength"], tool_latent_size=s["BreakCodeLength"],
                         num_dims=3, do_code_regularization=True,
                         **s["NetworkSpecs"], **s["SubnetSpecs"])
    if state_dict is not None:
        dec.load_state_dict(state_dict)
    return dec.eval()
