"""``import core`` shim: put <repo> and <repo>/deepjoin_b200/dropin on PYTHONPATH in place of the
reference's deepjoin/python and ``core`` / ``core.reconstruct`` / ``core.errors`` / ``core.data`` /
``core.utils_multiprocessing`` resolve to deepjoin_b200.core (one module identity, so exception
classes match across both import names)."""
import importlib
import sys

_real = importlib.import_module("deepjoin_b200.core")
for _sub in ("errors", "reconstruct", "data", "utils_multiprocessing", "utils_deepsdf", "metrics", "workspace", "handler"):
    sys.modules[__name__ + "." + _sub] = importlib.import_module("deepjoin_b200.core." + _sub)
sys.modules[__name__] = _real
