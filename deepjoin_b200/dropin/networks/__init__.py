"""``__import__("networks." + specs["NetworkArch"], fromlist=["Decoder"])`` shim
(reference reconstruct.py:434): resolves to deepjoin_b200.networks."""
import importlib
import sys

_real = importlib.import_module("deepjoin_b200.networks")
for _sub in ("decoder_z_lb_both_leaky_n_sdf", "decoder_deepsdf"):
    sys.modules[__name__ + "." + _sub] = importlib.import_module("deepjoin_b200.networks." + _sub)
sys.modules[__name__] = _real
