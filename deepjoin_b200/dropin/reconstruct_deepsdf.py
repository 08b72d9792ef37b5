"""``from reconstruct_deepsdf import reconstruct`` shim (reference reconstruct_deepsdf.py:17)."""
from deepjoin_b200.reconstruct_deepsdf import reconstruct  # noqa: F401
