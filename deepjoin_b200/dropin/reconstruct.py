"""``from reconstruct import reconstruct`` shim (reference reconstruct.py:19)."""
from deepjoin_b200.reconstruct import reconstruct  # noqa: F401
