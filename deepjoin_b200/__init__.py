"""deepjoin_b200 -- B200 (sm_100a) implementation of the DeepJoin inference hot path:
decoder grid queries, latent-code optimisation and marching cubes, behind the reference's own
Python call surface (``networks.<arch>.Decoder``, ``core.reconstruct.*``, ``reconstruct()``)
and a C ABI (include/deepjoin_b200.h)."""
__version__ = "0.1.0"
