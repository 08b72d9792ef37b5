"""Exception types of the reference's core/errors.py (same names, same base class)."""


class NoSamplesError(RuntimeError):
    """Raised when the data cannot be sampled correctly (core/errors.py:1)."""


class IsosurfaceExtractionError(RuntimeError):
    """Raised when the network cannot extract an isosurface (core/errors.py:7)."""


class DecoderNotLoadedError(RuntimeError):
    """Raised when the user tries to generate a mesh without a decoder (core/errors.py:13)."""


class PathAccessError(RuntimeError):
    """Raised when the user accesses a path that they shouldn't (core/errors.py:19)."""


class MeshEmptyError(RuntimeError):
    """Raised when a mesh is empty (core/errors.py:25)."""


class MeshContainsError(RuntimeError):
    """Raised when check_contains fails (core/errors.py:31)."""
