"""Exception types a caller of the reference catches by name (its ``core/errors.py:1-34``): the six
``RuntimeError`` subclasses below keep those names and that base class.  What is specific to this
library is where they come from: the C ABI never throws, it returns a status code
(``include/deepjoin_b200.h``) and ``deepjoin_b200._lib.check`` maps the codes onto these types."""

# name -> (what raises it here, C status code that maps onto it or None)
_TYPES = {
    "NoSamplesError": ("the sampler finds a class (sdf > 0 / sdf <= 0) without rows -- core/data.py:56-59,66-69; "
                       "dj_sample_schedule", "DJ_ERR_NO_SAMPLES"),
    "IsosurfaceExtractionError": ("marching cubes: level outside the volume's range, no surface, or an invalid "
                                  "gradient_direction -- core/reconstruct.py:175-176; dj_mc_count / dj_create_mesh_count",
                                  "DJ_ERR_LEVEL_RANGE / DJ_ERR_NO_SURFACE"),
    "DecoderNotLoadedError": ("a mesh is requested from a handler without a decoder (kept for callers that catch it)", None),
    "PathAccessError": ("a handler is asked for a path outside its experiment (kept for callers that catch it)", None),
    "MeshEmptyError": ("a metric receives a mesh without vertices -- core/metrics.py:29,65", None),
    "MeshContainsError": ("the inside/outside test of the protrusion metric fails (kept for callers that catch it)", None),
}


def _make(name, doc, code):
    text = "Raised when " + doc + "." + ("  C status: " + code + "." if code else "")
    return type(name, (RuntimeError,), {"__doc__": text, "__module__": __name__})


for _name, (_doc, _code) in _TYPES.items():
    globals()[_name] = _make(_name, _doc, _code)

__all__ = sorted(_TYPES)
