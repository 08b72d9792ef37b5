"""ctypes front end of oracle/mc_oracle.c + the create_mesh() post-processing
(core/reconstruct.py:168-187).  TEST INFRASTRUCTURE ONLY; parity vs scikit-image UNPINNED.
The sweep's case tables come from oracle/mc_case_tables.py (derived there, not the product's header)."""
import ctypes

import numpy as np

from . import build_oracle, mc_case_tables

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_oracle.build())
        _lib.dj_oracle_mc.restype = ctypes.c_int
        _lib.dj_oracle_mc.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                      ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                      ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
        _lib.dj_oracle_free.argtypes = [ctypes.c_void_p]
        _lib.dj_oracle_mc_variant.argtypes = [ctypes.c_void_p]
        _lib.dj_oracle_mc_variant.restype = ctypes.c_int
        t = mc_case_tables.build()
        arr = lambda x: (ctypes.c_int * len(x))(*x)
        _lib.dj_oracle_mc_set_tables.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_void_p]
        _lib.dj_oracle_mc_set_tables(arr(t["var_base"]), arr(t["n_amb"]), arr(t["amb_face"]), arr(t["n_tri"]), arr(t["tri"]),
                                     len(t["n_tri"]), arr(t["face_corners"]))
    return _lib


def marching_cubes(volume, level, gradient_direction="descent"):
    """Returns (verts float32 [V,3] in index units, columns (axis0,axis1,axis2); faces int32 [F,3]).
    ValueError / RuntimeError as skimage raises them."""
    vol = np.ascontiguousarray(volume, np.float32)
    if vol.ndim != 3:
        raise ValueError("Input volume should be a 3D numpy array.")
    if gradient_direction not in ("descent", "ascent"):
        raise ValueError("Incorrect input %s in `gradient_direction`" % gradient_direction)
    v, f = ctypes.c_void_p(), ctypes.c_void_p()
    nv, nf = ctypes.c_int64(), ctypes.c_int64()
    rc = lib().dj_oracle_mc(vol.ctypes.data, vol.shape[0], vol.shape[1], vol.shape[2], float(level),
                            int(gradient_direction == "descent"), ctypes.byref(v), ctypes.byref(f),
                            ctypes.byref(nv), ctypes.byref(nf))
    if rc == 5:
        raise ValueError("Surface level must be within volume data range.")
    if rc == 4:
        raise RuntimeError("No surface found at the given iso value.")
    if rc != 0:
        raise ValueError("bad volume")
    verts = np.ctypeslib.as_array(ctypes.cast(v, ctypes.POINTER(ctypes.c_float)), (nv.value, 3)).copy()
    faces = np.ctypeslib.as_array(ctypes.cast(f, ctypes.POINTER(ctypes.c_int32)), (nf.value, 3)).copy()
    lib().dj_oracle_free(v)
    lib().dj_oracle_free(f)
    return verts, faces


def mesh_from_volume(values, level, dims, padding, gradient_direction):
    """create_mesh() after decode_shape: MC with spacing (1+2p)/d, swap columns 0/1, centre
    (core/reconstruct.py:168-187).  float64 vertices, int32 faces."""
    verts, faces = marching_cubes(values, level, gradient_direction)
    spacing = np.r_[[(1 + (padding * 2)) / d for d in dims]]
    vertices = verts * spacing
    vertices = np.hstack([np.expand_dims(v, axis=1) for v in (vertices[:, 1], vertices[:, 0], vertices[:, 2])])
    vertices -= 0.5 + padding
    return vertices, faces
