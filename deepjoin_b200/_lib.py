"""ctypes binding of the C ABI (include/deepjoin_b200.h).  There is no CPU fallback:
if the shared library is missing or a call fails, this raises."""
import ctypes
import os

import numpy as np

from .core import errors

_HERE = os.path.dirname(os.path.abspath(__file__))
# DEEPJOIN_B200_LIB: load another build of the same ABI (the -DDJ_DEBUG diagnostics library of tools/)
LIB_PATH = os.environ.get("DEEPJOIN_B200_LIB") or os.path.join(_HERE, "lib", "libdeepjoin_b200.so")

(OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NO_SURFACE, ERR_LEVEL_RANGE, ERR_NO_SAMPLES, ERR_BAD_NET, ERR_DEVICE,
 ERR_PERMISSION, ERR_IO) = range(11)
PREC = {"fp32": 0, "bf16": 1, "fp16": 2, "fp16x2": 3}
ABI_VERSION = 4  # == DJ_ABI_VERSION in include/deepjoin_b200.h (checked by tests/test_host_cpu.py)

c_i, c_i64, c_d, c_f, c_p = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_float, ctypes.c_void_p


class NetDesc(ctypes.Structure):
    _fields_ = [("latent_size", ctypes.c_int32), ("tool_latent_size", ctypes.c_int32), ("hidden", ctypes.c_int32),
                ("f_layers", ctypes.c_int32), ("h_layers", ctypes.c_int32), ("latent_in", ctypes.c_int32),
                ("slope", ctypes.c_float)]


class LatentCfg(ctypes.Structure):
    _fields_ = [("num_iterations", ctypes.c_int32), ("num_samples", ctypes.c_int32), ("lr", c_f), ("lambda1", c_f),
                ("lambda3", c_f), ("clamp_dist", c_f), ("code_reg_lambda", c_f), ("l2reg", ctypes.c_int32),
                ("beta1", c_f), ("beta2", c_f), ("eps", c_f), ("log_every", ctypes.c_int32),
                ("precision", ctypes.c_int32), ("loss_kind", ctypes.c_int32), ("iter_offset", ctypes.c_int32),
                ("total_iterations", ctypes.c_int32)]


# name -> (restype, argtypes); every symbol include/deepjoin_b200.h declares
SIGNATURES = {
    "dj_abi_version": (c_i, []),
    "dj_last_error": (ctypes.c_char_p, []),
    "dj_pack_create": (c_i, [ctypes.POINTER(NetDesc), c_p, c_p, c_p, c_p, c_i, ctypes.POINTER(c_p)]),
    "dj_pack_destroy": (c_i, [c_p]),
    "dj_pack_device": (c_i, [c_p]),
    "dj_query_points": (c_i, [c_p, c_p, c_p, c_i64, c_i, c_i, c_i, c_i, c_p, c_p]),
    "dj_query_all": (c_i, [c_p, c_p, c_p, c_i64, c_i, c_p, c_p]),
    "dj_query_grid": (c_i, [c_p, c_p, c_i, c_i, c_i, c_d, c_d, c_i64, c_i64, c_i, c_i, c_i, c_i, c_p, c_p]),
    "dj_decode_samples_host": (c_i, [c_p, c_p, c_p, c_i64, c_i, c_i, c_i, c_i, c_p]),
    "dj_decode_samples_host_strided": (c_i, [c_p, c_p, c_p, c_i64, c_i64, c_i64, c_i, c_i, c_i, c_i, c_p]),
    "dj_latent_grad": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i64, ctypes.POINTER(LatentCfg), c_p, c_p, c_p]),
    "dj_latent_optimize": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_p, ctypes.POINTER(LatentCfg), c_p, c_p]),
    "dj_latent_optimize_batch": (c_i, [c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, ctypes.POINTER(LatentCfg), c_p, c_p]),
    "dj_sample_schedule": (c_i, [c_p, c_i64, c_i, c_i, c_f, ctypes.c_uint64, c_p, c_p]),
    "dj_host_mt19937_permutation": (c_i, [c_p, c_p, c_i64, c_p]),
    "dj_host_write_ply": (c_i, [ctypes.c_char_p, c_p, c_i64, c_p, c_i64]),
    "dj_host_write_obj": (c_i, [ctypes.c_char_p, c_p, c_i64, c_p, c_i64]),
    "dj_mc_create": (c_i, [c_i, ctypes.POINTER(c_p)]),
    "dj_mc_destroy": (c_i, [c_p]),
    "dj_mc_count": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_d, c_i, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), c_p]),
    "dj_mc_fill": (c_i, [c_p, c_i, c_p, c_p, c_p]),
    "dj_mc_top_plane": (c_i, [c_p, c_p, c_p]),
    "dj_create_mesh_count": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_d, c_i, c_i, c_i, c_d, c_i, c_i,
                                   ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]),
    "dj_create_mesh_fetch": (c_i, [c_p, c_p, c_p]),
    "dj_create_mesh_fetch64": (c_i, [c_p, c_p, c_p, c_p]),
    "dj_mesh_merge_vertices": (c_i, [c_p, c_p, c_i64, c_p, c_i64, ctypes.POINTER(c_i64), c_p]),
    "dj_nn_query": (c_i, [c_p, c_i64, c_p, c_i64, c_p, c_p, c_p]),
    "dj_launch_count": (c_i64, [c_i]),
    "dj_host_sample_schedule": (c_i, [c_p, c_p, c_p, c_i64, ctypes.c_double, c_i, c_i, c_i, c_p, c_p]),
}

# diagnostics that only a -DDJ_DEBUG build exports (python -m deepjoin_b200.build --debug; DEEPJOIN_B200_LIB selects it)
DEBUG_SIGNATURES = {
    "dj_debug_umma_probe": (c_i, [c_p, c_p, c_i, c_i, c_i, ctypes.c_uint32, ctypes.c_uint64, c_p]),
    "dj_debug_read_prof": (c_i, [c_p, c_p, c_i]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "deepjoin_b200: %s is missing. Build it with `python -m deepjoin_b200.build` "
                "(needs nvcc; sm_100a). There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        for name, (res, args) in DEBUG_SIGNATURES.items():
            if hasattr(L, name):
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
        got = L.dj_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError("deepjoin_b200: ABI version mismatch: library %d, bindings %d (rebuild: python -m deepjoin_b200.build --force)" % (got, ABI_VERSION))
        _lib = L
    return _lib


def last_error():
    return lib().dj_last_error().decode("utf-8", "replace")


def check(rc):
    """Map C status codes to the reference's exception types (core/errors.py, decoder :454-457)."""
    if rc == OK:
        return
    msg = last_error()
    if rc in (ERR_NO_SURFACE, ERR_LEVEL_RANGE):
        raise errors.IsosurfaceExtractionError(msg)
    if rc == ERR_NO_SAMPLES:
        raise errors.NoSamplesError(msg)
    if rc == ERR_BAD_NET:
        raise RuntimeError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_PERMISSION:
        raise PermissionError(msg)
    if rc == ERR_IO:
        raise OSError(msg)
    raise RuntimeError("deepjoin_b200 [%d]: %s" % (rc, msg))


def ptr(t):
    """Device / host pointer of a torch tensor or numpy array (must be contiguous)."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        assert t.flags["C_CONTIGUOUS"]
        return t.ctypes.data
    assert t.is_contiguous()
    return t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


class Pack:
    """Folded + packed decoder weights on one device (opaque DjPack*)."""

    def __init__(self, desc, f_w, f_b, h_w, h_b, device):
        self.handle = c_p()
        self.device = int(device)
        keep = [np.ascontiguousarray(a, np.float32) for a in list(f_w) + list(f_b) + list(h_w) + list(h_b)]
        nf, nh = len(f_w), len(h_w)

        def arr(xs):
            return (c_p * len(xs))(*[a.ctypes.data for a in xs])

        fw, fb = arr(keep[:nf]), arr(keep[nf:2 * nf])
        hw, hb = arr(keep[2 * nf:2 * nf + nh]), arr(keep[2 * nf + nh:])
        check(lib().dj_pack_create(ctypes.byref(desc), fw, fb, hw, hb, self.device, ctypes.byref(self.handle)))

    def close(self):
        if self.handle:
            lib().dj_pack_destroy(self.handle)
            self.handle = c_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class McWork:
    def __init__(self, device):
        self.handle = c_p()
        self.device = int(device)
        check(lib().dj_mc_create(self.device, ctypes.byref(self.handle)))

    def close(self):
        if self.handle:
            lib().dj_mc_destroy(self.handle)
            self.handle = c_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
