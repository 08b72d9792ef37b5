"""Drop-in for the reference's ``core/reconstruct.py`` (same function names / arguments /
return types); grid construction, batched decoder queries and marching cubes run on the GPU
behind the C ABI.  Line numbers cite the reference file."""
import ctypes
import logging
import threading

import numpy as np
import torch

from deepjoin_b200 import _lib
from deepjoin_b200.mesh import Mesh
from . import errors


def get_query_points(dims=(256, 256, 256), padding=0.1):
    """Return a grid of query points (core/reconstruct.py:25-30; np.meshgrid 'xy' order).
    Host-side utility only: the query path generates these in-kernel."""
    grid_pts = np.meshgrid(*[np.linspace(0, 1.0 + (padding * 2), d) - (0.5 + padding) for d in dims])
    return np.vstack([p.flatten() for p in grid_pts]).T


def tensor_sdf_to_occ(data):
    """core/reconstruct.py:33-34."""
    return torch.clamp(torch.sgn(-data), -1, 0) + 1


def tensor_sdf_to_udf(data):
    return torch.abs(data)


def tensor_udf_to_sdf(udf, occ):
    return udf * (((occ.round() * 2) - 1) * -1)


def _device_of(decoder):
    dev = next(decoder.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("deepjoin_b200: the decoder must live on a CUDA device (decoder.cuda())")
    return dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())


def _code_host(vec):
    if isinstance(vec, torch.Tensor):
        vec = vec.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(vec, dtype=np.float32).reshape(-1))


def decode_samples(vec, decoder, pts, use_net=1, batch_size=2 ** 16, sigmoid=True):
    """core/reconstruct.py:89-132.  ``pts`` [P,3] numpy -> values [P,1] float64.
    ``batch_size`` is accepted for signature compatibility; chunking is internal."""
    decoder._check_eval()
    if use_net not in (0, 1, 2, 3):
        raise RuntimeError("Requested output from non-existent network: {}".format(use_net))
    dev = _device_of(decoder)
    pack = decoder.pack(dev)
    pts = np.asarray(pts, dtype=np.float64)
    if pts.ndim != 2 or pts.shape[1] != 3:
        raise ValueError("pts must be [P, 3]")
    n = pts.shape[0]
    strides = (pts.strides[0] // 8, pts.strides[1] // 8)
    # rows, or the three coordinate arrays behind get_query_points' np.vstack(...).T: passed as they lie
    if not (strides == (3, 1) or (strides[0] == 1 and strides[1] >= n and n > 1)):
        pts = np.ascontiguousarray(pts)
        strides = (3, 1)
    # the result lands directly in an array from torch's pinned pool (no staging copy, no page faults)
    out = host_empty((pts.shape[0], 1), torch.float64).numpy()
    code = _code_host(vec)
    _lib.check(_lib.lib().dj_decode_samples_host_strided(
        pack.handle, code.ctypes.data, pts.ctypes.data, n, strides[0], strides[1], int(use_net),
        int(bool(decoder._return_occ)), int(bool(sigmoid)), decoder._prec(), out.ctypes.data))
    return out


def decode_shape_device(vec, decoder, dims, use_net=1, padding=0.0, sigmoid=True, rows=None):
    """Grid query that stays on the device: f32 CUDA tensor of the flat rows
    [r_lo, r_hi) of decode_shape(..., reshape=False)."""
    decoder._check_eval()
    if use_net not in (0, 1, 2, 3):
        raise RuntimeError("Requested output from non-existent network: {}".format(use_net))
    dev = _device_of(decoder)
    pack = decoder.pack(dev)
    d0, d1, d2 = [int(d) for d in dims]
    total = d0 * d1 * d2
    r_lo, r_hi = (0, total) if rows is None else rows
    if isinstance(vec, torch.Tensor) and vec.device == dev:
        code = vec.detach().reshape(-1).to(torch.float32).contiguous()  # stays on the device: no sync
    else:
        code = torch.from_numpy(_code_host(vec)).to(dev)
    out = torch.empty(r_hi - r_lo, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dj_query_grid(
            pack.handle, code.data_ptr(), d0, d1, d2, 1.0 + (padding * 2), 0.5 + padding, r_lo, r_hi, int(use_net),
            int(bool(decoder._return_occ)), int(bool(sigmoid)), decoder._prec(), out.data_ptr(), _lib.stream_ptr()))
    return out


def decode_shape(vec, decoder, dims, use_net=1, batch_size=2 ** 16, padding=0.0, reshape=True, sigmoid=True):
    """core/reconstruct.py:45-86: values over the np.meshgrid('xy') grid, float64."""
    # float64 (what the reference returns) is made on the device and copied into pinned host memory
    values = to_host(decode_shape_device(vec, decoder, dims, use_net, padding, sigmoid).to(torch.float64)).reshape(-1, 1)
    if reshape:
        return values.reshape([d for d in dims]).T
    return values


_PINNED_MESH_LIMIT = 1 << 28  # larger meshes go to pageable memory
_tls = threading.local()


def _work_for(device_index, stream=None):
    """Marching-cubes / meshing workspace (DjMcWork) of the calling THREAD for this device and stream: a workspace
    is one call at a time, so nothing is shared between threads or between streams (no process-global state;
    SURVEY 8b).  ``stream=None``: the host-facing create_mesh path, which runs on the workspace's own stream."""
    cache = getattr(_tls, "work", None)
    if cache is None:
        cache = _tls.work = {}
    key = (device_index, stream)
    w = cache.get(key)
    if w is None:
        w = cache[key] = _lib.McWork(device_index)
    return w


def marching_cubes(volume, level, gradient_direction="descent"):
    """skimage.measure.marching_cubes(volume, level, gradient_direction=...)[:2] as called at
    core/reconstruct.py:169-174 (spacing 1): CUDA volume [n0,n1,n2] f32 ->
    (verts f32 [V,3] in index units, faces int32 [F,3]) CUDA tensors."""
    if gradient_direction not in ("descent", "ascent"):
        raise ValueError("Incorrect input %s in `gradient_direction`" % gradient_direction)
    if volume.dim() != 3:
        raise ValueError("Input volume should be a 3D numpy array.")
    vol = volume.detach().to(torch.float32).contiguous()
    dev = vol.device
    nv, nf = ctypes.c_int64(), ctypes.c_int64()
    with torch.cuda.device(dev):
        st = _lib.stream_ptr()
        w = _work_for(dev.index, st)
        _lib.check(_lib.lib().dj_mc_count(w.handle, vol.data_ptr(), vol.shape[0], vol.shape[1], vol.shape[2], 0, 0,
                                          float(level), 1, ctypes.byref(nv), ctypes.byref(nf), st))
        verts = torch.empty((nv.value, 3), device=dev, dtype=torch.float32)
        faces = torch.empty((nf.value, 3), device=dev, dtype=torch.int32)
        _lib.check(_lib.lib().dj_mc_fill(w.handle, int(gradient_direction == "descent"), verts.data_ptr(),
                                         faces.data_ptr(), st))
    return verts, faces


_pool_warm = False


def _warm_pinned_pool():
    """torch's pinned allocator hands out power-of-two blocks and keeps freed ones; a first request of a size class
    costs a cudaHostAlloc (milliseconds, and it synchronises).  Touch every class up to 64 MB once (127 MB, ~40 ms)
    so that later results of any size find a block."""
    global _pool_warm
    if _pool_warm:
        return
    _pool_warm = True
    try:
        blocks = [torch.empty(1 << k, dtype=torch.uint8, pin_memory=True) for k in range(20, 27)]
        del blocks
    except RuntimeError:
        pass


def host_empty(shape, dtype):
    """Uninitialised host tensor for a result that a device->host copy fills: from torch's pinned-memory pool when
    it is small enough (link-rate DMA, no page faults on first touch, block recycled when the result is dropped),
    pageable otherwise or when the pool cannot grow."""
    n = 1
    for d in shape:
        n *= int(d)
    nbytes = n * torch.empty((), dtype=dtype).element_size()
    if 0 < nbytes <= _PINNED_MESH_LIMIT:
        _warm_pinned_pool()
        try:
            return torch.empty(shape, dtype=dtype, pin_memory=True)
        except RuntimeError:  # cudaHostAlloc failed: locked-memory limit of the host
            pass
    return torch.empty(shape, dtype=dtype)


def to_host(t):
    """CUDA tensor -> numpy array backed by torch's pinned-memory pool (link-rate copy, no page faults on a fresh
    allocation; the block is recycled when the array is dropped); larger results go to pageable memory."""
    t = t.detach().contiguous()
    out = host_empty(t.shape, t.dtype) if t.is_cuda else torch.empty(t.shape, dtype=t.dtype)
    out.copy_(t)
    return out.numpy()


def merge_vertices(vertices, faces):
    """Device-side ``Mesh.merge_vertices`` (trimesh's process=True merge, core/reconstruct.py:189):
    CUDA tensors vertices f64 [V,3] / faces int32 [F,3] -> (merged vertices, re-indexed faces)."""
    v = vertices.detach().to(torch.float64).contiguous().clone()
    f = faces.detach().to(torch.int32).contiguous().clone()
    dev = v.device
    n = ctypes.c_int64()
    with torch.cuda.device(dev):
        st = _lib.stream_ptr()
        _lib.check(_lib.lib().dj_mesh_merge_vertices(_work_for(dev.index, st).handle, v.data_ptr(), v.shape[0], f.data_ptr(),
                                                     f.shape[0], ctypes.byref(n), st))
    return v[:n.value], f


def create_mesh(vec, decoder, use_net, sigmoid=True, level=0.5, gradient_direction="descent", f_out=None,
                save_values=None, dims=(256, 256, 256), batch_size=2 ** 16, padding=0.1):
    """core/reconstruct.py:135-192: grid query -> marching cubes (spacing (1+2p)/d) -> swap
    vertex columns 0/1 -> centre -> mesh.  Raises IsosurfaceExtractionError like the reference."""
    decoder._check_eval()
    if use_net not in (0, 1, 2, 3):
        raise RuntimeError("Requested output from non-existent network: {}".format(use_net))
    if gradient_direction not in ("descent", "ascent"):
        raise errors.IsosurfaceExtractionError  # skimage raises ValueError, caught at :175
    dev = _device_of(decoder)
    pack = decoder.pack(dev)
    w = _work_for(dev.index)
    d0, d1, d2 = [int(d) for d in dims]
    code = _code_host(vec)
    nv, nf = ctypes.c_int64(), ctypes.c_int64()
    L = _lib.lib()
    if save_values is not None:
        logging.debug("Saving values to {}".format(save_values))
        np.save(save_values, decode_shape(vec, decoder, dims, use_net, batch_size, padding, False, sigmoid)
                .reshape([d for d in dims]))
    _lib.check(L.dj_create_mesh_count(pack.handle, w.handle, code.ctypes.data, d0, d1, d2, float(padding), int(use_net),
                                      int(bool(decoder._return_occ)), int(bool(sigmoid)), float(level),
                                      int(gradient_direction == "descent"), decoder._prec(),
                                      ctypes.byref(nv), ctypes.byref(nf)))
    # host arrays come from torch's pinned-memory pool (blocks are recycled when the mesh is dropped), so the
    # device->host copy runs at link rate and lands directly in the arrays the mesh keeps: float64
    # vertices, int64 faces as trimesh holds them (faces are widened on the device)
    vertices = host_empty((nv.value, 3), torch.float64).numpy()
    faces = host_empty((nf.value, 3), torch.int64).numpy()
    bad = ctypes.c_int(0)
    _lib.check(L.dj_create_mesh_fetch64(w.handle, vertices.ctypes.data, faces.ctypes.data, ctypes.byref(bad)))
    # trimesh's process=True (:189): the duplicate-vertex merge already ran on the device
    # (csrc/mesh_merge.cuh); non-finite vertices cannot come out of marching cubes on a finite grid,
    # the device-side check keeps the contract anyway
    mesh = Mesh(vertices=vertices, faces=faces, process=False)
    if bad.value:
        mesh.remove_infinite_values()
        mesh.merge_vertices()
    if f_out is not None:
        mesh.export(f_out)
    return mesh
