"""Drop-in for the function ``reconstruct()`` of the reference's ``reconstruct.py:19-226``
(duplicated in ``reconstruct_single.py:19-226``): per-shape Adam optimisation of
[latent (L), break_latent (Lb)] through the frozen decoder.

Same arguments, same return value ``(loss_dict, code[1, L+Lb])``, same initialisation
(``torch.ones(1,L).normal_(0, stat)`` on the CPU generator, :48-50) and the same mini-batch
stream (numpy global RNG consumed exactly like core/data.py does), but the 800-iteration loop
-- batch gather, forward, clamped-L1 + BCE + normal loss, backward to the code only, Adam, lr
step, loss logging every 10th iteration -- runs on the device without host round trips
(dj_latent_optimize).  The reference's sampler (62 ms per iteration in numpy) is ONE native call on numpy's own
MT19937 stream (dj_host_sample_schedule, ~1 ms per iteration, bit-identical rows) on a background thread; the rows of
its finished leading iterations go to the device in 10-iteration blocks while the rest is still being drawn.  ``sampling_method="device"`` (extension) draws
the schedule on the device instead (dj_sample_schedule; same distribution, not numpy's stream).

Arithmetic follows ``decoder.precision`` (override: ``decoder.latent_precision``):
  "fp16x2" (default)  split-operand tensor-core step, forward and backward (csrc/latent_tc2.cuh, X2): loss terms within
                      1e-5 and gradient within 1e-4 of the fp32 / autograd ones -- reproduces the reference's own seeded
                      run to the tolerance the fp32 path does (tests/test_gpu_latent.py)
  "fp32"              the reference's arithmetic on CUDA cores (13x slower)
  "bf16" / "fp16"     single-pass tensor-core step, 3.8x faster than fp16x2; gradient with ~1e-2 / 3e-3 relative
                      rounding noise, loss gates decided on a 16-bit forward: opt-in throughput modes"""
import os
from collections import defaultdict

import numpy as np
import torch

from deepjoin_b200 import _lib
from deepjoin_b200.core import data as _data

LOG_KEYS = ("epoch", "loss", "sdf_data_loss", "occ_data_loss", "norm_loss", "reg_loss", "mag", "h_mag")


def make_schedule(b_sdf_gt, num_iterations, num_samples, uniform_ratio=0.2, n_threads=0, out=None, progress=None):
    """Row indices of every iteration's mini-batch, consuming numpy's global RNG exactly as the
    reference loop does (reconstruct.py:114-118 -> core/data.py:15-73).  ``n_threads``: worker threads of the native
    sampler (0 = one per core but one).  ``out`` / ``progress``: a preallocated int32 [num_iterations, num_samples]
    array and a ctypes.c_int32 the native sampler keeps at the number of finished leading iterations (another
    thread may consume those rows while the call is still running)."""
    sched = np.empty((num_iterations, num_samples), dtype=np.int32) if out is None else out
    vals = np.ascontiguousarray(_data._f32(np.asarray(b_sdf_gt).reshape(-1)), dtype=np.float32)  # widened once
    state = np.random.get_state()
    if (state[0] == "MT19937" and uniform_ratio is not None and 0.0 < uniform_ratio < 1.0 and vals.shape[0] >= 2
            and num_iterations > 0 and os.environ.get("DEEPJOIN_B200_PY_SAMPLER") != "1"):
        # all iterations in one native call on numpy's own stream (dj_host_sample_schedule): the calling thread walks
        # the generator, worker threads finish the shuffles -- ~1 ms instead of ~20 ms per iteration, same rows
        import ctypes
        key = np.ascontiguousarray(state[1], dtype=np.uint32).copy()
        pos = ctypes.c_int32(int(state[2]))
        rc = _lib.lib().dj_host_sample_schedule(key.ctypes.data, ctypes.byref(pos), vals.ctypes.data, vals.shape[0],
                                                float(uniform_ratio), int(num_iterations), int(num_samples), int(n_threads),
                                                sched.ctypes.data, ctypes.byref(progress) if progress is not None else None)
        np.random.set_state((state[0], key, pos.value, state[3], state[4]))
        _lib.check(rc)
        return sched
    for e in range(num_iterations):
        sched[e] = _data.select_sample_indices(vals, num_samples, uniform_ratio=uniform_ratio)
        if progress is not None:
            progress.value = e + 1
    return sched


def make_schedule_surface_interior(b_sdf_gt, num_iterations, num_samples):
    """Row indices of every iteration's mini-batch for ``sampling_method="surface_interior_only"``
    (reconstruct.py:60-98): int(0.2 N) of the INTERIOR uniform rows (second half of the sample set, sdf < 0) drawn
    without replacement (with, if there are too few), then the fair positive / non-positive windows of
    select_samples over the surface rows (first half, no ratio step, no shuffle); numpy's global RNG consumed
    exactly as the reference does."""
    vals = _data._f32(np.asarray(b_sdf_gt).reshape(-1))
    half = int(vals.shape[0] / 2)
    uni_rows = half + np.where(vals[half:] < 0)[0]
    surf_vals = vals[:half]
    num_uni = int(num_samples * 0.2)
    num_surf = int(num_samples - num_uni)
    sched = np.empty((num_iterations, num_samples), dtype=np.int32)
    for e in range(num_iterations):
        if num_uni <= uni_rows.shape[0]:
            uni_inds = _data._permutation(uni_rows.shape[0])[:num_uni]  # == np.random.choice(n, num_uni, replace=False)
        else:
            uni_inds = np.random.choice(uni_rows.shape[0], num_uni, replace=True)
        sched[e, :num_uni] = uni_rows[uni_inds]
        sched[e, num_uni:] = _data._fair_windows(surf_vals, num_surf)
    return sched


def init_code(latent_size, break_latent_size, stat):
    """reconstruct.py:48-53 (CPU generator, so torch.manual_seed reproduces the reference)."""
    if type(stat) == type(0.1):
        latent = torch.ones(1, latent_size).normal_(mean=0, std=stat)
        break_latent = torch.ones(1, break_latent_size).normal_(mean=0, std=stat)
    else:
        latent = torch.normal(stat[0].detach(), stat[1].detach())
        break_latent = torch.normal(stat[0].detach(), stat[1].detach())
    return torch.cat([latent.reshape(1, -1), break_latent.reshape(1, -1)], dim=1).float()


def optimize_code(decoder, code0, pool_xyz, pool_sdf, pool_nrm, schedule, lr, lambda1, lambda3, clamp_dist,
                  l2reg=True, code_reg_lambda=1e-4, log_every=10, precision="fp32", loss_kind=0, schedule_blocks=None,
                  total_iterations=None):
    """Device loop.  pool_* CUDA f32 tensors ([M,3], [M], [M,3]); schedule CUDA int32
    [iterations, num_samples].  Returns (log [ceil(it/log_every), 8] CUDA, code [1, L+Lb] CUDA).

    ``schedule_blocks`` (instead of ``schedule``): an iterator of host int32 arrays [n_k, num_samples] whose lengths sum
    to ``total_iterations`` -- the loop runs block by block (DjLatentCfg.iter_offset), each block uploaded and enqueued
    as soon as it arrives, so that a host sampler producing the schedule overlaps with the device work."""
    dev = pool_xyz.device
    pack = decoder.pack(dev)
    if schedule_blocks is None:
        schedule_blocks = [schedule]
        total_iterations = int(schedule.shape[0])
    total = int(total_iterations)
    code = code0.detach().reshape(-1).to(device=dev, dtype=torch.float32).clone().contiguous()
    m = torch.zeros_like(code)
    v = torch.zeros_like(code)
    log = torch.zeros(((total + log_every - 1) // log_every, 8), device=dev, dtype=torch.float32)
    done, keep = 0, []
    side, stage, k = None, [None, None], 0
    for block in schedule_blocks:
        if not isinstance(block, torch.Tensor):
            block = torch.from_numpy(np.ascontiguousarray(block, dtype=np.int32))
        if block.device.type == "cpu":
            # Host blocks go through one of two pinned staging buffers on a side stream: the library enqueues a block
            # without waiting for the previous one (only the last block synchronises), so this copy overlaps the
            # running iterations.
            if side is None:
                side = torch.cuda.Stream(dev)
            buf, ev = stage[k & 1] or (None, None)
            if ev is not None:
                ev.synchronize()  # the copy that last read this staging buffer
            if buf is None or buf.numel() < block.numel():
                # sized for the largest block the producer hands out (100 iterations), so that it is allocated once
                # (torch's pinned pool then serves the same two blocks to every later call)
                buf = torch.empty(max(block.numel(), 100 * int(block.shape[1])), dtype=torch.int32, pin_memory=True)
            pinned = buf[:block.numel()].view(block.shape)
            pinned.copy_(block)
            with torch.cuda.stream(side):
                block = pinned.to(device=dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
            torch.cuda.current_stream(dev).wait_stream(side)
            stage[k & 1] = (buf, ev)
            k += 1
        else:
            block = block.to(device=dev, dtype=torch.int32).contiguous()
        keep.append(block)  # the enqueued iterations read it
        n_k, ns = int(block.shape[0]), int(block.shape[1])
        cfg = _lib.LatentCfg(n_k, ns, float(lr), float(lambda1), float(lambda3), float(clamp_dist),
                             float(code_reg_lambda), int(bool(l2reg)), 0.9, 0.999, 1e-8, int(log_every),
                             _lib.PREC[precision], int(loss_kind), done, total)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().dj_latent_optimize(
                pack.handle, code.data_ptr(), m.data_ptr(), v.data_ptr(), pool_xyz.data_ptr(), pool_sdf.data_ptr(),
                pool_nrm.data_ptr(), pool_xyz.shape[0], block.data_ptr(), _lib.ctypes.byref(cfg), log.data_ptr(),
                _lib.stream_ptr()))
        done += n_k
    if done != total:
        raise ValueError("schedule blocks cover %d of %d iterations" % (done, total))
    if len(keep) > 1:
        torch.cuda.current_stream(dev).synchronize()
    return log, code.reshape(1, -1)


def _schedule_blocks(b_sdf_gt, num_iterations, num_samples, uniform_ratio, block=100, first=10):
    """The reference sampler's schedule of a whole reconstruct() call, drawn by ONE make_schedule call on a background
    thread (numpy's global stream, in order; the native sampler releases the GIL) and handed out while it is being
    drawn: yields views of the leading finished iterations, `first` iterations as soon as they exist (the device starts
    after ~15 ms), then whatever has accumulated, at most `block`, in multiples of 10 (the log interval)."""
    import ctypes
    import threading
    import time
    sched = np.empty((num_iterations, num_samples), dtype=np.int32)
    progress = ctypes.c_int32(0)
    failure = []
    # the thread that walks numpy's generator is the sampler's critical path: leave two cores to the thread that
    # launches the device work and to the driver
    workers = max(1, (os.cpu_count() or 4) - 3)

    def produce():
        try:
            make_schedule(b_sdf_gt, num_iterations, num_samples, uniform_ratio, n_threads=workers, out=sched, progress=progress)
        except BaseException as exc:  # noqa: BLE001 -- re-raised in the consumer (NoSamplesError keeps its type)
            failure.append(exc)

    t = threading.Thread(target=produce, daemon=True)
    t.start()
    try:
        emitted, want = 0, max(1, min(first, block))
        while emitted < num_iterations:
            alive = t.is_alive()
            ready = min(int(progress.value), num_iterations)
            if failure:
                raise failure[0]
            if not alive and ready < num_iterations:
                ready = num_iterations  # a sampler that finished without moving the counter
            take = min(ready - emitted, block)
            if ready < num_iterations:
                take -= take % 10
            if take >= min(want, num_iterations - emitted) and take > 0:
                yield sched[emitted:emitted + take]
                emitted += take
                want = 10
            else:
                time.sleep(0.0005)
    finally:
        # also when the consumer gives up early: the sampler call runs to its end and restores numpy's generator state
        # before control returns to the caller, never behind its back
        t.join()
    if failure:
        raise failure[0]


def optimize_codes(decoder, codes0, pools, schedules, lr, lambda1, lambda3, clamp_dist, l2reg=True, code_reg_lambda=1e-4,
                   log_every=10, precision="bf16", loss_kind=0):
    """The device loop for S independent shapes at once (extension; the reference's worker runs them one after
    the other, core/utils_multiprocessing.py:33-71).  codes0 [S, L+Lb]; pools: S tuples (xyz [M,3], sdf [M],
    nrm [M,3]) of CUDA f32 tensors; schedules CUDA int32 [S, iterations, num_samples].
    Returns (logs [S, ceil(it/log_every), 8] CUDA, codes [S, L+Lb] CUDA).  With precision "bf16" every
    iteration is one fused tensor-core launch over all shapes (dj_latent_optimize_batch)."""
    import ctypes
    S = int(schedules.shape[0])
    if len(pools) != S or int(codes0.shape[0]) != S:
        raise ValueError("need one pool and one code per schedule")
    dev = schedules.device
    pack = decoder.pack(dev)
    iters, ns = int(schedules.shape[1]), int(schedules.shape[2])
    cfg = _lib.LatentCfg(iters, ns, float(lr), float(lambda1), float(lambda3), float(clamp_dist),
                         float(code_reg_lambda), int(bool(l2reg)), 0.9, 0.999, 1e-8, int(log_every),
                         _lib.PREC[precision], int(loss_kind))
    codes = codes0.detach().to(device=dev, dtype=torch.float32).reshape(S, -1).clone().contiguous()
    m = torch.zeros_like(codes)
    v = torch.zeros_like(codes)
    logs = torch.zeros((S, (iters + log_every - 1) // log_every, 8), device=dev, dtype=torch.float32)
    f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
    keep = [(f32(x), f32(s).reshape(-1), f32(n)) for x, s, n in pools]
    arr = lambda k: (ctypes.c_void_p * S)(*[t[k].data_ptr() for t in keep])
    pn = (ctypes.c_int64 * S)(*[int(t[0].shape[0]) for t in keep])
    sched = schedules.to(torch.int32).contiguous()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dj_latent_optimize_batch(
            pack.handle, S, codes.data_ptr(), m.data_ptr(), v.data_ptr(), arr(0), arr(1), arr(2), pn, sched.data_ptr(),
            ctypes.byref(cfg), logs.data_ptr(), _lib.stream_ptr()))
        torch.cuda.current_stream().synchronize()  # `keep` / the pointer arrays must outlive the enqueued loop
    return logs, codes


def latent_grad(decoder, code, xyz, sdf, nrm, lambda1=1.0, lambda3=1.0, clamp_dist=0.1, l2reg=True,
                code_reg_lambda=1e-4, precision="fp32", loss_kind=0):
    """One forward+backward: (dL/dcode [L+Lb], loss terms [5] = loss, sdf, occ, norm, reg)."""
    dev = xyz.device
    pack = decoder.pack(dev)
    cfg = _lib.LatentCfg(1, int(xyz.shape[0]), 0.0, float(lambda1), float(lambda3), float(clamp_dist),
                         float(code_reg_lambda), int(bool(l2reg)), 0.9, 0.999, 1e-8, 10, _lib.PREC[precision],
                         int(loss_kind))
    code = code.detach().reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
    grad = torch.empty_like(code)
    terms = torch.empty(5, device=dev, dtype=torch.float32)
    f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
    xyz, sdf, nrm = f32(xyz), f32(sdf).reshape(-1), f32(nrm)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dj_latent_grad(pack.handle, code.data_ptr(), xyz.data_ptr(), sdf.data_ptr(),
                                             nrm.data_ptr(), xyz.shape[0], _lib.ctypes.byref(cfg), grad.data_ptr(),
                                             terms.data_ptr(), _lib.stream_ptr()))
    return grad, terms


def reconstruct(decoder, num_iterations, latent_size, break_latent_size, test_sdf, stat, lambda1, lambda2, lambda3,
                clamp_dist, sampling_method=None, num_samples=30000, lr=5e-4, l2reg=False, code_reg_lambda=1e-4,
                iter_path=None, loss_version=None):
    decoder.eval()
    dev = next(decoder.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("deepjoin_b200: the decoder must live on a CUDA device")
    if latent_size != decoder.latent_size or break_latent_size != decoder.tool_latent_size:
        raise ValueError("latent sizes do not match the decoder")
    code0 = init_code(latent_size, break_latent_size, stat)

    if sampling_method == "surface_interior_only":
        # partial-view sampling (:60-98): (pts, sdf) only, no normals -- the reference's normal term would compare
        # [N,3] with [N,0] and raise, so it is only usable with lambda1 == 0
        pts_gt, b_sdf_gt = test_sdf
        if lambda1 != 0:
            raise RuntimeError("The size of tensor a (3) must match the size of tensor b (0) at non-singleton dimension 1"
                               " (sampling_method='surface_interior_only' carries no normals: use lambda1=0)")
        b_n_gt = np.zeros((np.asarray(pts_gt).shape[0], 3), dtype=np.float32)
    else:
        pts_gt, b_sdf_gt, b_n_gt = test_sdf
    pool_xyz = torch.from_numpy(np.ascontiguousarray(pts_gt, dtype=np.float32)).to(dev)
    pool_sdf = torch.from_numpy(np.ascontiguousarray(np.asarray(b_sdf_gt).reshape(-1), dtype=np.float32)).to(dev)
    pool_nrm = torch.from_numpy(np.ascontiguousarray(b_n_gt, dtype=np.float32)).to(dev)

    if sampling_method == "surface_interior_only":
        sched = torch.from_numpy(make_schedule_surface_interior(b_sdf_gt, num_iterations, num_samples)).to(dev)
    elif sampling_method == "device":
        sched = torch.empty((num_iterations, num_samples), device=dev, dtype=torch.int32)
        seed = int(np.random.randint(0, 2 ** 31 - 1))
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().dj_sample_schedule(pool_sdf.data_ptr(), pool_sdf.shape[0], int(num_iterations),
                                                     int(num_samples), 0.2, seed, sched.data_ptr(), _lib.stream_ptr()))
    else:
        sched = None  # the reference's sampler, drawn block by block while the device runs (same stream, same rows)

    prec = getattr(decoder, "latent_precision", None) or getattr(decoder, "precision", "fp16x2")
    if sched is None:
        vals = np.ascontiguousarray(_data._f32(np.asarray(b_sdf_gt).reshape(-1)), dtype=np.float32)
        log, code = optimize_code(decoder, code0, pool_xyz, pool_sdf, pool_nrm, None, lr, lambda1, lambda3, clamp_dist,
                                  l2reg=l2reg, code_reg_lambda=code_reg_lambda, log_every=10, precision=prec,
                                  schedule_blocks=_schedule_blocks(vals, num_iterations, num_samples, 0.2),
                                  total_iterations=num_iterations)
    else:
        log, code = optimize_code(decoder, code0, pool_xyz, pool_sdf, pool_nrm, sched, lr, lambda1, lambda3, clamp_dist,
                                  l2reg=l2reg, code_reg_lambda=code_reg_lambda, log_every=10, precision=prec)
    log = log.cpu().numpy()
    loss_dict = defaultdict(lambda: [])
    for row in log:
        loss_dict["epoch"].append(int(row[0]))
        for k, val in zip(LOG_KEYS[1:], row[1:]):
            loss_dict[k].append(float(val))
    return dict(loss_dict), code
