"""Where reconstruct() spends its time (one shape, 800 x 30,000, 500,000-row sample set): the reference-stream sampler
path against the device-sampler path, plus the pieces of the prologue."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import synth_data
from bench import make_decoder
from deepjoin_b200 import reconstruct as RR


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dec = make_decoder(0, dev, "fp16x2")
    xyz, sdf, nrm = synth_data.make_sample_set(500000, seed=0)
    args = (dec, 800, 128, 64, (xyz, sdf, nrm), 0.01, 1.0, 0.0, 1.0, 0.1)
    kw = dict(num_samples=30000, lr=5e-3, l2reg=True)
    RR.reconstruct(dec, 20, 128, 64, (xyz, sdf, nrm), 0.01, 1.0, 0.0, 1.0, 0.1, **kw)
    for method in (None, "device", None, "device"):
        np.random.seed(0); torch.manual_seed(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        RR.reconstruct(*args, sampling_method=method, **kw)
        torch.cuda.synchronize()
        print("reconstruct(sampling_method=%r): %.3f s" % (method, time.perf_counter() - t0), flush=True)
    vals = np.ascontiguousarray(np.asarray(sdf).reshape(-1), dtype=np.float32)
    t0 = time.perf_counter()
    n = 0
    for blk in RR._schedule_blocks(vals, 800, 30000, 0.2):
        n += blk.shape[0]
    print("schedule blocks alone (producer thread, no device work): %.3f s for %d iterations" % (time.perf_counter() - t0, n))
    t0 = time.perf_counter()
    RR.make_schedule(vals, 800, 30000, 0.2)
    print("make_schedule, 800 iterations in one call: %.3f s" % (time.perf_counter() - t0))
    t0 = time.perf_counter()
    a = torch.from_numpy(np.ascontiguousarray(xyz, dtype=np.float32)).to(dev)
    b = torch.from_numpy(np.ascontiguousarray(np.asarray(sdf).reshape(-1), dtype=np.float32)).to(dev)
    c = torch.from_numpy(np.ascontiguousarray(nrm, dtype=np.float32)).to(dev)
    torch.cuda.synchronize()
    print("pool conversion + upload: %.1f ms (dtypes %s %s %s)" % (1e3 * (time.perf_counter() - t0), xyz.dtype, sdf.dtype, nrm.dtype))
    t0 = time.perf_counter()
    RR.make_schedule(sdf, 10, 30000, 0.2)
    print("first 10 iterations of the host sampler: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
    t0 = time.perf_counter()
    RR.make_schedule(sdf, 100, 30000, 0.2)
    print("100 iterations of the host sampler: %.1f ms" % (1e3 * (time.perf_counter() - t0)))




def trace():
    """Per-block timeline of the pipelined reference-sampler path: when the block left the producer, when the main
    thread had it enqueued, when the device finished it."""
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dec = make_decoder(0, dev, "fp16x2")
    xyz, sdf, nrm = synth_data.make_sample_set(500000, seed=0)
    kw = dict(num_samples=30000, lr=5e-3, l2reg=True)
    RR.reconstruct(dec, 20, 128, 64, (xyz, sdf, nrm), 0.01, 1.0, 0.0, 1.0, 0.1, **kw)
    orig = RR._schedule_blocks
    marks = []

    def wrapped(*a, **k):
        for blk in orig(*a, **k):
            t_got = time.perf_counter()
            ev = torch.cuda.Event(enable_timing=True)
            yield blk
            # control returns here when the consumer asks for the next block = after this block was enqueued
            ev.record()
            marks.append((blk.shape[0], t_got, time.perf_counter(), ev))

    RR._schedule_blocks = wrapped
    np.random.seed(0); torch.manual_seed(0)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e0.record()
    t0 = time.perf_counter()
    RR.reconstruct(dec, 800, 128, 64, (xyz, sdf, nrm), 0.01, 1.0, 0.0, 1.0, 0.1, **kw)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("total %.3f s" % (t1 - t0))
    for n, tg, te, ev in marks:
        print("block of %3d: produced at %7.1f ms, enqueued at %7.1f ms, device done at %7.1f ms" % (n, 1e3 * (tg - t0), 1e3 * (te - t0), e0.elapsed_time(ev)))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "trace":
    trace()
    sys.exit(0)

if __name__ == "__main__":
    main()
