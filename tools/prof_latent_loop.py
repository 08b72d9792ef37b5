"""Kernel durations INSIDE the running latent loop (CUPTI through torch.profiler; kernels stay concurrent with the
host as in production, unlike an ncu launch list)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import synth_data
from bench import make_decoder
from deepjoin_b200 import reconstruct as RR
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
dec = make_decoder(0, dev, "bf16")
S, iters, ns = int(os.environ.get("SHAPES", "8")), 300, 30000
pools = []
for s in range(S):
    x, d, n = synth_data.make_sample_set(100000, seed=s)
    pools.append(tuple(torch.from_numpy(a.astype(np.float32)).to(dev) for a in (x, d.reshape(-1), n)))
sched = torch.randint(0, 100000, (S, iters, ns), device=dev, dtype=torch.int32)
codes0 = torch.randn(S, 192) * 0.01
RR.optimize_codes(dec, codes0, pools, sched, 5e-3, 1.0, 1.0, 0.1, precision="bf16")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    e0.record()
    RR.optimize_codes(dec, codes0, pools, sched, 5e-3, 1.0, 1.0, 0.1, precision="bf16")
    e1.record(); torch.cuda.synchronize()
print("loop %.1f ms = %.1f us per iteration" % (e0.elapsed_time(e1), 1e3 * e0.elapsed_time(e1) / iters))
tot = 0
for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:10]:
    print("%-70s n=%5d total %9.1f us  mean %8.1f us" % (ev.key[:70], ev.count, ev.device_time_total, ev.device_time_total / max(ev.count, 1)))
    tot += ev.device_time_total
print("sum of kernel time %.1f ms" % (tot / 1e3))
