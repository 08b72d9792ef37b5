"""Where the per-shape time of configs[2] goes: schedule draw, latent loop (single shape and batch of 8), mesh."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import synth_data
from bench import make_decoder
from deepjoin_b200 import reconstruct as RR, _lib
from deepjoin_b200.core import reconstruct as R
dev = torch.device("cuda", 0)
dec = make_decoder(0, dev, "bf16")
S = 8
pools = []
for s in range(S):
    xyz, sdf, nrm = synth_data.make_sample_set(100000, seed=s)
    pools.append(tuple(torch.from_numpy(a.astype(np.float32)).to(dev) for a in (xyz, sdf.reshape(-1), nrm)))
def ev(): return torch.cuda.Event(enable_timing=True)
for iters in (100, 800):
    sched = torch.randint(0, 100000, (S, iters, 30000), device=dev, dtype=torch.int32)
    codes0 = torch.cat([RR.init_code(128, 64, 0.01) for _ in range(S)])
    for nb in (1, 2, 4, 8):
        RR.optimize_codes(dec, codes0[:nb], pools[:nb], sched[:nb, :5], 5e-3, 1.0, 1.0, 0.1, precision="bf16")
        e0, e1 = ev(), ev()
        e0.record()
        logs, codes = RR.optimize_codes(dec, codes0[:nb], pools[:nb], sched[:nb], 5e-3, 1.0, 1.0, 0.1, precision="bf16")
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("iters %d batch %d: %.1f ms total, %.3f ms/iter, %.3f ms per shape-iteration, %.0f TF alg" % (iters, nb, ms, ms/iters, ms/iters/nb, nb*30000*11.03e6/(ms/iters*1e-3)/1e12), flush=True)
e0, e1 = ev(), ev()
e0.record()
m = R.create_mesh(codes[0], dec, 2, sigmoid=False, level=0.0, gradient_direction="ascent", dims=(256,)*3, padding=0.1)
e1.record(); torch.cuda.synchronize()
print("mesh %.1f ms" % e0.elapsed_time(e1), len(m.vertices))
