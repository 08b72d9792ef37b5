"""configs[2] pieces on one GPU: schedule, Adam loop, then the per-shape restoration meshes, each timed."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import synth_data
from bench import make_decoder
from deepjoin_b200 import _lib, reconstruct as RR
from deepjoin_b200.core import reconstruct as R
dev = torch.device("cuda", 0)
dec = make_decoder(0, dev, "bf16")
n_sh, iters, ns = 8, int(os.environ.get("ITERS", "800")), 30000
pools = []
for sh in range(n_sh):
    x, s, n = synth_data.make_sample_set(100000, seed=sh)
    pools.append(tuple(torch.from_numpy(a.astype(np.float32)).to(dev) for a in (x, s.reshape(-1), n)))
sched = torch.empty((n_sh, iters, ns), device=dev, dtype=torch.int32)
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(2):
    t0 = sync()
    for sh in range(n_sh):
        _lib.check(_lib.lib().dj_sample_schedule(pools[sh][1].data_ptr(), pools[sh][1].shape[0], iters, ns, 0.2, 1000 + sh, sched[sh].data_ptr(), _lib.stream_ptr()))
    t1 = sync()
    codes0 = []
    for sh in range(n_sh):
        torch.manual_seed(sh); codes0.append(RR.init_code(128, 64, 0.01))
    logs, codes = RR.optimize_codes(dec, torch.cat(codes0), pools, sched, 5e-3, 1.0, 1.0, 0.1, l2reg=True, precision="bf16")
    t2 = sync()
    tm = []
    for sh in range(n_sh):
        a = time.perf_counter()
        try:
            m = R.create_mesh(codes[sh], dec, 2, sigmoid=False, level=0.0, gradient_direction="ascent", dims=(256,) * 3, padding=0.1)
            nv = len(m.vertices)
        except Exception as e:
            nv = -1
        tm.append((round(1e3 * (time.perf_counter() - a), 1), nv))
    t3 = sync()
    print("rep %d: schedule %.1f ms, adam loop %.1f ms, meshes %.1f ms" % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)), tm)
