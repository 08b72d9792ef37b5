import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import synth_data
from bench import make_decoder
from deepjoin_b200 import reconstruct as RR, _lib
from deepjoin_b200.core import reconstruct as R
dev = torch.device("cuda", 0)
dec = make_decoder(0, dev, "bf16")
xyz, sdf, nrm = synth_data.make_sample_set(100000, seed=0)
px = torch.from_numpy(xyz.astype(np.float32)).to(dev); ps = torch.from_numpy(sdf.astype(np.float32).reshape(-1)).to(dev); pn = torch.from_numpy(nrm.astype(np.float32)).to(dev)
sched = torch.empty((800, 30000), device=dev, dtype=torch.int32)
def ev(): return torch.cuda.Event(enable_timing=True)
for rep in range(2):
    e = [ev() for _ in range(5)]
    e[0].record()
    _lib.check(_lib.lib().dj_sample_schedule(ps.data_ptr(), ps.shape[0], 800, 30000, 0.2, 1000, sched.data_ptr(), _lib.stream_ptr()))
    e[1].record()
    log, c = RR.optimize_code(dec, RR.init_code(128, 64, 0.01), px, ps, pn, sched, 5e-3, 1.0, 1.0, 0.1, l2reg=True, precision="bf16")
    e[2].record()
    t0 = time.perf_counter()
    m = R.create_mesh(c, dec, 2, sigmoid=False, level=0.0, gradient_direction="ascent", dims=(256,)*3, padding=0.1)
    e[3].record()
    torch.cuda.synchronize()
    print("schedule %.1f ms, latent %.1f ms, mesh %.1f ms (wall %.1f)" % (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), 1e3*(time.perf_counter()-t0)), len(m.vertices))
