"""Small end-to-end pass over every kernel family (for compute-sanitizer --tool memcheck; one tool per gpurun call)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import synth_data
from bench import make_decoder
from deepjoin_b200 import reconstruct as RR
from deepjoin_b200.core import reconstruct as R
dev = torch.device("cuda", 0)
code = synth_data.trained_code(0)
pts = synth_data.make_points(1000, 0).to(dev)
for prec in ("bf16", "fp32"):
    dec = make_decoder(0, dev, prec)
    for u in (0, 1, 2, 3):
        out = dec.query(code.to(dev), pts, use_net=u)
    allo = dec.query_all(code.to(dev), pts)
    mesh = R.create_mesh(code, dec, 1, sigmoid=False, level=0.0, gradient_direction="ascent", dims=(24, 20, 28), padding=0.1)
    print(prec, "forward ok", float(out.abs().max()), len(mesh.vertices), len(mesh.faces), flush=True)
xyz, sdf, nrm = synth_data.make_sample_set(4000, seed=1)
x, s, n = (torch.from_numpy(a[:700].astype(np.float32)).to(dev) for a in (xyz, sdf.reshape(-1), nrm))
dec = make_decoder(0, dev, "bf16")
for prec in ("bf16", "fp32"):
    g, t = RR.latent_grad(dec, code.to(dev), x, s, n, precision=prec)
    print(prec, "latent grad ok", float(g.abs().max()), float(t[0]), flush=True)
pools = [tuple(torch.from_numpy(a.astype(np.float32)).to(dev) for a in (xyz, sdf.reshape(-1), nrm))] * 2
sched = torch.randint(0, 4000, (2, 3, 300), device=dev, dtype=torch.int32)
logs, codes = RR.optimize_codes(dec, torch.cat([RR.init_code(128, 64, 0.01)] * 2), pools, sched, 5e-3, 1.0, 1.0, 0.1, precision="bf16")
torch.cuda.synchronize()
print("batch ok", float(logs[0, 0, 1]))
