#!/bin/bash
# ncu evidence for the latent step: launch list of a short loop + one full capture of the fused kernel
mkdir -p gpurun_out
cat > /tmp/lat_cmd.py <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import synth_data
from bench import make_decoder
from deepjoin_b200 import reconstruct as RR
dev = torch.device("cuda", 0)
dec = make_decoder(0, dev, "bf16")
xyz, sdf, nrm = synth_data.make_sample_set(100000, seed=0)
px = torch.from_numpy(xyz.astype(np.float32)).to(dev); ps = torch.from_numpy(sdf.astype(np.float32).reshape(-1)).to(dev); pn = torch.from_numpy(nrm.astype(np.float32)).to(dev)
sched = torch.randint(0, 100000, (12, 30000), device=dev, dtype=torch.int32)
log, code = RR.optimize_code(dec, RR.init_code(128, 64, 0.01), px, ps, pn, sched, 5e-3, 1.0, 1.0, 0.1, l2reg=True, precision="bf16")
torch.cuda.synchronize()
print("loss", float(log[0, 1]), float(log[-1, 1]))
PY
CMD="python /tmp/lat_cmd.py"
$CMD > gpurun_out/lat_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_latent.csv $CMD > gpurun_out/ncu_launches_latent.log 2>&1
$CMD > gpurun_out/lat_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:latent_tc2 -s 5 -c 1 -f -o gpurun_out/prof_latent_tc2 $CMD > gpurun_out/ncu_full_latent.log 2>&1
tail -2 gpurun_out/lat_plain.log; tail -2 gpurun_out/ncu_full_latent.log
