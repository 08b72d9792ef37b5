#!/bin/bash
# ncu full capture of the fused latent kernel with 8 shapes per launch (configs[2] steady state)
mkdir -p gpurun_out
cat > /tmp/lat8_cmd.py <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import synth_data
from bench import make_decoder
from deepjoin_b200 import reconstruct as RR
dev = torch.device("cuda", 0)
dec = make_decoder(0, dev, "bf16")
S, iters, ns = 8, 8, 30000
pools, codes0 = [], []
for s in range(S):
    x, d, n = synth_data.make_sample_set(100000, seed=s)
    pools.append(tuple(torch.from_numpy(a.astype(np.float32)).to(dev) for a in (x, d.reshape(-1), n)))
    torch.manual_seed(s); codes0.append(RR.init_code(128, 64, 0.01))
sched = torch.randint(0, 100000, (S, iters, ns), device=dev, dtype=torch.int32)
logs, codes = RR.optimize_codes(dec, torch.cat(codes0), pools, sched, 5e-3, 1.0, 1.0, 0.1, l2reg=True, precision="bf16")
torch.cuda.synchronize()
print("loss", float(logs[0, 0, 1]))
PY
CMD="python /tmp/lat8_cmd.py"
$CMD > gpurun_out/lat8_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:latent_tc2 -s 4 -c 1 -f -o gpurun_out/prof_latent8_tc2 $CMD > gpurun_out/ncu_full_latent8.log 2>&1
tail -1 gpurun_out/lat8_plain.log; tail -2 gpurun_out/ncu_full_latent8.log
