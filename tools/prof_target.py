"""Short profiling targets for ncu (B200_PROFILING.md): python tools/prof_target.py fwd|lat|mc [precision]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import synth_data
from bench import make_decoder
from deepjoin_b200.core import reconstruct as R
from deepjoin_b200 import reconstruct as RR

def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
    prec = sys.argv[2] if len(sys.argv) > 2 else "fp16x2"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dec = make_decoder(0, dev, prec)
    code = synth_data.trained_code(0).to(dev)
    if what in ("fwd", "mc"):
        dims = (256, 256, 256)
        for _ in range(3):
            vol = R.decode_shape_device(code, dec, dims, use_net=1, padding=0.1, sigmoid=False)
            v, f = R.marching_cubes(vol.reshape(dims), 0.0, "ascent")
        torch.cuda.synchronize()
        print("fwd ok", prec, v.shape, f.shape)
    else:
        S, it, ns = 8, 4, 30000
        pools = []
        for sh in range(S):
            x, s, n = synth_data.make_sample_set(100000, seed=sh)
            pools.append((torch.from_numpy(x.astype(np.float32)).to(dev), torch.from_numpy(s.astype(np.float32).reshape(-1)).to(dev), torch.from_numpy(n.astype(np.float32)).to(dev)))
        sched = torch.randint(0, 100000, (S, it, ns), dtype=torch.int32, device=dev)
        codes0 = torch.cat([RR.init_code(128, 64, 0.01) for _ in range(S)])
        logs, codes = RR.optimize_codes(dec, codes0, pools, sched, 5e-3, 1.0, 1.0, 0.1, l2reg=True, precision=prec)
        torch.cuda.synchronize()
        print("lat ok", prec, float(logs[0, 0, 1]))

if __name__ == "__main__":
    main()
