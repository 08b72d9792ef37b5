"""Split-operand (fp16x2) forward against the fp64 oracle on trained-norm weights + timing (GPU box)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import decoder_oracle as O, synth
from tests import util

def main():
    torch.cuda.set_device(0)
    rows = []
    for seed in (0, 1):
        net64 = util.oracle_net(seed, "trained_norm", "float64")
        code = util.trained_code(seed)
        pts = synth.make_points(20000, seed)
        for prec in ("fp16x2", "fp16", "bf16", "fp32"):
            dec = util.make_decoder(seed, "trained_norm", precision=prec)
            for use_net in (0, 1, 2, 3):
                ref = O.forward(net64, code.double(), pts.double(), use_net).numpy()
                out = dec.query(code.cuda(), pts.cuda(), use_net=use_net).cpu().numpy().astype(np.float64)
                err = np.abs(out - ref)
                flips = int(((out > 0) != (ref > 0)).sum())
                rows.append(dict(seed=seed, prec=prec, use_net=use_net, max=float(err.max()), mean=float(err.mean()), flips=flips))
                print(rows[-1], flush=True)
            if prec == "fp16x2":
                ref = O.forward(net64, code.double(), pts.double(), None)
                out = dec.query_all(code.cuda(), pts.cuda())
                for i, (a, b) in enumerate(zip(out, ref)):
                    e = float((a.cpu().double() - b).abs().max())
                    print("  query_all[%d] max err %.3e" % (i, e))
    # timing
    dec = util.make_decoder(0, "trained_norm", precision="fp16x2")
    code = util.trained_code(0).cuda()
    from deepjoin_b200.core import reconstruct as R
    for prec in ("fp16x2", "bf16"):
        dec.precision = prec
        for _ in range(2):
            R.decode_shape_device(code, dec, (256, 256, 256), use_net=1, padding=0.1, sigmoid=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 3
        for _ in range(n):
            R.decode_shape_device(code, dec, (256, 256, 256), use_net=1, padding=0.1, sigmoid=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print("256^3 use_net=1 %s: %.2f ms  %.3e q/s" % (prec, ms, 256**3 / ms * 1e3), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/x2_check.json", "w"), indent=1)

if __name__ == "__main__":
    main()
