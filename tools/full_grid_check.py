"""Whole-grid comparison of the split-operand forward with the fp32 CUDA-core path (both product kernels, no oracle):
every tile / pass of the persistent kernel, not just the first few.  python tools/full_grid_check.py [dims]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synth_data
from bench import make_decoder
from deepjoin_b200.core import reconstruct as R


def main():
    d = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    code = synth_data.trained_code(0).to(dev)
    vols = {}
    for prec in ("fp32", "fp16x2"):
        dec = make_decoder(0, dev, prec)
        for u in (1, 0):
            vols[(prec, u)] = R.decode_shape_device(code, dec, (d, d, d), use_net=u, padding=0.1, sigmoid=False).double()
    for u in (1, 0):
        a, b = vols[("fp32", u)], vols[("fp16x2", u)]
        diff = (a - b).abs()
        flips = ((a > 0) != (b > 0))
        print("use_net %d, %d^3 = %d points: max |fp16x2 - fp32| %.3e, mean %.3e, sign differences %d (largest |fp32| among them %.3e)"
              % (u, d, a.numel(), diff.max().item(), diff.mean().item(), int(flips.sum()), a[flips].abs().max().item() if flips.any() else 0.0), flush=True)
        worst = int(diff.argmax())
        print("   worst point: flat row %d (tile %d, row %d of the tile)" % (worst, worst // 64, worst % 64))


if __name__ == "__main__":
    main()
