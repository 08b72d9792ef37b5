"""Split-operand latent step (DJ_PREC_FP16X2) against autograd (fp64 oracle) + timing (GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import decoder_oracle as O, synth
from tests import util
from deepjoin_b200 import reconstruct as RR

def main():
    torch.cuda.set_device(0)
    for seed in (0, 1):
        net64 = util.oracle_net(seed, "trained_norm", "float64")
        code = util.trained_code(seed)
        dec = util.make_decoder(seed, "trained_norm", precision="fp16x2")
        xyz, sdf, nrm = synth.make_sample_set(40000, seed=seed + 1)
        for n in (1, 63, 64, 65, 512, 2000, 30000):
            x = torch.from_numpy(xyz[:n].astype(np.float32)); s = torch.from_numpy(sdf[:n].astype(np.float32)); m = torch.from_numpy(nrm[:n].astype(np.float32))
            gr, tr = O.latent_grad(net64, code.double(), x.double(), s.double(), m.double())
            for prec in ("fp16x2", "fp16", "bf16", "fp32"):
                g, t = RR.latent_grad(dec, code.cuda(), x.cuda(), s.cuda(), m.cuda(), precision=prec)
                g = g.cpu().double()
                rel = float((g - gr).abs().max() / gr.abs().max())
                cos = float((g * gr).sum() / (g.norm() * gr.norm()))
                print("seed %d n=%5d %-6s grad rel err %.3e cos %.8f  loss %.7f (ref %.7f) terms err %.2e" % (
                    seed, n, prec, rel, cos, float(t[0]), tr[0], max(abs(float(t[i]) - tr[i]) for i in range(5))), flush=True)
    # DeepSDF objective (second net skipped)
    # timing: 8 shapes x 200 iterations x 30000 samples
    dec = util.make_decoder(0, "trained_norm", precision="fp16x2")
    pools = []
    for sh in range(8):
        xyz_s, sdf_s, nrm_s = synth.make_sample_set(100000, seed=sh)
        pools.append((torch.from_numpy(xyz_s.astype(np.float32)).cuda(), torch.from_numpy(sdf_s.astype(np.float32).reshape(-1)).cuda(), torch.from_numpy(nrm_s.astype(np.float32)).cuda()))
    sched = torch.randint(0, 100000, (8, 200, 30000), dtype=torch.int32, device="cuda")
    codes0 = torch.cat([RR.init_code(128, 64, 0.01) for _ in range(8)])
    for prec in ("fp16x2", "fp16", "bf16"):
        for S in (8, 1):
            RR.optimize_codes(dec, codes0[:S], pools[:S], sched[:S, :10], 5e-3, 1.0, 1.0, 0.1, l2reg=True, precision=prec)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            logs, codes = RR.optimize_codes(dec, codes0[:S], pools[:S], sched[:S], 5e-3, 1.0, 1.0, 0.1, l2reg=True, precision=prec)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print("%-6s %d shapes: %.3f ms per shape-iteration, %.0f TFLOP/s useful, loss %.4f -> %.4f" % (
                prec, S, ms / (200 * S), S * 200 * 30000 * 11.03e6 / (ms * 1e-3) / 1e12, float(logs[0, 0, 1]), float(logs[0, -1, 1])), flush=True)

if __name__ == "__main__":
    main()
