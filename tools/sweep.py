#!/usr/bin/env python
"""BASELINE configs[4]: decoder query throughput sweep -- P points (random in the unit cube, resident in HBM) x
precision x use_net, CodeLength 128 and 256.  One JSON line per case (device-timed, CUDA events).
Usage: python tools/sweep.py [--max-log2 30]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
F256 = 2 * (3 * 512 + 512 * 512 * 2 + 253 * 512 + 3 * 512 + 256 * 512 + 3 * 512 * 512 + 5 * 512)
FLOP = {128: {0: 3412992, 1: 5518336}, 256: {0: F256, 1: F256 + 2105344}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-log2", type=int, default=26)
    ap.add_argument("--fp32-max-log2", type=int, default=22)
    args = ap.parse_args()
    import torch
    import synth_data
    from bench import SPECS
    from deepjoin_b200.networks.decoder_z_lb_both_leaky_n_sdf import Decoder
    dev = torch.device("cuda", 0)
    for L in (128, 256):
        dec = Decoder(latent_size=L, tool_latent_size=64, num_dims=3, do_code_regularization=True,
                      **SPECS["NetworkSpecs"], **SPECS["SubnetSpecs"])
        dec.load_state_dict(synth_data.make_state_dict(0, "default", latent_size=L, tool_latent_size=64))
        dec = dec.to(dev).eval()
        code = synth_data.make_code(0, L, 64).to(dev)
        for log2 in range(20, args.max_log2 + 1, 2):
            n = 1 << log2
            pts = torch.rand((n, 3), device=dev) * 1.2 - 0.6
            # fp16x2: the parity mode (split operands).  fp16: IEEE half = tf32's 10-bit mantissa at the bf16 rate; a
            # kind::tf32 instantiation of the fused kernel does not exist (4-byte operands: a 128-row layer input is
            # 256 KB, neither shared nor tensor memory holds it next to the accumulators -- DESIGN.md 3.1)
            for precision in ("fp16x2", "bf16", "fp16", "fp32"):
                if precision == "fp32" and log2 > args.fp32_max_log2:
                    continue
                for u in (0, 1):
                    for _ in range(2 if log2 <= 26 else 1):
                        dec.query(code, pts, use_net=u, precision=precision)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    reps = 3 if log2 <= 24 else 1
                    e0.record()
                    for _ in range(reps):
                        dec.query(code, pts, use_net=u, precision=precision)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / reps
                    flop = FLOP.get(L, {}).get(u)
                    print(json.dumps({"code_length": L, "points": n, "precision": precision, "use_net": u, "ms": ms,
                                      "queries_per_s": n / (ms * 1e-3),
                                      "tflops_alg": (n * flop / (ms * 1e-3) / 1e12) if flop else None}), flush=True)
            del pts


if __name__ == "__main__":
    main()
