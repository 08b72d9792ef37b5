"""Golden run of the reference's reconstruct() with sampling_method="surface_interior_only" (reconstruct.py:60-98:
partial-view sampling -- interior uniform points + surface points, no normals), executed in place
(oracle/ref_harness.py) -> tests/golden/reconstruct_surface_interior.npz.  Build container only."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness, synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ns = ref_harness.load()
    torch.set_num_threads(8)
    xyz, sdf, _ = synth.make_sample_set(4000, seed=3)
    rec = {"xyz": xyz, "sdf": sdf}
    for kind in ("default", "trained_norm"):
        dec = ref_harness.make_decoder(ns, synth.make_state_dict(0, kind))
        np.random.seed(7)
        torch.manual_seed(7)
        loss, code = ns.reconstruct(dec, 12, 128, 64, [xyz, sdf], 0.01, 0.0, 0.0, 1.0, 0.1,
                                    sampling_method="surface_interior_only", num_samples=500, lr=5e-3, l2reg=True,
                                    code_reg_lambda=1e-4)
        rec[kind + "_code"] = code.detach().numpy()
        for k, val in loss.items():
            rec[kind + "_log_" + k] = np.asarray(val, dtype=np.float64)
        print(kind, {k: v for k, v in loss.items() if k in ("loss", "sdf_data_loss", "occ_data_loss", "norm_loss")})
    np.savez_compressed(os.path.join(OUT, "reconstruct_surface_interior.npz"), **rec)


if __name__ == "__main__":
    main()
