"""Extra grid fixtures made by EXECUTING THE REFERENCE in place (oracle/ref_harness.py; build container only):
dims with an axis of length 1 (np.linspace(0, stop, 1) is [0.0]) -> tests/golden/grid_extra.npz

    python tests/golden/make_golden_grid_extra.py
"""
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
    rows = np.load(os.path.join(OUT, "latent_rows.npz"))
    code = torch.from_numpy(np.concatenate([rows["c2000"][0], rows["c2000_tool"][0]]))
    dec = ref_harness.make_decoder(ns, synth.make_state_dict(0, "trained_norm"))
    grid = {"code": code.numpy()}
    with torch.no_grad():
        for dims in ((1, 9, 7), (8, 1, 5), (4, 6, 1), (1, 1, 1)):
            t = "d%dx%dx%d" % dims
            grid[t + "_pts"] = ns.core_reconstruct.get_query_points(dims, 0.1)
            for u in (1, 3):
                grid[t + "_u%d_flat" % u] = ns.core_reconstruct.decode_shape(
                    code, dec, dims, use_net=u, batch_size=500, padding=0.1, reshape=False, sigmoid=False)
    np.savez_compressed(os.path.join(OUT, "grid_extra.npz"), **grid)


if __name__ == "__main__":
    main()
