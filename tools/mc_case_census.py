"""Census of marching-cubes configurations on a decoder volume: how many surface cells fall into each of the 15 base
cases (Chernyaev / Lewiner numbering), how many have an ambiguous face, and how many are in the cases where Lewiner's
implementation may additionally run its interior test (4, 6, 7, 10, 12, 13) -- the cells in which this repo's tiling
(face tests only) can differ from Lewiner's in TOPOLOGY, not just in how a polygon is split.

    python tools/mc_case_census.py --dims 256 --precision fp16x2          # product path on cuda:0
    python tools/mc_case_census.py --npy volume.npy --level 0.0           # any saved volume, CPU only

The base case of a configuration is found from scratch here (minority corner set -> multiset of pairwise squared
distances, which separates the 15 classes up to the 11/14 mirror pair); no table of the product or oracle is used."""
import argparse
import itertools
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

CORNER_XYZ = ((0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1))
# one representative corner set per base case
REPRESENTATIVES = {1: (0,), 2: (0, 1), 3: (0, 2), 4: (0, 6), 5: (0, 1, 2), 6: (0, 1, 6), 7: (0, 2, 5), 8: (0, 1, 2, 3),
                   9: (0, 1, 3, 4), 10: (0, 1, 6, 7), 11: (0, 1, 2, 6), 12: (0, 1, 2, 7), 13: (0, 2, 5, 7)}
INTERIOR_TEST_CASES = (4, 6, 7, 10, 12, 13)
AMBIGUOUS_FACE_CASES = (3, 6, 7, 10, 12, 13)


def signature(corners):
    d = sorted(sum((a - b) ** 2 for a, b in zip(CORNER_XYZ[i], CORNER_XYZ[j])) for i, j in itertools.combinations(corners, 2))
    return (len(corners), tuple(d))


def base_cases():
    """-> int array [256]: base case of every configuration (0 for the two empty ones; 11 stands for 11 and 14)."""
    by_sig = {signature(c): k for k, c in REPRESENTATIVES.items()}
    assert len(by_sig) == len(REPRESENTATIVES)
    out = np.zeros(256, dtype=np.int64)
    for cfg in range(1, 255):
        on = [i for i in range(8) if cfg >> i & 1]
        if len(on) > 4:
            on = [i for i in range(8) if not cfg >> i & 1]
        out[cfg] = by_sig[signature(on)]
    return out


def census(volume, level):
    """volume [d0,d1,d2] indexed [x][y][z] as the reference's grid is (corner i of a cell at CORNER_XYZ[i])."""
    above = (volume.astype(np.float64) - level) > 0
    cfg = np.zeros(tuple(d - 1 for d in volume.shape), dtype=np.int64)
    for i, (x, y, z) in enumerate(CORNER_XYZ):
        cfg |= above[x:x + cfg.shape[0], y:y + cfg.shape[1], z:z + cfg.shape[2]].astype(np.int64) << i
    cases = base_cases()[cfg]
    hist = np.bincount(cases.ravel(), minlength=14)
    surface = int(hist[1:].sum())
    return {"cells": int(cfg.size), "surface_cells": surface, "by_case": {int(k): int(hist[k]) for k in range(1, 14) if hist[k]},
            "ambiguous_face_cells": int(sum(hist[k] for k in AMBIGUOUS_FACE_CASES)),
            "interior_test_candidates": int(sum(hist[k] for k in INTERIOR_TEST_CASES))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, default=256)
    ap.add_argument("--precision", default="fp16x2")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--level", type=float, default=0.0)
    ap.add_argument("--npy", default=None)
    args = ap.parse_args()
    if args.npy:
        vol = np.load(args.npy)
    else:
        import torch
        import bench
        import synth_data
        from deepjoin_b200.core import reconstruct as R
        dev = torch.device("cuda", 0)
        dec = bench.make_decoder(args.seed, dev, args.precision)
        code = synth_data.trained_code(args.seed).to(dev)
        vol = R.decode_shape_device(code, dec, (args.dims,) * 3, use_net=1, padding=0.1, sigmoid=False)
        vol = vol.reshape((args.dims,) * 3).cpu().numpy()
    import json
    r = census(vol, args.level)
    r["volume"] = args.npy or "synthetic trained-norm decoder seed %d, %d^3, %s" % (args.seed, args.dims, args.precision)
    print(json.dumps(r))


if __name__ == "__main__":
    main()
