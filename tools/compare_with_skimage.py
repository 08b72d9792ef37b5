"""Compare deepjoin_b200's marching cubes (oracle restatement, and the GPU kernels if a GPU is
present) with the real scikit-image, wherever it is installed (it is NOT in the offline build image,
which is why DESIGN.md calls marching-cubes parity 'unpinned').

    python tools/compare_with_skimage.py [N]

Reports, on sphere / torus / gyroid+noise volumes: vertex and face counts, whether faces are identical
(same order and ids), and otherwise the size of the symmetric difference of the triangle sets keyed by
quantised vertex positions."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def compare(n=64, verbose=True):
    """-> list of dicts (one per volume and direction).  Raises ImportError without scikit-image."""
    from skimage import measure
    from oracle import mc_oracle
    ax = np.linspace(-1, 1, n)
    z, y, x = np.meshgrid(ax, ax, ax, indexing="ij")
    rng = np.random.default_rng(0)
    vols = {
        "sphere": np.sqrt(x * x + y * y + z * z) - 0.7,
        "torus": np.sqrt((np.sqrt(x * x + y * y) - 0.55) ** 2 + z * z) - 0.25,
        "gyroid+noise": np.sin(4 * x) * np.cos(4 * y) + np.sin(4 * y) * np.cos(4 * z) + np.sin(4 * z) * np.cos(4 * x)
        + 0.2 * rng.uniform(-1, 1, x.shape),
    }
    rows = []
    for name, vol in vols.items():
        vol = vol.astype(np.float32)
        for direction in ("descent", "ascent"):
            vs, fs, _, _ = measure.marching_cubes(vol, level=0.0, gradient_direction=direction)
            vo, fo = mc_oracle.marching_cubes(vol, 0.0, direction)
            same = vs.shape == vo.shape and fs.shape == fo.shape and np.array_equal(fs, fo)
            row = dict(volume=name, direction=direction, skimage=(len(vs), len(fs)), ours=(len(vo), len(fo)), faces_identical=bool(same))
            msg = "%-13s %-7s skimage V/F %d/%d ours %d/%d faces identical: %s" % (
                name, direction, len(vs), len(fs), len(vo), len(fo), same)
            if same:
                row["max_vertex_diff"] = float(np.abs(vs - vo).max())
                msg += " | max vertex diff %.3e" % row["max_vertex_diff"]
            else:
                q = lambda v, f: {tuple(sorted(map(tuple, np.round(v[t] * 4096).astype(np.int64)))) for t in f}
                a, b = q(vs, fs), q(vo, fo)
                row.update(common=len(a & b), only_skimage=len(a - b), only_ours=len(b - a))
                msg += " | triangle sets: common %d, only skimage %d, only ours %d" % (len(a & b), len(a - b), len(b - a))
            rows.append(row)
            if verbose:
                print(msg)
    return rows


def main():
    try:
        compare(int(sys.argv[1]) if len(sys.argv) > 1 else 64)
    except ImportError:
        print("scikit-image is not installed: nothing to compare (parity stays unpinned)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
