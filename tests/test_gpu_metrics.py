"""the nearest-neighbour metrics on the GPU against scipy's cKDTree (what the reference calls)."""
import numpy as np
import pytest
import torch
from scipy.spatial import cKDTree

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,m", [(1, 1), (2000, 2000), (30000, 30000), (777, 5001)])
def test_nn_query_equals_ckdtree(n, m):
    from deepjoin_b200.core import metrics as M
    rng = np.random.default_rng(n + m)
    q = rng.uniform(-0.5, 0.5, (n, 3)).astype(np.float32)
    t = rng.uniform(-0.5, 0.5, (m, 3)).astype(np.float32)
    d, i = M.nn_query(q, t)
    dr, ir = cKDTree(t.astype(np.float64)).query(q.astype(np.float64))
    assert np.allclose(d, dr, rtol=1e-5, atol=1e-7)
    same = i == ir
    # fp32 distances: a different index only where two targets are (numerically) equally near
    assert same.mean() > 0.999 and np.allclose(np.linalg.norm(q[~same] - t[i[~same]], axis=1), dr[~same], rtol=1e-5, atol=1e-7)


def test_chamfer_and_normal_consistency_on_spheres():
    from deepjoin_b200.core import errors, metrics as M
    from deepjoin_b200.core import reconstruct as R
    from deepjoin_b200.mesh import Mesh

    def sphere(radius, n=48):
        ax = (np.arange(n) - (n - 1) / 2) / n
        vol = (np.sqrt(ax[:, None, None] ** 2 + ax[None, :, None] ** 2 + ax[None, None, :] ** 2) - radius).astype(np.float32)
        v, f = R.marching_cubes(torch.from_numpy(vol).cuda(), 0.0, "ascent")
        return Mesh(v.cpu().numpy().astype(np.float64) / n, f.cpu().numpy(), process=False)

    np.random.seed(0)
    a, b = sphere(0.30), sphere(0.35)
    c_same = M.chamfer(a, a, 2000)
    c_diff = M.chamfer(a, b, 2000)
    # two independent 2000-point samples of one sphere: E[nn dist^2] ~ area / (pi N) per direction (3.6e-4 in total);
    # concentric spheres add ~ (dr)^2 per direction
    assert 1e-4 < c_same < 6e-4 and abs(c_diff - 2 * 0.05 ** 2) < 1.5e-3
    nc = M.normal_consistency(a, b, 5000)
    assert nc > 0.97
    pts = M.sample_surface(a, 2000)[0]
    assert abs(np.linalg.norm(pts - pts.mean(0), axis=1).mean() - 0.30) < 0.01
    assert abs(M.chamfer(pts, a, 2000) - c_same) < 3e-4
    with pytest.raises(errors.MeshEmptyError):
        M.chamfer(a, Mesh(), 2000)
