"""CPU stand-in for one rank's marching-cubes slab (deepjoin_b200.parallel.slab_fragment) built from the
C oracle: run the oracle on the slab's sub-volume, then re-express it in the GPU kernels' seam
convention (deepjoin_b200/csrc/mc.cuh): x/y-edge vertices on the slab's first plane belong to the
slab below -- they are dropped, later ids close up, and faces refer to them as -2 - (3*sample + dir);
the top-plane table maps 3*sample + dir of the slab's LAST plane to the slab-local vertex id."""
import numpy as np

from oracle import mc_oracle


def _plane_slots(v, plane, n2):
    """slot = 3*(y*n2 + x) + dir for vertices lying on sample plane `plane` (axis-0 coordinate)."""
    on = np.where(v[:, 0] == np.float32(plane))[0]
    y, x = v[on, 1], v[on, 2]
    fy, fx = np.floor(y), np.floor(x)
    is_x_edge = x != fx          # x-edge: the axis-2 coordinate is fractional
    is_y_edge = y != fy
    assert (is_x_edge ^ is_y_edge).all(), "vertex exactly on a grid sample: use a noise field"
    d = np.where(is_x_edge, 0, 1)
    slot = 3 * (fy.astype(np.int64) * n2 + fx.astype(np.int64)) + d
    return on, slot


def oracle_slab_fragment(vol, z_lo, z_hi, seam, level=0.0):
    """-> (verts f32 [V,3] global index units, faces int32 [F,3], top int32 [n1*n2*3])."""
    n1, n2 = vol.shape[1], vol.shape[2]
    top = np.full(n1 * n2 * 3, -1, np.int32)
    if z_hi - z_lo < 1:
        return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32), top
    sub = np.ascontiguousarray(vol[z_lo:z_hi + 1])
    try:
        v, f = mc_oracle.marching_cubes(sub, level, "ascent")
    except RuntimeError:
        return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32), top
    f = f.astype(np.int64)
    newid = np.arange(len(v), dtype=np.int64)
    if seam:
        on, slot = _plane_slots(v, 0, n2)
        keep = np.ones(len(v), bool)
        keep[on] = False
        newid = np.cumsum(keep) - 1
        newid[on] = -2 - slot
        f = newid[f]
        v = v[keep]
    on, slot = _plane_slots(v, z_hi - z_lo, n2)
    top[slot] = on
    v = v.copy()
    v[:, 0] += np.float32(z_lo)
    return v, f.astype(np.int32), top
