"""Property tests (hypothesis) of the host-side logic around the hot path: shard arithmetic, slab merging against
the C oracle's single sweep for random volumes and cuts, the fair-window sampler, the strided-point contract."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from deepjoin_b200 import parallel as P
from deepjoin_b200.core import data, errors
from oracle import mc_oracle
from tests import slab_util

SET = settings(max_examples=40, deadline=None, derandomize=True, database=None)


@SET
@given(n0=st.integers(2, 700), world=st.integers(1, 16))
def test_slab_ranges_tile_the_cell_layers(n0, world):
    r = P.slab_ranges(n0, world)
    assert len(r) == world and r[0][0] == 0 and r[-1][1] == n0 - 1
    assert all(lo <= hi for lo, hi in r) and all(a[1] == b[0] for a, b in zip(r, r[1:]))   # one shared plane per seam
    assert sum(hi - lo for lo, hi in r) == n0 - 1                                          # every cell layer once
    sizes = [hi - lo for lo, hi in r if hi > lo]
    assert max(sizes) - min(sizes[:-1] or sizes) <= max(sizes)                             # contiguous ceil-sized slabs
    assert all(s == -(-(n0 - 1) // world) for s in sizes[:-1])


@SET
@given(n=st.integers(0, 500), world=st.integers(1, 16))
def test_shard_items_partition(n, world):
    parts = [list(P.shard_items(n, world, k)) for k in range(world)]
    assert sum(parts, []) == list(range(n))
    assert all(len(p) <= -(-n // world) for p in parts) if n else all(not p for p in parts)


@settings(max_examples=25, deadline=None, derandomize=True, database=None)
@given(seed=st.integers(0, 10 ** 6), n0=st.integers(3, 14), n1=st.integers(2, 9), n2=st.integers(2, 9),
       world=st.integers(1, 5), level=st.floats(-0.3, 0.3))
def test_random_slab_cuts_reproduce_the_single_sweep(seed, n0, n1, n2, world, level):
    """Any volume, any number of slabs (including empty ones): merged fragments == one sequential sweep."""
    rng = np.random.default_rng(seed)
    vol = rng.uniform(-1, 1, (n0, n1, n2)).astype(np.float32)   # noise: every ambiguous case, no sample on the level
    level = float(np.float32(level))
    try:
        vo, fo = mc_oracle.marching_cubes(vol, level, "ascent")
    except (RuntimeError, ValueError):
        return
    frags = [slab_util.oracle_slab_fragment(vol, lo, hi, seam=k > 0 and lo > 0, level=level)
             for k, (lo, hi) in enumerate(P.slab_ranges(n0, world))]
    v, f = P.merge_slab_meshes([(a, b) for a, b, _ in frags], [c for _, _, c in frags])
    assert np.array_equal(f, fo) and np.allclose(v, vo, rtol=0, atol=2e-6)
    tv, tf = P.merge_fragments([tuple(torch.from_numpy(np.ascontiguousarray(x)) for x in fr) for fr in frags])
    assert np.array_equal(tf.numpy(), fo)


@SET
@given(seed=st.integers(0, 10 ** 6), n=st.integers(4, 400), frac_pos=st.floats(0.0, 1.0), num=st.integers(1, 60))
def test_fair_windows_are_balanced_contiguous_runs(seed, n, frac_pos, num):
    rng = np.random.default_rng(seed)
    d = np.where(rng.random(n) < frac_pos, 1.0, -1.0).astype(np.float32)
    np.random.seed(seed % (2 ** 31))
    pos_num, neg_num = num // 2, num - num // 2
    try:
        idx = data._fair_windows(d, num)
    except errors.NoSamplesError:
        assert (pos_num and not (d > 0).any()) or (neg_num and not (d <= 0).any())
        return
    assert idx.shape == (num,)
    assert (d[idx[:pos_num]] > 0).all() and (d[idx[pos_num:]] <= 0).all()      # [pos..., neg...] (core/data.py:46-73)
    for part, cls in ((idx[:pos_num], np.where(d > 0)[0]), (idx[pos_num:], np.where(d <= 0)[0])):
        if len(part) and len(cls) > len(part):                                   # a contiguous window of the class's rows
            start = np.searchsorted(cls, part[0])
            assert np.array_equal(part, cls[start:start + len(part)])


@SET
@given(n=st.integers(2, 50), layout=st.sampled_from(["rows", "soa", "strided"]))
def test_point_layouts_accepted_without_a_copy_are_exactly_rows_or_coordinate_arrays(n, layout):
    """The stride contract of dj_decode_samples_host_strided as core.reconstruct.decode_samples applies it."""
    base = np.arange(3 * n, dtype=np.float64)
    pts = {"rows": base.reshape(n, 3), "soa": base.reshape(3, n).T, "strided": np.arange(6 * n, dtype=np.float64).reshape(n, 6)[:, ::2]}[layout]
    s = (pts.strides[0] // 8, pts.strides[1] // 8)
    direct = s == (3, 1) or (s[0] == 1 and s[1] >= n and n > 1)
    assert direct == (layout != "strided")
    if layout == "soa":   # element (i, d) at base[i*1 + d*n]
        assert all(pts[i, d] == base[i * s[0] + d * s[1]] for i in (0, n - 1) for d in range(3))
