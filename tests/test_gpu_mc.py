"""GPU marching cubes == the oracle's sequential sweep: identical faces (ids and order),
bit-identical float32 vertices."""
import numpy as np
import pytest
import torch

from oracle import mc_oracle
from tests.test_mc_cpu import _closed_manifold, _fields

pytestmark = pytest.mark.gpu


def _gpu_mc(vol, level, direction):
    from deepjoin_b200.core import reconstruct as R
    v, f = R.marching_cubes(torch.from_numpy(np.ascontiguousarray(vol, np.float32)).cuda(), level, direction)
    return v.cpu().numpy(), f.cpu().numpy()


@pytest.mark.parametrize("name", ["sphere", "torus", "gyroid"])
@pytest.mark.parametrize("n", [33, 64])
def test_analytic_fields_bit_exact(name, n):
    vol = _fields(n)[name].astype(np.float32)
    for direction in ("ascent", "descent"):
        v, f = _gpu_mc(vol, 0.0, direction)
        vo, fo = mc_oracle.marching_cubes(vol, 0.0, direction)
        assert f.dtype == np.int32 and v.dtype == np.float32
        assert np.array_equal(f, fo)
        assert np.array_equal(v.view(np.uint32), vo.view(np.uint32))
    if name != "gyroid":
        assert _closed_manifold(v, f)


@pytest.mark.parametrize("shape,seed,level", [((9, 17, 33), 0, 0.0), ((40, 7, 12), 1, 0.1), ((2, 2, 2), 2, 0.0),
                                              ((31, 32, 33), 3, -0.2), ((2, 50, 3), 4, 0.0)])
def test_noise_volumes_every_case_bit_exact(shape, seed, level):
    rng = np.random.default_rng(seed)
    vol = rng.uniform(-1, 1, shape).astype(np.float32)
    vol[0, 0, 0], vol[-1, -1, -1] = -1.0, 1.0
    v, f = _gpu_mc(vol, level, "ascent")
    vo, fo = mc_oracle.marching_cubes(vol, level, "ascent")
    assert np.array_equal(f, fo) and np.array_equal(v.view(np.uint32), vo.view(np.uint32))


def test_values_equal_to_level_and_ties():
    rng = np.random.default_rng(9)
    vol = rng.integers(-1, 2, (12, 13, 14)).astype(np.float32)  # many exact zeros and decider ties
    v, f = _gpu_mc(vol, 0.0, "descent")
    vo, fo = mc_oracle.marching_cubes(vol, 0.0, "descent")
    assert np.array_equal(f, fo) and np.array_equal(v.view(np.uint32), vo.view(np.uint32))


@pytest.mark.parametrize("level", [0.1, np.nextafter(np.float32(0.1), np.float32(1)).item(), float(np.float32(0.1)), -0.3, 1e-40])
def test_levels_between_float32_values(level):
    """The kernels classify float32 samples against the largest float32 <= level (exactly the double-precision test
    `v - level > 0` of the sweep): samples one ulp below / on / above a level that is not a float32 itself."""
    rng = np.random.default_rng(3)
    lf = np.float32(level)
    choices = np.array([lf, np.nextafter(lf, np.float32(-9)), np.nextafter(lf, np.float32(9)), lf - np.float32(1), lf + np.float32(1)], np.float32)
    vol = choices[rng.integers(0, 5, (9, 11, 40))]
    v, f = _gpu_mc(vol, level, "ascent")
    vo, fo = mc_oracle.marching_cubes(vol, level, "ascent")
    assert np.array_equal(f, fo) and np.array_equal(v.view(np.uint32), vo.view(np.uint32))


def test_errors_match_skimage_contract():
    from deepjoin_b200.core import errors
    vol = _fields(16)["sphere"].astype(np.float32)
    with pytest.raises(errors.IsosurfaceExtractionError):
        _gpu_mc(vol, 10.0, "ascent")
    with pytest.raises(errors.IsosurfaceExtractionError):
        _gpu_mc(np.zeros((5, 5, 5), np.float32), 0.0, "ascent")
    with pytest.raises(ValueError):
        _gpu_mc(vol, 0.0, "sideways")


def test_large_volume_256_against_oracle_and_manifold():
    vol = (_fields(256)["torus"]).astype(np.float32)
    v, f = _gpu_mc(vol, 0.0, "ascent")
    vo, fo = mc_oracle.marching_cubes(vol, 0.0, "ascent")
    assert np.array_equal(f, fo) and np.array_equal(v.view(np.uint32), vo.view(np.uint32))
    assert _closed_manifold(v, f)


def test_slab_sharding_reproduces_single_sweep():
    """Two slabs along axis 0 (one shared sample plane), seam references patched with the lower
    slab's top-plane table: identical to the single-volume result (BASELINE config 4)."""
    import ctypes
    from deepjoin_b200 import _lib
    rng = np.random.default_rng(4)
    vol = (_fields(40)["gyroid"] + 0.2 * rng.uniform(-1, 1, (40, 40, 40))).astype(np.float32)
    vo, fo = mc_oracle.marching_cubes(vol, 0.0, "ascent")
    cut = 23
    parts = []
    L = _lib.lib()
    for lo, hi, seam in ((0, cut, 0), (cut, 39, 1)):
        sub = torch.from_numpy(vol[lo:hi + 1].copy()).cuda()
        w = _lib.McWork(0)
        nv, nf = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(L.dj_mc_count(w.handle, sub.data_ptr(), sub.shape[0], 40, 40, lo, seam, 0.0, 0,
                                 ctypes.byref(nv), ctypes.byref(nf), None))
        v = torch.empty((nv.value, 3), dtype=torch.float32, device="cuda")
        f = torch.empty((nf.value, 3), dtype=torch.int32, device="cuda")
        _lib.check(L.dj_mc_fill(w.handle, 0, v.data_ptr(), f.data_ptr(), None))
        top = torch.empty(40 * 40 * 3, dtype=torch.int32, device="cuda")
        _lib.check(L.dj_mc_top_plane(w.handle, top.data_ptr(), None))
        torch.cuda.synchronize()
        parts.append((v.cpu().numpy(), f.cpu().numpy().astype(np.int64), top.cpu().numpy().astype(np.int64)))
    from deepjoin_b200.parallel import merge_slab_meshes
    v, f = merge_slab_meshes([(p[0], p[1]) for p in parts], [p[2] for p in parts])
    assert np.array_equal(f, fo) and np.array_equal(v.view(np.uint32), vo.view(np.uint32))


def test_random_volumes_levels_and_quantised_values_bit_exact():
    """Property test: any dims (incl. 2-sample axes and rows longer than a block), any level, noise or coarsely
    quantised values (many samples exactly ON the level, equal corner values, degenerate interpolation weights):
    faces and float32 vertices identical to the sequential sweep, or the same error."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None, derandomize=True, database=None)
    @given(seed=st.integers(0, 10 ** 6), n0=st.integers(2, 12), n1=st.integers(2, 12), n2=st.sampled_from([2, 3, 7, 33, 300]),
           level=st.sampled_from([0.0, 0.25, -0.5, 0.1234]), quant=st.sampled_from([0, 4, 2]), direction=st.sampled_from(["ascent", "descent"]))
    def run(seed, n0, n1, n2, level, quant, direction):
        rng = np.random.default_rng(seed)
        vol = rng.uniform(-1, 1, (n0, n1, n2)).astype(np.float32)
        if quant:
            vol = (np.round(vol * quant) / quant).astype(np.float32)
        try:
            vo, fo = mc_oracle.marching_cubes(vol, level, direction)
        except (RuntimeError, ValueError) as e:
            with pytest.raises(Exception) as ei:
                _gpu_mc(vol, level, direction)
            assert type(ei.value).__name__ in ("IsosurfaceExtractionError", type(e).__name__)
            return
        v, f = _gpu_mc(vol, level, direction)
        assert np.array_equal(f, fo) and np.array_equal(v.view(np.uint32), vo.view(np.uint32))

    run()
