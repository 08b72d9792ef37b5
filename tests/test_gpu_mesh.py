"""create_mesh drop-in: GPU grid query + GPU marching cubes + f64 post-processing against the
oracle pipeline run on the SAME f32 grid (topology is only comparable on identical grids)."""
import numpy as np
import pytest
import torch

from oracle import mc_oracle
from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("dims,use_net,level,direction,sigmoid",
                         [((48, 48, 48), 1, 0.0, "ascent", False), ((40, 32, 24), 0, 0.0, "ascent", False),
                          ((32, 32, 32), 2, 0.5, "descent", True)])
def test_create_mesh_equals_oracle_pipeline_on_the_same_grid(precision, dims, use_net, level, direction, sigmoid):
    from deepjoin_b200.core import reconstruct as R
    dec = util.make_decoder(0, "trained_norm", precision=precision)
    code = util.trained_code(0)
    mesh = R.create_mesh(code, dec, use_net, sigmoid=sigmoid, level=level, gradient_direction=direction, dims=dims,
                         padding=0.1)
    values = R.decode_shape(code, dec, dims, use_net=use_net, padding=0.1, reshape=False, sigmoid=sigmoid)
    values = values.reshape([d for d in dims])
    vo, fo = mc_oracle.mesh_from_volume(values, level, dims, 0.1, direction)
    from deepjoin_b200.mesh import Mesh
    ref = Mesh(vo, fo)
    assert np.array_equal(mesh.faces, ref.faces)
    assert np.array_equal(mesh.vertices, ref.vertices)
    assert mesh.vertices.dtype == np.float64 and len(mesh.faces) > 0


def test_create_mesh_raises_isosurface_error_and_worker_makes_empty_mesh(tmp_path):
    from deepjoin_b200.core import errors, reconstruct as R
    dec = util.make_decoder(0, "default", precision="fp32")  # degenerate field: no zero crossing
    code = util.code_for(0, "default")
    with pytest.raises(errors.IsosurfaceExtractionError):
        R.create_mesh(code, dec, 0, sigmoid=False, level=0.0, gradient_direction="ascent", dims=(16, 16, 16))
    with pytest.raises(errors.IsosurfaceExtractionError):
        R.create_mesh(code, dec, 0, sigmoid=False, level=0.0, gradient_direction="sideways", dims=(16, 16, 16))


def test_create_mesh_exports_and_saves_values(tmp_path):
    from deepjoin_b200.core import reconstruct as R
    from deepjoin_b200.mesh import load_ply
    dec = util.make_decoder(1, "trained_norm", precision="bf16")
    code = util.trained_code(1)
    f_out, vals = str(tmp_path / "m.ply"), str(tmp_path / "v.npy")
    mesh = R.create_mesh(code, dec, 1, sigmoid=False, level=0.0, gradient_direction="ascent", f_out=f_out,
                         save_values=vals, dims=(32, 32, 32))
    back = load_ply(f_out)
    assert np.array_equal(back.faces, mesh.faces)
    v = np.load(vals)
    assert v.shape == (32, 32, 32) and v.dtype == np.float64


def test_full_size_256_mesh_is_consistent():
    """BASELINE config 2: 256^3 grid + marching cubes; MC of the device grid == oracle MC of the
    same grid copied to the host."""
    from deepjoin_b200.core import reconstruct as R
    dec = util.make_decoder(0, "trained_norm", precision="bf16")
    code = util.trained_code(0)
    dims = (256, 256, 256)
    mesh = R.create_mesh(code, dec, 1, sigmoid=False, level=0.0, gradient_direction="ascent", dims=dims)
    vol = R.decode_shape_device(code, dec, dims, use_net=1, padding=0.1, sigmoid=False).reshape(dims).cpu().numpy()
    vo, fo = mc_oracle.mesh_from_volume(vol, 0.0, dims, 0.1, "ascent")
    from deepjoin_b200.mesh import Mesh
    ref = Mesh(vo, fo)
    assert np.array_equal(mesh.faces, ref.faces) and np.array_equal(mesh.vertices, ref.vertices)


@pytest.mark.parametrize("nv,dup_frac", [(1, 0.0), (1000, 0.0), (5000, 0.3), (200000, 0.05), (3000, 1.0)])
def test_device_vertex_merge_equals_host_merge(nv, dup_frac):
    """csrc/mesh_merge.cuh against Mesh.merge_vertices (np.round(.., 8) + first-occurrence order)."""
    from deepjoin_b200.core import reconstruct as R
    from deepjoin_b200.mesh import Mesh
    rng = np.random.default_rng(nv)
    v = rng.uniform(-0.6, 0.6, (nv, 3))
    ndup = int(nv * dup_frac)
    if ndup:
        src = rng.integers(0, nv, ndup)
        dst = rng.integers(0, nv, ndup)
        v[dst] = v[src] + rng.uniform(-4e-9, 4e-9, (ndup, 3))  # mostly the same key after rounding, sometimes not
    if dup_frac == 1.0:
        v[:] = v[0]
    f = rng.integers(0, nv, (2 * nv + 1, 3)).astype(np.int32)
    ref = Mesh(v.copy(), f.copy())
    gv, gf = R.merge_vertices(torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda())
    assert np.array_equal(gv.cpu().numpy(), ref.vertices)
    assert np.array_equal(gf.cpu().numpy().astype(np.int64), ref.faces)


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("direction", ["ascent", "descent"])
def test_slab_sharded_create_mesh_is_byte_identical(world, direction):
    """BASELINE config 4 on one GPU: the ranks' slab fragments computed one after the other, merged as the
    gather step does, finished as create_mesh does -> identical to the unsharded create_mesh."""
    from deepjoin_b200 import parallel as P
    from deepjoin_b200.core import reconstruct as R
    dec = util.make_decoder(0, "trained_norm", precision="bf16")
    code = util.trained_code(0)
    dims = (48, 40, 44)
    ref = R.create_mesh(code, dec, 1, sigmoid=False, level=0.0, gradient_direction=direction, dims=dims, padding=0.1)
    frags = []
    for k, (lo, hi) in enumerate(P.slab_ranges(dims[0], world)):
        v, f, top, _ = P.slab_fragment(code, dec, 1, False, 0.0, dims, 0.1, lo, hi, seam=k > 0 and lo > 0)
        frags.append((v.clone(), f.clone(), top.clone()))
    v, f = P.merge_fragments(frags)
    mesh = P.finish_mesh(v, f, dims, 0.1, direction)
    assert np.array_equal(mesh.faces, ref.faces) and np.array_equal(mesh.vertices, ref.vertices)
    # world size 1 (no process group): the collective entry point degenerates to create_mesh
    one = P.create_mesh_sharded(code, dec, 1, sigmoid=False, level=0.0, gradient_direction=direction, dims=dims, padding=0.1)
    assert np.array_equal(one.faces, ref.faces) and np.array_equal(one.vertices, ref.vertices)


def test_slab_sharded_create_mesh_raises_like_the_reference():
    from deepjoin_b200 import parallel as P
    from deepjoin_b200.core import errors
    dec = util.make_decoder(0, "default", precision="fp32")  # degenerate field: no zero crossing
    code = util.code_for(0, "default")
    with pytest.raises(errors.IsosurfaceExtractionError):
        P.create_mesh_sharded(code, dec, 0, sigmoid=False, level=0.0, gradient_direction="ascent", dims=(16, 16, 16))


def test_full_size_512_in_8_slabs_equals_unsharded():
    """BASELINE configs[3] at its full size on one GPU: 512^3 (134 M queries), the 8 ranks' slabs computed one after
    the other and merged as the gather step does == the unsharded create_mesh, byte for byte; every face index is in range and every vertex is used."""
    from deepjoin_b200 import parallel as P
    from deepjoin_b200.core import reconstruct as R
    dec = util.make_decoder(0, "trained_norm", precision="bf16")
    code = util.trained_code(0)
    dims = (512, 512, 512)
    ref = R.create_mesh(code, dec, 1, sigmoid=False, level=0.0, gradient_direction="ascent", dims=dims, padding=0.1)
    assert len(ref.vertices) > 1_000_000 and ref.faces.min() == 0 and ref.faces.max() == len(ref.vertices) - 1
    frags = []
    for k, (lo, hi) in enumerate(P.slab_ranges(dims[0], 8)):
        v, f, top, _ = P.slab_fragment(code, dec, 1, False, 0.0, dims, 0.1, lo, hi, seam=k > 0 and lo > 0)
        frags.append((v.clone(), f.clone(), top.clone()))
    v, f = P.merge_fragments(frags)
    mesh = P.finish_mesh(v, f, dims, 0.1, "ascent")
    assert np.array_equal(mesh.faces, ref.faces) and np.array_equal(mesh.vertices, ref.vertices)
    # every vertex is used, no face refers to one vertex twice unless the duplicate-vertex merge collapsed it
    fa = torch.from_numpy(ref.faces).cuda()
    assert torch.unique(fa).numel() == len(ref.vertices)
    degenerate = ((fa[:, 0] == fa[:, 1]) | (fa[:, 1] == fa[:, 2]) | (fa[:, 0] == fa[:, 2])).sum().item()
    assert degenerate < 1e-3 * len(ref.faces)
