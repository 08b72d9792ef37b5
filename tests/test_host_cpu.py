"""CPU-only checks of the host layer: C-ABI library loads and exports every declared symbol,
sampler index schedule is bit-exact with the reference's sampler, mesh container, grid
coordinate arithmetic, drop-in import shims."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol():
    from deepjoin_b200 import _lib, build
    build.build()
    header = open(os.path.join(ROOT, "include", "deepjoin_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    debug_block = re.search(r"#ifdef DJ_DEBUG(.*?)#endif", header, flags=re.S).group(1)
    assert set(re.findall(r"\b(dj_[a-z0-9_]+)\s*\(", debug_block)) == set(_lib.DEBUG_SIGNATURES)
    header = header.replace(debug_block, "")  # diagnostics are not part of the product library
    declared = set(re.findall(r"\b(dj_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    for name in _lib.DEBUG_SIGNATURES:
        assert not hasattr(L, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    # one ABI number everywhere: header #define, ctypes bindings, built library (and __graft_entry__.build(), below)
    abi = int(re.search(r"#define\s+DJ_ABI_VERSION\s+(\d+)", header).group(1))
    assert abi == _lib.ABI_VERSION
    assert _lib.lib().dj_abi_version() == abi


def test_graft_entry_build_runs_on_an_up_to_date_tree():
    """The driver's entry point (it asserted a stale ABI number and no test called it)."""
    sys.path.insert(0, ROOT)
    import __graft_entry__ as E
    E.build()


def test_status_codes_map_to_reference_exceptions():
    from deepjoin_b200 import _lib
    from deepjoin_b200.core import errors
    _lib.lib()
    with pytest.raises(errors.IsosurfaceExtractionError):
        _lib.check(_lib.ERR_NO_SURFACE)
    with pytest.raises(errors.IsosurfaceExtractionError):
        _lib.check(_lib.ERR_LEVEL_RANGE)
    with pytest.raises(errors.NoSamplesError):
        _lib.check(_lib.ERR_NO_SAMPLES)
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.ERR_UNSUPPORTED)
    with pytest.raises(RuntimeError):
        _lib.check(_lib.ERR_BAD_NET)
    assert issubclass(errors.IsosurfaceExtractionError, RuntimeError)


def test_tensor_sdf_helpers_known_answers():
    """core/reconstruct.py:33-42 (a8): occ = 1 where sdf <= 0 (zero counts as inside), udf = |sdf|, and the way back."""
    from deepjoin_b200.core import reconstruct as R
    sdf = torch.tensor([[-2.0], [-1e-30], [-0.0], [0.0], [1e-30], [3.5], [float("inf")], [-float("inf")]])
    occ = R.tensor_sdf_to_occ(sdf)
    assert occ.dtype == sdf.dtype and occ.reshape(-1).tolist() == [1.0, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 1.0]
    fin = sdf[:6]
    udf = R.tensor_sdf_to_udf(fin)
    assert torch.equal(udf.reshape(-1), torch.tensor([2.0, 1e-30, 0.0, 0.0, 1e-30, 3.5]))
    back = R.tensor_udf_to_sdf(udf, R.tensor_sdf_to_occ(fin))
    assert torch.equal(back.abs(), fin.abs()) and torch.equal(back[[0, 1, 4, 5]], fin[[0, 1, 4, 5]])
    soft = R.tensor_udf_to_sdf(torch.tensor([1.0, 1.0, 1.0]), torch.tensor([0.49, 0.51, 1.0]))  # occ is rounded first
    assert soft.tolist() == [1.0, -1.0, -1.0]


def test_null_arguments_are_rejected_without_a_gpu():
    from deepjoin_b200 import _lib
    L = _lib.lib()
    assert L.dj_query_points(None, None, None, 0, 0, 0, 0, 0, None, None) == _lib.ERR_INVALID
    assert L.dj_mc_fill(None, 0, None, None, None) == _lib.ERR_INVALID
    assert b"dj_mc_fill" in L.dj_last_error()


def test_sample_indices_match_reference_sampler():
    from deepjoin_b200.core import data
    g = np.load(os.path.join(G, "sampler.npz"))
    pts = np.hstack((g["xyz"], g["n"]))
    for seed, n, key in ((11, 600, "sel"), (12, 601, "sel2")):
        np.random.seed(seed)
        idx = data.select_sample_indices(g["sdf"], n, uniform_ratio=0.2)
        assert np.array_equal(pts[idx], g[key + "_pts"]) and np.array_equal(g["sdf"][idx], g[key + "_vals"])
        np.random.seed(seed)
        p, v = data.select_samples(pts, g["sdf"], n, uniform_ratio=0.2)
        assert np.array_equal(p, g[key + "_pts"]) and np.array_equal(v, g[key + "_vals"])


def test_partial_sample_indices_match_reference_sampler():
    """core.data.select_partial_samples (core/data.py:76-119) on row indices, against the reference's output."""
    from deepjoin_b200.core import data
    from deepjoin_b200 import reconstruct_deepsdf as RD
    g = np.load(os.path.join(G, "deepsdf_reconstruct.npz"))
    np.random.seed(21)
    idx = data.select_partial_sample_indices(g["sdf"], g["mask"], 500, uniform_ratio=0.2)
    assert g["mask"][idx].all()
    assert np.array_equal(g["xyz"][idx], g["sel_pts"]) and np.array_equal(g["sdf"][idx], g["sel_vals"])
    np.random.seed(21)
    p, v = data.select_partial_samples(g["xyz"], g["sdf"], g["mask"], 500, uniform_ratio=0.2)
    assert np.array_equal(p, g["sel_pts"]) and np.array_equal(v, g["sel_vals"])
    # the schedule of the whole loop consumes the RNG like the reference loop did (seed 5, 12 x 512)
    np.random.seed(5)
    sched = RD.make_schedule(g["sdf"], g["mask"], 12, 512, 0.2)
    assert np.array_equal(g["xyz"][sched].astype(np.float32), g["default_batch_pts"])


def test_library_permutation_continues_numpys_legacy_stream():
    """dj_host_mt19937_permutation == np.random.permutation, element for element, and leaves numpy's global
    generator in the state numpy itself would be in (mid-block positions, several twister refills, cached gaussian)."""
    from deepjoin_b200.core import data
    for seed, n in ((0, 65536), (1, 250000), (2, 312500), (3, 70001)):
        np.random.seed(seed)
        np.random.rand(seed * 7 + 3)
        np.random.randn()  # leaves a cached gaussian in the state
        a, ra, ga = np.random.permutation(n), np.random.randint(1 << 30), np.random.randn()
        np.random.seed(seed)
        np.random.rand(seed * 7 + 3)
        np.random.randn()
        b, rb, gb = data._permutation(n), np.random.randint(1 << 30), np.random.randn()
        assert b.dtype == a.dtype and np.array_equal(a, b) and ra == rb and ga == gb
    np.random.seed(5)
    c1 = np.random.choice(250000, size=62500, replace=False)
    np.random.seed(5)
    assert np.array_equal(c1, data._permutation(250000)[:62500])


@pytest.mark.parametrize("n_pts,iters,num_samples,ratio,seed", [
    (500000, 4, 30000, 0.2, 0),   # the reference's sizes: 250 k surface + 250 k uniform rows, 30,000 per batch
    (120001, 5, 8000, 0.5, 1),    # no ratio step, odd pool
    (90000, 3, 8001, 0.7, 2),     # surface rows sub-sampled, odd batch
    (4001, 5, 6000, 0.2, 3),      # fewer rows of a class than asked for: drawn with replacement
    (70000, 3, 1, 0.3, 4),        # one sample per batch (no positive window at all)
])
def test_native_schedule_is_the_reference_sampler_on_numpys_stream(n_pts, iters, num_samples, ratio, seed):
    """dj_host_sample_schedule (all iterations in one native call, threaded) == calling the reference-pinned Python
    sampler once per iteration: same rows, and numpy's global generator is left exactly where it would be."""
    import synth_data
    from deepjoin_b200 import reconstruct as R
    sdf = synth_data.make_sample_set(n_pts, seed=seed)[1].reshape(-1)
    np.random.seed(seed)
    np.random.randn()  # a cached gaussian and a mid-block position
    a = R.make_schedule(sdf, iters, num_samples, ratio)
    tail_a = (np.random.randint(1 << 30), np.random.randn())
    os.environ["DEEPJOIN_B200_PY_SAMPLER"] = "1"
    try:
        np.random.seed(seed)
        np.random.randn()
        b = R.make_schedule(sdf, iters, num_samples, ratio)
        tail_b = (np.random.randint(1 << 30), np.random.randn())
    finally:
        del os.environ["DEEPJOIN_B200_PY_SAMPLER"]
    assert a.dtype == np.int32 and np.array_equal(a, b) and tail_a == tail_b


def test_native_schedule_matches_the_reference_made_fixture():
    """... and directly against rows the reference's own select_samples produced (tests/golden/sampler.npz)."""
    from deepjoin_b200 import reconstruct as R
    g = np.load(os.path.join(G, "sampler.npz"))
    pts = np.hstack((g["xyz"], g["n"]))
    for seed, n, key in ((11, 600, "sel"), (12, 601, "sel2")):
        np.random.seed(seed)
        idx = R.make_schedule(g["sdf"], 1, n, 0.2)[0]
        assert np.array_equal(pts[idx], g[key + "_pts"]) and np.array_equal(g["sdf"][idx], g[key + "_vals"])


@pytest.mark.parametrize("iters,py_sampler", [(137, False), (6, False), (30, False), (23, True)])
def test_streamed_schedule_blocks_equal_one_call(iters, py_sampler):
    """reconstruct()'s producer (_schedule_blocks: one sampler call on a background thread, finished leading
    iterations handed out through the progress counter) yields exactly make_schedule's rows, in blocks that are
    multiples of the log interval except the last, and leaves numpy's generator where one call would."""
    import synth_data
    from deepjoin_b200 import reconstruct as R
    sdf = np.ascontiguousarray(synth_data.make_sample_set(60000, seed=5)[1].reshape(-1), dtype=np.float32)
    if py_sampler:
        os.environ["DEEPJOIN_B200_PY_SAMPLER"] = "1"
    try:
        np.random.seed(9)
        whole = R.make_schedule(sdf, iters, 4000, 0.2)
        tail_a = np.random.randint(1 << 30)
        np.random.seed(9)
        blocks = [b.copy() for b in R._schedule_blocks(sdf, iters, 4000, 0.2)]
        tail_b = np.random.randint(1 << 30)
    finally:
        os.environ.pop("DEEPJOIN_B200_PY_SAMPLER", None)
    assert all(b.shape[0] % 10 == 0 for b in blocks[:-1]) and all(0 < b.shape[0] <= 100 for b in blocks)
    assert np.array_equal(np.concatenate(blocks), whole) and tail_a == tail_b


def test_streamed_schedule_blocks_propagate_no_samples_error():
    from deepjoin_b200 import reconstruct as R
    from deepjoin_b200.core import errors
    with pytest.raises(errors.NoSamplesError):
        list(R._schedule_blocks(np.full(1000, 0.5, np.float32), 20, 100, 0.2))


def test_native_schedule_raises_no_samples_error():
    from deepjoin_b200 import reconstruct as R
    from deepjoin_b200.core import errors
    with pytest.raises(errors.NoSamplesError):
        R.make_schedule(np.full(1000, 0.5, np.float32), 2, 100, 0.2)
    with pytest.raises(errors.NoSamplesError):
        R.make_schedule(np.full(1000, -0.5, np.float32), 2, 100, 0.2)


@pytest.mark.parametrize("kind", ["default", "trained_norm"])
def test_surface_interior_schedule_reproduces_the_reference_run(kind):
    """sampling_method="surface_interior_only" (reconstruct.py:60-98): the host schedule, fed to the oracle's Adam loop,
    reproduces the log and the code of the reference's own seeded run (tests/golden/make_golden_partial.py)."""
    from deepjoin_b200 import reconstruct as R
    from oracle import decoder_oracle as O, synth
    g = np.load(os.path.join(G, "reconstruct_surface_interior.npz"))
    xyz, sdf = g["xyz"].astype(np.float32), g["sdf"].astype(np.float32)
    np.random.seed(7)
    torch.manual_seed(7)
    code0 = R.init_code(128, 64, 0.01)
    sched = R.make_schedule_surface_interior(g["sdf"], 12, 500)
    assert sched.shape == (12, 500) and (sdf.reshape(-1)[sched[:, :100]] < 0).all() and (sched[:, :100] >= 2000).all()
    assert (sched[:, 100:] < 2000).all()
    net = O.fold(synth.make_state_dict(0, kind))
    zeros = torch.zeros(500, 3)
    log, code = O.adam_loop(net, code0, lambda e: (torch.from_numpy(xyz[sched[e]]), torch.from_numpy(sdf[sched[e]]), zeros),
                            12, lr=5e-3, lambda1=0.0)
    assert np.abs(np.asarray(log["loss"]) - g[kind + "_log_loss"]).max() < 1e-4
    assert np.abs(code.numpy() - g[kind + "_code"]).max() < 2e-3


def test_sampler_raises_no_samples_error():
    from deepjoin_b200.core import data, errors
    with pytest.raises(errors.NoSamplesError):
        data.select_samples(np.random.rand(50, 6), np.ones((50, 1)), 10, uniform_ratio=None)


def test_schedule_consumes_rng_like_the_reference_loop():
    from deepjoin_b200 import reconstruct as R
    from oracle import decoder_oracle as O
    g = np.load(os.path.join(G, "sampler.npz"))
    pts = np.hstack((g["xyz"], g["n"]))
    np.random.seed(3)
    sched = R.make_schedule(g["sdf"], 3, 200, 0.2)
    np.random.seed(3)
    for e in range(3):
        p, v = O.select_samples(pts, g["sdf"], 200, uniform_ratio=0.2)
        assert np.array_equal(pts[sched[e]], p)


def test_init_code_matches_reference_rng_use():
    from deepjoin_b200 import reconstruct as R
    g = np.load(os.path.join(G, "reconstruct.npz"))
    torch.manual_seed(5)
    assert np.array_equal(R.init_code(128, 64, 0.01).numpy(), g["default_code0"])


def test_grid_axis_arithmetic_is_numpy_linspace():
    # the kernels evaluate (t == d-1 ? stop : t*step) - offs in float64 (mlp_tc.cuh lin_axis)
    for d in (2, 3, 12, 128, 256, 257, 512):
        for padding in (0.0, 0.1, 0.25):
            stop, offs = 1.0 + (padding * 2), 0.5 + padding
            step = stop / (d - 1)
            t = np.arange(d, dtype=np.float64)
            mine = np.where(t == d - 1, stop, t * step) - offs
            ref = np.linspace(0, 1.0 + (padding * 2), d) - (0.5 + padding)
            assert np.array_equal(mine, ref)


def test_mesh_container(tmp_path):
    from deepjoin_b200.mesh import Mesh, load_ply
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 0, 0], [np.nan, 0, 0]], dtype=np.float64)
    f = np.array([[0, 1, 2], [0, 3, 2], [0, 1, 4]])
    m = Mesh(v, f)
    assert len(m.vertices) == 3 and m.faces.tolist() == [[0, 1, 2], [0, 1, 2]]
    raw = Mesh(v[:4], f[:2], process=False)
    assert len(raw.vertices) == 4
    path = str(tmp_path / "m.ply")
    m.export(path)
    back = load_ply(path)
    assert np.allclose(back.vertices, m.vertices) and np.array_equal(back.faces, m.faces)
    m.export(str(tmp_path / "m.obj"))
    assert Mesh().is_empty
    Mesh().export(str(tmp_path / "e.ply"))


def test_decoder_rejects_unsupported_configurations():
    from deepjoin_b200.networks.decoder_z_lb_both_leaky_n_sdf import Decoder
    from tests.util import SPECS
    base = dict(latent_size=128, tool_latent_size=64, num_dims=3, **SPECS["NetworkSpecs"], **SPECS["SubnetSpecs"])
    for bad in (dict(use_tanh=True), dict(xyz_in_all=True), dict(latent_in=[2, 4]), dict(num_dims=2),
                dict(subnet_xyz=False), dict(dims=[256] * 8)):
        kw = dict(base)
        kw.update(bad)
        with pytest.raises(NotImplementedError):
            Decoder(**kw)
    dec = Decoder(**base)
    keys = set(dec.state_dict())
    assert {"lin0.weight_g", "lin0.weight_v", "lin0.bias", "lin8.weight", "hnet_lin5.weight", "hnet_lin4.weight_g"} <= keys
    assert dec.lin3.weight_v.shape == (381, 512) and dec.lin4.weight_v.shape == (512, 512)
    with pytest.raises(NotImplementedError):  # training mode has no kernel
        dec.train()._check_eval()
    with pytest.raises(RuntimeError):
        dec.eval().forward(torch.zeros(2, 195), use_net=7)
    with pytest.raises(NotImplementedError):
        dec.forward(torch.zeros(2, 195), return_normals=True)


def test_product_never_imports_the_oracle():
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "deepjoin_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "oracle/" in txt:
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_dropin_import_names():
    code = ("import core, core.reconstruct, core.errors, core.data, networks.decoder_z_lb_both_leaky_n_sdf as n;"
            "import deepjoin_b200.core as c;"
            "from reconstruct import reconstruct;"
            "assert core.errors.IsosurfaceExtractionError is c.errors.IsosurfaceExtractionError;"
            "assert core.reconstruct.create_mesh is c.reconstruct.create_mesh;"
            "a = __import__('networks.' + 'decoder_z_lb_both_leaky_n_sdf', fromlist=['Decoder']);"
            "assert a.Decoder is n.Decoder; print('ok')")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "deepjoin_b200", "dropin"))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


def test_bench_own_arm_does_not_touch_the_oracle():
    """bench.py may execute oracle/ only in its cpu_baseline / --impl reference legs: the decoder rate, the
    reference's latent iteration (B3) and the CPU marching-cubes sweep (B5)."""
    import ast
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef):
            for sub in ast.walk(node):
                names = []
                if isinstance(sub, ast.ImportFrom):
                    names = [sub.module or ""]
                elif isinstance(sub, ast.Import):
                    names = [a.name for a in sub.names]
                for nm in names:
                    if nm.split(".")[0] in ("oracle", "tests"):
                        assert node.name in ("cpu_decoder_rate", "cpu_latent_iteration", "cpu_marching_cubes_ms"), (node.name, nm)
    for node in tree.body:  # module level
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            nm = getattr(node, "module", None) or node.names[0].name
            assert nm.split(".")[0] not in ("oracle", "tests")
    # the synthetic-input module is oracle-free
    sd = open(os.path.join(ROOT, "synth_data.py")).read()
    assert not re.search(r"^\s*(from|import)\s+(oracle|tests)\b", sd, flags=re.M)


def test_synthetic_weights_do_not_need_the_oracle():
    """synth_data (committed head shifts) and oracle.synth (fp64 oracle centring) build identical weights."""
    import torch
    import synth_data
    from oracle import synth
    a = synth_data.make_state_dict(3, "trained_norm")
    shifts = synth.compute_head_shifts({k: v.clone() for k, v in synth_data.make_state_dict(3, "default").items()}, synth.MUGS, 3)
    assert len(shifts) == 2
    b = synth.make_state_dict(3, "trained_norm")
    assert all(torch.equal(a[k], b[k]) for k in a)
    c = synth.make_state_dict(11, "trained_norm")  # no committed shift -> oracle centring
    assert set(c) == set(a)


def test_deepsdf_dropin_surface():
    """the baseline decoder's import names and its rejection of unsupported configurations."""
    from deepjoin_b200.networks.decoder_deepsdf import Decoder
    from deepjoin_b200.core import utils_deepsdf
    ok = dict(dims=[512] * 8, dropout=list(range(8)), dropout_prob=0.2, norm_layers=list(range(8)), latent_in=[4],
              xyz_in_all=False, use_tanh=False, latent_dropout=False, weight_norm=True)
    dec = Decoder(128, **ok)
    import synth_data
    dec.load_state_dict(synth_data.make_deepsdf_state_dict(0, "default", prefix="module."))
    assert sorted(k for k in dec.state_dict() if k.startswith("lin8")) == ["lin8.bias", "lin8.weight"]
    assert tuple(dec.state_dict()["lin8.weight"].shape) == (1, 512)
    for bad in (dict(use_tanh=True), dict(xyz_in_all=True), dict(latent_in=[2, 4]), dict(dims=[256] * 8), dict(latent_dropout=True)):
        with pytest.raises(NotImplementedError):
            Decoder(128, **dict(ok, **bad))
    assert all(hasattr(utils_deepsdf, n) for n in ("decode_sdf", "create_mesh", "values2mesh"))
    with pytest.raises(RuntimeError):
        dec.eval().pack()  # CPU decoder: no fallback


def test_reconstruction_handler_names_and_round_trips(tmp_path):
    """the result tree and file names of the reference's ReconstructionHandler (core/handler.py:787-860,
    core/workspace.py:113-203) and load / save of codes, loss logs, value grids and meshes."""
    from deepjoin_b200.core import errors, handler, workspace
    from deepjoin_b200.mesh import Mesh
    exp = str(tmp_path / "mugs")
    os.makedirs(exp)
    h = handler.ReconstructionHandler(exp, signiture=[800, 0.005, 30000], name="ours", checkpoint="2000",
                                      lat_signiture=[800, 0.005])
    base = os.path.join(exp, "Reconstructions", "ours@2000")
    assert h.path_code(7) == os.path.join(base, "Codes", "7_800_0.005_.pth")
    assert h.path_loss(7) == os.path.join(base, "Codes", "7_800_0.005___loss.pth")
    assert h.path_mesh(7, 2) == os.path.join(base, "Meshes", "7_2_800_0.005_30000_.ply")
    assert h.path_values(7, 2) == os.path.join(base, "Meshes", "7_2_800_0.005_30000__values.npy")
    assert h.path_eval(7, 2, 0, "chamfer") == os.path.join(base, "Codes", "7_2_0_chamfer_800_0.005_30000_.npy")
    assert not os.path.isdir(base)  # nothing is created until asked to
    assert workspace.get_reconstructions_dir(exp, None, "latest") == os.path.join(exp, "Reconstructions", "latest")
    assert handler.iterify_path("a/b.pth", 3) == "a/b__iter-3.pth" and handler.lossify_log_path("a/b.pth") == "a/b__loss.pth"
    code = torch.arange(192, dtype=torch.float32).reshape(1, 192)
    h.set_code(7, code, save=True)
    assert os.path.isfile(h.path_code(7))
    fresh = handler.ReconstructionHandler(exp, [800, 0.005, 30000], "ours", "2000", lat_signiture=[800, 0.005])
    assert torch.equal(fresh.get_code(7), code)
    with pytest.raises(FileNotFoundError):
        fresh.get_code(8)
    vals = np.random.default_rng(0).random((4, 4, 4))
    h.set_values(7, 2, vals, save=True)
    assert np.array_equal(fresh.get_values(7, 2), vals)
    mesh = Mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]]), np.array([[0, 1, 2], [0, 2, 3]]), process=False)
    h.set_mesh(mesh, 7, 2, save=True)
    back = fresh.get_mesh(7, 2)
    assert np.array_equal(back.faces, mesh.faces) and np.allclose(back.vertices, mesh.vertices)
    with pytest.raises(errors.IsosurfaceExtractionError):
        fresh.get_mesh(7, 3)
    obj = str(tmp_path / "m.obj")
    handler.saver(obj, mesh)
    again = handler.loader(obj)
    assert np.array_equal(again.faces, mesh.faces) and np.allclose(again.vertices, mesh.vertices)
    with pytest.raises(RuntimeError, match="Unhandled file type"):
        handler.loader(str(tmp_path / "x.gif"))
    json.dump({"NetworkArch": "decoder_deepsdf"}, open(os.path.join(exp, "specs.json"), "w"))
    assert workspace.load_experiment_specifications(exp)["NetworkArch"] == "decoder_deepsdf"


def test_native_mesh_writers_produce_the_python_writers_bytes(tmp_path):
    """dj_host_write_ply / dj_host_write_obj (what Mesh.export uses) == the numpy / Python writers, byte for byte;
    empty meshes; an unwritable path is a PermissionError / OSError as open() would raise."""
    import filecmp
    from deepjoin_b200.mesh import Mesh, load_obj, load_ply
    rng = np.random.default_rng(3)
    m = Mesh(rng.normal(size=(2500, 3)) * 10.0 ** rng.integers(-6, 3, size=(2500, 1)), rng.integers(0, 2500, (4000, 3)), process=False)
    for ext, ref in (("ply", m._export_ply_numpy), ("obj", m._export_obj_python)):
        a, b = str(tmp_path / ("a." + ext)), str(tmp_path / ("b." + ext))
        m.export(a)
        ref(b)
        assert filecmp.cmp(a, b, shallow=False), ext
    e = str(tmp_path / "empty.ply")
    Mesh().export(e)
    assert load_ply(e).is_empty
    Mesh().export(str(tmp_path / "empty.obj"))
    assert load_obj(str(tmp_path / "empty.obj")).is_empty
    with pytest.raises(OSError):  # PermissionError is an OSError
        m.export(str(tmp_path / "no_such_dir" / "x.ply"))


def test_worker_loop_overlaps_the_deferred_part_and_reports_progress_in_order(tmp_path):
    """core.utils_multiprocessing._each: the callable an item returns runs in the background while the next item
    starts; at most one in flight; the callback fires once per item, after the item is really finished; finished
    items are skipped; a repeated path waits for its first write."""
    import threading
    import time
    from deepjoin_b200.core import utils_multiprocessing as U
    log, lock = [], threading.Lock()

    def note(x):
        with lock:
            log.append(x)

    paths = [str(tmp_path / n) for n in ("a", "b", "done", "c", "c")]
    open(paths[2], "w").write("x")
    os.utime(paths[2], (2e9, 2e9))  # newer than the reference's overwrite stamp: a finished item

    def work(item, path):
        note(("gpu", item))

        def rest():
            time.sleep(0.05)
            open(path, "w").write("y")
            note(("saved", item))
        return rest

    U._each(list(range(5)), paths, False, lambda: note(("cb", None)), work)
    kinds = [k for k, _ in log]
    assert kinds.count("cb") == 5 and [i for k, i in log if k == "gpu"] == [0, 1, 3]   # item 2 finished before, item 4 = same path as 3
    assert log.index(("gpu", 1)) < log.index(("saved", 0))                               # overlap: next item started first
    cbs = [n for n, (k, _) in enumerate(log) if k == "cb"]
    assert cbs[0] > log.index(("saved", 0))         # progress is reported once an item has really been written
    assert cbs[1] > log.index(("saved", 1))
    assert cbs[3] > log.index(("saved", 3)) and kinds[-1] == "cb"
    assert all(os.path.exists(p) for p in paths)
