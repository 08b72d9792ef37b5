"""GPU parity of the decoder forward (fp32 CUDA-core path, split-operand and single-pass tcgen05 paths) against
the oracle restatement and the golden vectors produced by the reference itself.

Tolerances (BASELINE.json north_star): SDF 1e-3, normals / occupancy probability 1e-2.  The split-operand
tensor-core mode ("fp16x2", the default precision) and the fp32 path are held 10x tighter than that on BOTH weight
sets; single-pass bf16 meets it only on default-init weights and is held to its measured budget on trained-norm
weights (it is the opt-in throughput mode)."""
import os

import numpy as np
import pytest
import torch

from oracle import decoder_oracle as O
from oracle import synth
from tests import util

pytestmark = pytest.mark.gpu
G = util.GOLDEN
NAMES = ("c_occ", "b_occ", "r_occ", "t_occ", "c_sdf", "b_sdf", "r_sdf", "t_sdf", "c_n", "b_n", "r_n", "t_n")
SDF_TOL = {"fp32": 2e-5, "fp16x2": 2e-5, "bf16": 1e-3}
PROB_TOL = {"fp32": 2e-5, "fp16x2": 2e-5, "bf16": 1e-2}
NORTH_STAR_SDF, NORTH_STAR_PROB = 1e-3, 1e-2


def _tols(precision, kind):
    # trained-norm weights amplify rounding: bf16 is reported, fp32 / fp16x2 stay tight
    if kind == "trained_norm" and precision in ("fp32", "fp16x2"):
        return 2e-4, 2e-4
    return SDF_TOL[precision], PROB_TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "fp16x2", "bf16"])
@pytest.mark.parametrize("seed", [0, 1])
def test_default_weights_all_branches_vs_golden(precision, seed):
    g = np.load(os.path.join(G, "decoder_forward.npz"))
    tag = "default_s%d" % seed
    dec = util.make_decoder(seed, "default", precision=precision)
    code = torch.from_numpy(g[tag + "_code"]).cuda()
    pts = torch.from_numpy(g[tag + "_pts"]).cuda()
    inp = torch.cat([code.expand(pts.shape[0], -1), pts], 1)
    st, pt = _tols(precision, "default")
    outs = dec(inp)
    assert len(outs) == 12
    for nm, o in zip(NAMES, outs):
        ref = g[tag + "_all_" + nm]
        assert tuple(o.shape) == ref.shape, nm
        tol = st if "sdf" in nm else pt
        assert np.abs(o.cpu().numpy() - ref).max() <= tol, (nm, np.abs(o.cpu().numpy() - ref).max())
    for u in (0, 1, 2, 3):
        o = dec(inp, use_net=u).cpu().numpy()
        assert o.shape == (pts.shape[0], 1)
        assert np.abs(o - g[tag + "_u%d" % u]).max() <= st, u
        o = dec(inp, use_net=u, return_occ=True).cpu().numpy()
        assert np.abs(o - g[tag + "_u%d_occ" % u]).max() <= pt, u


@pytest.mark.parametrize("precision", ["fp32", "fp16x2"])
@pytest.mark.parametrize("seed", [0, 1])
def test_trained_norm_weights_vs_golden(seed, precision):
    """The reference's own outputs (fp32 torch, tests/golden/make_golden.py) on weights with the trained norms."""
    g = np.load(os.path.join(G, "decoder_forward.npz"))
    tag = "trained_norm_s%d" % seed
    dec = util.make_decoder(seed, "trained_norm", precision=precision)
    code = torch.from_numpy(g[tag + "_code"]).cuda()
    pts = torch.from_numpy(g[tag + "_pts"]).cuda()
    inp = torch.cat([code.expand(pts.shape[0], -1), pts], 1)
    for nm, o in zip(NAMES, dec(inp)):
        ref = g[tag + "_all_" + nm]
        scale = max(1.0, float(np.abs(ref).max()))
        assert np.abs(o.cpu().numpy() - ref).max() <= 2e-4 * scale, nm
    for u in (0, 1, 2, 3):
        assert np.abs(dec(inp, use_net=u).cpu().numpy() - g[tag + "_u%d" % u]).max() <= 2e-4


@pytest.mark.parametrize("seed", [0, 1])
def test_trained_norm_weights_split_operand_parity(seed):
    """The parity mode on realistic weights: 20,000 points x every use_net x the 12-tuple against the
    fp64 oracle.  SDF within 1e-4 (north star 1e-3), normals / probabilities within 1e-3 (north star 1e-2), and no
    sign flip outside the +-2e-5 band in which the reference's own fp32 arithmetic is undecided (its error against
    fp64 on these inputs is 1.5e-5, see the fp32 rows of tools/x2_check.py)."""
    dec = util.make_decoder(seed, "trained_norm", precision="fp16x2")
    net = util.oracle_net(seed, "trained_norm", "float64")
    code = util.trained_code(seed)
    pts = synth.make_points(20000, seed)
    total_flips = in_band = 0
    for u in (0, 1, 2, 3):
        ref = O.forward(net, code.double(), pts.double(), u).numpy()
        out = dec.query(code.cuda(), pts.cuda(), use_net=u).cpu().numpy().astype(np.float64)
        err = np.abs(out - ref)
        flip = (out > 0) != (ref > 0)
        total_flips += int(flip.sum())
        in_band += int((np.abs(ref) <= 2e-5).sum())
        print("fp16x2 trained-norm seed %d use_net=%d: max %.3e mean %.3e sign flips %d (|ref| at the flips: %s)"
              % (seed, u, err.max(), err.mean(), int(flip.sum()), np.abs(ref[flip]).tolist()))
        assert err.max() <= 1e-4 <= NORTH_STAR_SDF
        assert not (flip & (np.abs(ref) > 2e-5)).any()
        occ_ref = O.forward(net, code.double(), pts.double(), u, return_occ=True).numpy()
        occ = dec.query(code.cuda(), pts.cuda(), use_net=u, return_occ=True).cpu().numpy()
        scale = max(1.0, float(np.abs(occ_ref).max()))  # use_net 0 / 3 return logits
        assert np.abs(occ - occ_ref).max() <= 1e-4 * scale
    assert total_flips <= in_band  # every flip sits on a value the fp32 reference cannot sign either
    ref12 = O.forward(net, code.double(), pts.double(), None)
    out12 = dec.query_all(code.cuda(), pts.cuda())
    for nm, a, b in zip(NAMES, out12, ref12):
        e = float((a.cpu().double() - b).abs().max())
        scale = max(1.0, float(b.abs().max()))
        assert e <= (1e-4 if "sdf" in nm else 1e-3) * scale, (nm, e)
        assert e <= (NORTH_STAR_SDF if "sdf" in nm else NORTH_STAR_PROB) * scale


def test_trained_norm_weights_bf16_error_budget():
    """Single-pass bf16 (the opt-in throughput mode) with trained-norm weights: the network amplifies operand rounding
    (max 5.7e-2, mean 2.8e-3).  Held to its measured budget, NOT claimed to meet 1e-3: sign flips are
    counted and bounded."""
    dec = util.make_decoder(0, "trained_norm", precision="bf16")
    net = util.oracle_net(0, "trained_norm", "float64")
    code = util.trained_code(0)
    pts = synth.make_points(20000, 3)
    ref = O.forward(net, code.double(), pts.double(), 1).numpy()
    out = dec.query(code.cuda(), pts.cuda(), use_net=1).cpu().numpy()
    err = np.abs(out - ref)
    flips = int(((out > 0) != (ref > 0)).sum())
    print("bf16 trained-norm use_net=1: max %.3e mean %.3e p99 %.3e sign flips %d / %d"
          % (err.max(), err.mean(), np.quantile(err, 0.99), flips, len(ref)))
    assert err.mean() < 1e-2 and err.max() < 0.1
    assert flips <= 400  # ~0.6 % of the points (measured 128): the iso-surface moves, which is why bf16 is not the default
    out32 = util.make_decoder(0, "trained_norm", precision="fp32").query(code.cuda(), pts.cuda(), use_net=1).cpu().numpy()
    assert np.abs(out32 - ref).max() < 2e-4


@pytest.mark.parametrize("precision", ["fp32", "fp16x2", "bf16"])
@pytest.mark.parametrize("n", [1, 63, 64, 65, 127, 128, 129, 1000, 70001])
def test_ragged_sizes(precision, n):
    dec = util.make_decoder(0, "default", precision=precision)
    net = util.oracle_net(0, "default")
    code = synth.make_code(0)
    pts = synth.make_points(n, 7)
    for u in (0, 3, 2):
        ref = O.forward(net, code, pts, u).numpy()
        out = dec.query(code.cuda(), pts.cuda(), use_net=u).cpu().numpy()
        assert out.shape == ref.shape
        assert np.abs(out - ref).max() <= SDF_TOL[precision]


def test_empty_input():
    dec = util.make_decoder(0, "default")
    out = dec.query(synth.make_code(0).cuda(), torch.zeros(0, 3).cuda(), use_net=1)
    assert tuple(out.shape) == (0, 1)


def test_error_contract():
    dec = util.make_decoder(0, "default")
    inp = torch.zeros(4, 195).cuda()
    with pytest.raises(RuntimeError, match="non-existent network"):
        dec(inp, use_net=4)
    with pytest.raises(NotImplementedError):
        dec(inp, return_normals=True)
    mixed = inp.clone()
    mixed[1, 0] = 1.0
    with pytest.raises(NotImplementedError):
        dec(mixed, use_net=0)


@pytest.mark.parametrize("precision", ["fp32", "fp16x2", "bf16"])
@pytest.mark.parametrize("dims", [(12, 12, 12), (6, 9, 11)])
def test_grid_query_matches_reference_decode_shape(precision, dims):
    from deepjoin_b200.core import reconstruct as R
    g = np.load(os.path.join(G, "grid.npz"))
    t = "d%dx%dx%d" % dims
    dec = util.make_decoder(0, "trained_norm", precision=precision)
    code = torch.from_numpy(g["code"])
    tol = {"fp32": 2e-4, "fp16x2": 2e-4, "bf16": 0.1}[precision]  # bf16: measured budget on trained-norm weights
    for u in (0, 1, 2, 3):
        v = R.decode_shape(code, dec, dims, use_net=u, batch_size=500, padding=0.1, reshape=False, sigmoid=False)
        ref = g[t + "_u%d_flat" % u]
        assert v.dtype == np.float64 and v.shape == ref.shape
        assert np.abs(v - ref).max() <= tol
    v = R.decode_shape(code, dec, dims, use_net=1, padding=0.1, reshape=True, sigmoid=True)
    assert v.shape == g[t + "_u1_reshape_sig"].shape
    assert np.abs(v - g[t + "_u1_reshape_sig"]).max() <= tol


@pytest.mark.parametrize("precision", ["fp32", "fp16x2"])
@pytest.mark.parametrize("dims", [(1, 9, 7), (8, 1, 5), (4, 6, 1), (1, 1, 1)])
def test_grid_query_with_an_axis_of_length_one(precision, dims):
    """np.linspace(0, stop, 1) is [0.0] (not [stop]): a 2-D slice query must evaluate the plane the reference does
    (fixture made by the reference itself, tests/golden/make_golden_grid_extra.py)."""
    from deepjoin_b200.core import reconstruct as R
    g = np.load(os.path.join(G, "grid_extra.npz"))
    t = "d%dx%dx%d" % dims
    dec = util.make_decoder(0, "trained_norm", precision=precision)
    code = torch.from_numpy(g["code"])
    assert np.array_equal(R.get_query_points(dims, 0.1), g[t + "_pts"])
    for u in (1, 3):
        v = R.decode_shape(code, dec, dims, use_net=u, batch_size=500, padding=0.1, reshape=False, sigmoid=False)
        assert v.shape == g[t + "_u%d_flat" % u].shape
        assert np.abs(v - g[t + "_u%d_flat" % u]).max() <= 2e-4


@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16x2"])
def test_grid_points_are_bit_identical_to_numpy_meshgrid(precision):
    """In-kernel coordinates == np.meshgrid/linspace coordinates: the same kernel fed the numpy
    points must return bit-identical values."""
    from deepjoin_b200.core import reconstruct as R
    dec = util.make_decoder(1, "trained_norm", precision=precision)
    code = util.trained_code(1)
    for dims, padding in (((17, 13, 9), 0.1), ((32, 32, 32), 0.0)):
        pts = torch.from_numpy(R.get_query_points(dims, padding)).float().cuda()
        a = dec.query(code.cuda(), pts, use_net=1).reshape(-1)
        b = R.decode_shape_device(code, dec, dims, use_net=1, padding=padding, sigmoid=False)
        assert torch.equal(a, b)
        lo, hi = 100, 1555
        c = R.decode_shape_device(code, dec, dims, use_net=1, padding=padding, sigmoid=False, rows=(lo, hi))
        assert torch.equal(c, b[lo:hi])


@pytest.mark.parametrize("precision", ["fp32", "fp16x2", "bf16"])
def test_decode_samples_host_buffers(precision):
    from deepjoin_b200.core import reconstruct as R
    dec = util.make_decoder(0, "default", precision=precision)
    net = util.oracle_net(0, "default")
    code = synth.make_code(0)
    pts = synth.make_points(5000, 2).numpy().astype(np.float64)
    for u, sig in ((1, True), (0, False)):
        ref = O.decode_samples(net, code, pts, use_net=u, sigmoid=sig)
        out = R.decode_samples(code, dec, pts, use_net=u, batch_size=2 ** 10, sigmoid=sig)
        assert out.dtype == np.float64 and out.shape == ref.shape
        assert np.abs(out - ref).max() <= SDF_TOL[precision]
    # point layouts: rows, the three coordinate arrays behind get_query_points' np.vstack(...).T (passed as they
    # lie), float32 input, an arbitrary strided view; several pipeline chunks (2^21 points each) with a ragged tail
    n = (1 << 22) + 12345
    big = np.vstack([c for c in (synth.make_points(n, 5).numpy().astype(np.float64).T)]).T
    assert big.strides == (8, 8 * n)
    a = R.decode_samples(code, dec, big, use_net=1, sigmoid=False)
    b = R.decode_samples(code, dec, np.ascontiguousarray(big), use_net=1, sigmoid=False)
    assert np.array_equal(a, b)
    sub = big[::7]
    c = R.decode_samples(code, dec, sub, use_net=1, sigmoid=False)
    assert np.array_equal(c, a[::7])
    d = R.decode_samples(code, dec, big[:1000].astype(np.float32), use_net=1, sigmoid=False)
    assert np.array_equal(d, a[:1000])


def test_large_grid_properties_256():
    """BASELINE config 2 size (256^3): size-independent properties instead of the oracle --
    slab queries tile the full query bit-exactly, B = max(C, -T) and R = max(C, T)."""
    from deepjoin_b200.core import reconstruct as R
    _large_grid_properties("bf16")


def test_large_grid_properties_256_split_operand():
    _large_grid_properties("fp16x2")


def _large_grid_properties(precision):
    from deepjoin_b200.core import reconstruct as R
    dec = util.make_decoder(0, "trained_norm", precision=precision)
    code = util.trained_code(0)
    dims = (256, 256, 256)
    tot = 256 ** 3
    b = R.decode_shape_device(code, dec, dims, use_net=1, padding=0.1, sigmoid=False)
    c = R.decode_shape_device(code, dec, dims, use_net=0, padding=0.1, sigmoid=False)
    t = R.decode_shape_device(code, dec, dims, use_net=3, padding=0.1, sigmoid=False)
    r = R.decode_shape_device(code, dec, dims, use_net=2, padding=0.1, sigmoid=False)
    assert torch.isfinite(b).all()
    assert torch.equal(b, torch.maximum(c, -t)) and torch.equal(r, torch.maximum(c, t))
    lo, hi = tot // 3 + 17, tot // 3 + 17 + 3_000_001
    s = R.decode_shape_device(code, dec, dims, use_net=1, padding=0.1, sigmoid=False, rows=(lo, hi))
    assert torch.equal(s, b[lo:hi])
    assert (b.min() < 0) and (b.max() > 0)


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("fp16x2", 2e-5), ("bf16", 1e-3)])
def test_latent_size_256_variant(precision, tol):
    """BASELINE configs[4]: CodeLength = 256 (fnet lin0 [512,259], lin3 [253,512]); every use_net branch
    against the oracle restatement."""
    from deepjoin_b200.networks.decoder_z_lb_both_leaky_n_sdf import Decoder
    arch = dict(latent_size=256, tool_latent_size=64)
    sd = synth.make_state_dict(2, "default", **arch)
    dec = Decoder(latent_size=256, tool_latent_size=64, num_dims=3, do_code_regularization=True,
                  **util.SPECS["NetworkSpecs"], **util.SPECS["SubnetSpecs"])
    dec.load_state_dict(sd)
    dec.precision = precision
    dec = dec.cuda().eval()
    net = O.fold(sd, **arch)
    code = synth.make_code(2, 256, 64)
    pts = synth.make_points(3000, 5)
    for u in (0, 1, 2, 3):
        ref = O.forward(net, code, pts, u).numpy()
        got = dec.query(code.cuda(), pts.cuda(), use_net=u).cpu().numpy()
        assert np.abs(got - ref).max() <= tol, (u, np.abs(got - ref).max())


def test_fp16_operand_mode_is_finer_than_bf16_at_the_same_kernel():
    """DJ_PREC_FP16: IEEE-half operands on the pair kernel.  All branches within 2e-5 on random-init weights;
    with trained-norm weights the SDF error is several times below the bf16 one (mean <= 2.5e-3)."""
    code = util.trained_code(0)
    pts = synth.make_points(8192, 3)
    for kind, tol in (("default", 3e-5), ("trained_norm", 2e-2)):
        net = util.oracle_net(0, kind, "float64")
        c = util.code_for(0, kind)
        d16 = util.make_decoder(0, kind, precision="fp16")
        dbf = util.make_decoder(0, kind, precision="bf16")
        for u in (0, 1, 2, 3):
            ref = O.forward(net, c.double(), pts.double(), u).numpy()
            e16 = np.abs(d16.query(c.cuda(), pts.cuda(), use_net=u).cpu().numpy() - ref)
            ebf = np.abs(dbf.query(c.cuda(), pts.cuda(), use_net=u).cpu().numpy() - ref)
            assert e16.max() <= tol, (kind, u, e16.max())
            assert e16.mean() < 0.5 * ebf.mean(), (kind, u, e16.mean(), ebf.mean())
            if kind == "trained_norm":
                assert e16.mean() <= 2.5e-3
    # grid mode + meshing go through the same precision switch
    from deepjoin_b200.core import reconstruct as R
    mesh = R.create_mesh(code, d16, 1, sigmoid=False, level=0.0, gradient_direction="ascent", dims=(32, 32, 32))
    assert len(mesh.faces) > 0


def test_grid_mode_equals_point_mode_for_any_dims_rows_and_branch():
    """Property test: the grid evaluator (coordinates made in-kernel) and the point evaluator fed numpy's meshgrid
    points return bit-identical values for any (non-cubic) dims, padding, row range, use_net and precision."""
    from hypothesis import given, settings, strategies as st
    from deepjoin_b200.core import reconstruct as R
    decs = {p: util.make_decoder(1, "trained_norm", precision=p) for p in ("bf16", "fp16", "fp16x2", "fp32")}
    code = util.trained_code(1).cuda()

    @settings(max_examples=40, deadline=None, derandomize=True, database=None)
    @given(d0=st.integers(2, 24), d1=st.integers(2, 24), d2=st.integers(2, 24), padding=st.sampled_from([0.0, 0.05, 0.1, 0.3]),
           u=st.sampled_from([0, 1, 2, 3]), precision=st.sampled_from(["bf16", "fp16", "fp16x2", "fp32"]), cut=st.tuples(st.floats(0, 1), st.floats(0, 1)))
    def run(d0, d1, d2, padding, u, precision, cut):
        dec = decs[precision]
        dims = (d0, d1, d2)
        tot = d0 * d1 * d2
        lo, hi = sorted(int(c * tot) for c in cut)
        hi = max(hi, lo + 1)
        pts = torch.from_numpy(R.get_query_points(dims, padding)).float().cuda()
        a = dec.query(code, pts[lo:hi].contiguous(), use_net=u).reshape(-1)
        b = R.decode_shape_device(code, dec, dims, use_net=u, padding=padding, sigmoid=False, rows=(lo, hi))
        assert a.shape == b.shape and torch.equal(a, b)

    run()


def test_four_pass_program_agrees_with_three_pass(tmp_path):
    """DJ_PREC_FP16X2 launches the three-pass program; the four-pass one (DEEPJOIN_B200_X2_FOURPASS=1, read once per
    process, and the fallback for layer shapes the three-pass builder does not cover) must give the same field: both
    within 1e-4 of the fp64 oracle, within 6e-5 of each other."""
    import subprocess
    import sys
    if os.environ.get("DEEPJOIN_B200_X2_FOURPASS"):
        pytest.skip("this process already runs the four-pass program")
    out_file = str(tmp_path / "four.npy")
    prog = ("import sys, numpy as np, torch; sys.path.insert(0, %r)\n"
            "from oracle import synth\nfrom tests import util\n"
            "dec = util.make_decoder(0, 'trained_norm', precision='fp16x2')\n"
            "out = dec.query(util.trained_code(0).cuda(), synth.make_points(5000, 4).cuda(), use_net=1)\n"
            "np.save(%r, out.cpu().numpy())\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), out_file)
    env = dict(os.environ, DEEPJOIN_B200_X2_FOURPASS="1")
    subprocess.run([sys.executable, "-c", prog], check=True, env=env, timeout=300)
    four = np.load(out_file).astype(np.float64)
    dec = util.make_decoder(0, "trained_norm", precision="fp16x2")
    pts = synth.make_points(5000, 4)
    three = dec.query(util.trained_code(0).cuda(), pts.cuda(), use_net=1).cpu().numpy().astype(np.float64)
    ref = O.forward(util.oracle_net(0, "trained_norm", "float64"), util.trained_code(0).double(), pts.double(), 1).numpy()
    assert np.abs(three - ref).max() <= 1e-4
    assert np.abs(four - ref).max() <= 1e-4
    assert np.abs(three - four).max() <= 6e-5
    assert not np.array_equal(three, four)  # the environment switch did select another program


def test_whole_grid_split_operand_vs_fp32_path():
    """Every tile and pass of the persistent split-operand kernel (129^3 = 2.1 M points: ~230 passes per CTA, a ragged
    last tile), against the fp32 CUDA-core path on the same grid: within 1e-4 everywhere, sign differences only where
    the fp32 value is itself within 5e-5 of zero.  (256^3: max 2.9e-5, 62 sign differences below 2e-5 --
    tools/full_grid_check.py.)"""
    from deepjoin_b200.core import reconstruct as R
    code = util.trained_code(0).cuda()
    dims = (129, 129, 129)
    vols = {}
    for prec in ("fp32", "fp16x2"):
        dec = util.make_decoder(0, "trained_norm", precision=prec)
        vols[prec] = R.decode_shape_device(code, dec, dims, use_net=1, padding=0.1, sigmoid=False).double()
    a, b = vols["fp32"], vols["fp16x2"]
    assert (a - b).abs().max().item() <= 1e-4
    flips = (a > 0) != (b > 0)
    assert not (flips & (a.abs() > 5e-5)).any()


@pytest.mark.parametrize("precision", ["fp16x2", "bf16"])
def test_output_is_independent_of_the_position_in_the_batch(precision):
    """A point's value must not depend on which tile, TMEM lane or CTA of a pair it lands in: shuffling the batch
    (ragged size: the last tile is partial) permutes the outputs bit for bit."""
    dec = util.make_decoder(0, "trained_norm", precision=precision)
    code = util.trained_code(0).cuda()
    pts = synth.make_points(7777, 11).cuda()
    g = torch.Generator(device="cpu").manual_seed(3)
    perm = torch.randperm(pts.shape[0], generator=g).cuda()
    for u in (1, 3):
        a = dec.query(code, pts, use_net=u)
        b = dec.query(code, pts[perm].contiguous(), use_net=u)
        assert torch.equal(a[perm], b)
