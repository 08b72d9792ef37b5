"""Pin oracle/decoder_oracle.py to the reference through the golden fixtures
(made by tests/golden/make_golden.py from the reference executed in place)."""
import os

import numpy as np
import pytest
import torch

from oracle import decoder_oracle as O
from oracle import synth

G = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ("c_occ", "b_occ", "r_occ", "t_occ", "c_sdf", "b_sdf", "r_sdf", "t_sdf", "c_n", "b_n", "r_n", "t_n")


@pytest.mark.parametrize("kind", ["default", "trained_norm"])
@pytest.mark.parametrize("seed", [0, 1])
def test_forward_matches_reference(kind, seed):
    g = np.load(os.path.join(G, "decoder_forward.npz"))
    tag = "%s_s%d" % (kind, seed)
    net = O.fold(synth.make_state_dict(seed, kind))
    code = torch.from_numpy(g[tag + "_code"])
    pts = torch.from_numpy(g[tag + "_pts"])
    tol = 2e-6 if kind == "default" else 5e-5  # fp32 summation-order noise through 9 layers
    outs = O.forward_all(net, code, pts)
    for nm, o in zip(NAMES, outs):
        ref = g[tag + "_all_" + nm]
        scale = max(1.0, float(np.abs(ref).max()))
        assert np.abs(o.numpy() - ref).max() <= tol * scale, nm
    for u in (0, 1, 2, 3):
        assert np.abs(O.forward(net, code, pts, u).numpy() - g[tag + "_u%d" % u]).max() <= tol
        ref = g[tag + "_u%d_occ" % u]
        scale = max(1.0, float(np.abs(ref).max()))
        assert np.abs(O.forward(net, code, pts, u, return_occ=True).numpy() - ref).max() <= tol * scale


def test_fp64_oracle_close_to_reference_fp32():
    g = np.load(os.path.join(G, "decoder_forward.npz"))
    net = O.fold(synth.make_state_dict(0, "trained_norm"), dtype=torch.float64)
    out = O.forward(net, torch.from_numpy(g["trained_norm_s0_code"]), torch.from_numpy(g["trained_norm_s0_pts"]), 1)
    assert np.abs(out.numpy() - g["trained_norm_s0_u1"]).max() < 1e-4


def test_bad_use_net_raises():
    net = O.fold(synth.make_state_dict(0, "default"))
    with pytest.raises(RuntimeError):
        O.forward(net, synth.make_code(0), synth.make_points(4), use_net=4)


@pytest.mark.parametrize("dims", [(1, 9, 7), (8, 1, 5), (4, 6, 1), (1, 1, 1)])
def test_grid_layout_with_an_axis_of_length_one(dims):
    g = np.load(os.path.join(G, "grid_extra.npz"))
    t = "d%dx%dx%d" % dims
    assert np.array_equal(O.get_query_points(dims, 0.1), g[t + "_pts"])
    net = O.fold(synth.make_state_dict(0, "trained_norm"))
    code = torch.from_numpy(g["code"])
    for u in (1, 3):
        v = O.decode_shape(net, code, dims, use_net=u, batch_size=500, padding=0.1, reshape=False, sigmoid=False)
        assert np.abs(v - g[t + "_u%d_flat" % u]).max() < 5e-5


@pytest.mark.parametrize("dims", [(12, 12, 12), (6, 9, 11)])
def test_grid_layout(dims):
    g = np.load(os.path.join(G, "grid.npz"))
    t = "d%dx%dx%d" % dims
    pts = O.get_query_points(dims, 0.1)
    assert pts.dtype == np.float64 and np.array_equal(pts, g[t + "_pts"])
    net = O.fold(synth.make_state_dict(0, "trained_norm"))
    code = torch.from_numpy(g["code"])
    for u in (0, 1, 2, 3):
        v = O.decode_shape(net, code, dims, use_net=u, batch_size=500, padding=0.1, reshape=False, sigmoid=False)
        assert v.dtype == np.float64 and v.shape == g[t + "_u%d_flat" % u].shape
        assert np.abs(v - g[t + "_u%d_flat" % u]).max() < 5e-5
    v = O.decode_shape(net, code, dims, use_net=1, padding=0.1, reshape=True, sigmoid=True)
    assert v.shape == g[t + "_u1_reshape_sig"].shape
    assert np.abs(v - g[t + "_u1_reshape_sig"]).max() < 5e-5


def test_sampler_bit_exact():
    g = np.load(os.path.join(G, "sampler.npz"))
    for seed, n, key in ((11, 600, "sel"), (12, 601, "sel2")):
        np.random.seed(seed)
        p, v = O.select_samples(np.hstack((g["xyz"], g["n"])), g["sdf"], n, uniform_ratio=0.2)
        assert np.array_equal(p, g[key + "_pts"]) and np.array_equal(v, g[key + "_vals"])
        assert (v[: n // 2] > 0).all() and (v[n // 2:] <= 0).all()


def test_sampler_no_samples_error():
    pts = np.random.rand(100, 6)
    vals = np.ones((100, 1))
    with pytest.raises(O.NoSamplesError):
        O.select_samples(pts, vals, 10, uniform_ratio=None)


@pytest.mark.parametrize("kind", ["default", "trained_norm"])
def test_latent_grad_and_adam_loop(kind):
    g = np.load(os.path.join(G, "reconstruct.npz"))
    net = O.fold(synth.make_state_dict(0, kind))
    bp, bs = g[kind + "_batch_pts"], g[kind + "_batch_sdf"]
    code0 = torch.from_numpy(g[kind + "_code0"])
    grad, terms = O.latent_grad(net, code0, torch.from_numpy(bp[0][:, :3]), torch.from_numpy(bs[0]),
                                torch.from_numpy(bp[0][:, 3:]))
    ref = g[kind + "_grad0"].reshape(-1)
    assert abs(terms[0] - float(g[kind + "_loss0"])) < 1e-5
    assert np.abs(grad.numpy() - ref).max() <= 2e-3 * np.abs(ref).max() + 1e-7

    def batches(e):
        return (torch.from_numpy(bp[e][:, :3]), torch.from_numpy(bs[e]), torch.from_numpy(bp[e][:, 3:]))

    log, code = O.adam_loop(net, code0, batches, bp.shape[0], lr=5e-3)
    # Adam's first steps are sign-like (lr-sized): tolerate gradient noise on tiny components
    assert np.abs(code.numpy() - g[kind + "_code"]).max() < 2e-3
    assert np.abs(np.asarray(log["loss"]) - g[kind + "_log_loss"]).max() < 1e-4
    assert np.allclose(log["mag"], g[kind + "_log_mag"], atol=1e-5)


@pytest.mark.parametrize("kind", ["default", "trained_norm"])
@pytest.mark.parametrize("seed", [0, 1])
def test_deepsdf_oracle_matches_reference(kind, seed):
    """oracle.deepsdf_forward against outputs of the reference's own networks/decoder_deepsdf.py
    (tests/golden/make_golden_deepsdf.py)."""
    g = np.load(os.path.join(G, "deepsdf_forward.npz"))
    sd = synth.make_deepsdf_state_dict(seed, kind)
    out = O.deepsdf_forward(sd, torch.from_numpy(g["%s_%d_code" % (kind, seed)]), torch.from_numpy(g["%s_%d_pts" % (kind, seed)]))
    assert np.abs(out.numpy() - g["%s_%d_out" % (kind, seed)]).max() < (5e-6 if kind == "default" else 2e-4)


@pytest.mark.parametrize("kind", ["default", "trained_norm"])
def test_deepsdf_latent_loop_matches_reference(kind):
    """The oracle's restatement of the baseline's latent loop (reconstruct_deepsdf.py:17-154) against the
    reference executed in place (tests/golden/make_golden_deepsdf.py): gradient, loss, Adam, logging."""
    g = np.load(os.path.join(G, "deepsdf_reconstruct.npz"))
    sd = synth.make_deepsdf_state_dict(0, kind)
    bp, bs = g[kind + "_batch_pts"], g[kind + "_batch_sdf"]
    lat0 = torch.from_numpy(g[kind + "_code0"])
    grad, terms = O.deepsdf_latent_grad(sd, lat0, torch.from_numpy(bp[0]), torch.from_numpy(bs[0]))
    ref = g[kind + "_grad0"].reshape(-1)
    assert abs(terms[0] - float(g[kind + "_loss0"])) < 1e-6
    assert np.abs(grad.numpy() - ref).max() <= 2e-3 * np.abs(ref).max() + 1e-7
    log, code = O.deepsdf_adam_loop(sd, lat0, lambda e: (torch.from_numpy(bp[e]), torch.from_numpy(bs[e])), bp.shape[0], lr=5e-3)
    assert np.abs(code.numpy() - g[kind + "_code"]).max() < 2e-3
    assert log["epoch"] == list(g[kind + "_log_epoch"].astype(int))
    assert np.abs(np.asarray(log["data_loss"]) - g[kind + "_log_data_loss"]).max() < 1e-4
    assert np.allclose(log["reg_loss"], g[kind + "_log_reg_loss"], rtol=1e-3, atol=1e-9)
    assert np.allclose(log["mag"], g[kind + "_log_mag"], atol=2e-4)


def test_partial_sampler_restatement_matches_reference():
    g = np.load(os.path.join(G, "deepsdf_reconstruct.npz"))
    np.random.seed(21)
    p, v = O.select_partial_samples(g["xyz"], g["sdf"], g["mask"], num_samples=500, uniform_ratio=0.2)
    assert np.array_equal(p, g["sel_pts"]) and np.array_equal(v, g["sel_vals"])
