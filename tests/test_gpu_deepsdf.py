"""the DeepSDF baseline decoder and its grid query / meshing on the fused kernels."""
import os

import numpy as np
import pytest
import torch

from oracle import decoder_oracle as O
from oracle import mc_oracle, synth
from tests import util

pytestmark = pytest.mark.gpu
G = util.GOLDEN
SPECS = dict(dims=[512] * 8, dropout=list(range(8)), dropout_prob=0.2, norm_layers=list(range(8)), latent_in=[4],
             xyz_in_all=False, use_tanh=False, latent_dropout=False, weight_norm=True)


def _decoder(seed, kind, precision):
    from deepjoin_b200.networks.decoder_deepsdf import Decoder
    dec = Decoder(128, **SPECS)
    dec.load_state_dict(synth.make_deepsdf_state_dict(seed, kind, prefix="module."))
    dec.precision = precision
    return dec.cuda().eval()


@pytest.mark.parametrize("kind,precision,tol", [("default", "fp32", 2e-5), ("trained_norm", "fp32", 3e-4),
                                                ("default", "bf16", 1e-3), ("trained_norm", "bf16", 0.3)])
@pytest.mark.parametrize("seed", [0, 1])
def test_forward_matches_the_reference_decoder(kind, precision, tol, seed):
    g = np.load(os.path.join(G, "deepsdf_forward.npz"))
    dec = _decoder(seed, kind, precision)
    code = torch.from_numpy(g["%s_%d_code" % (kind, seed)])
    pts = torch.from_numpy(g["%s_%d_pts" % (kind, seed)])
    out = dec(torch.cat([code.reshape(1, -1).expand(pts.shape[0], -1), pts], 1).cuda()).cpu().numpy()
    ref = g["%s_%d_out" % (kind, seed)]
    assert out.shape == ref.shape and np.abs(out - ref).max() <= tol
    if precision == "bf16" and kind == "trained_norm":
        assert np.abs(out - ref).mean() < 2e-2


def test_create_mesh_and_values2mesh():
    from deepjoin_b200.core import reconstruct as R, utils_deepsdf as U
    from deepjoin_b200.mesh import Mesh
    dec = _decoder(0, "trained_norm", "fp32")
    latent = synth.trained_code(0)[:128]
    N = 40
    mesh = U.create_mesh(dec, latent.reshape(1, -1), N=N, padding=0.1)
    vol = R.decode_shape_device(dec.code192(latent), dec, (N, N, N), use_net=0, padding=0.1, sigmoid=False).reshape(N, N, N)
    # the grid itself against the oracle restatement of the reference decoder
    pts = torch.from_numpy(R.get_query_points((N, N, N), 0.1)).float()
    ref = O.deepsdf_forward(synth.make_deepsdf_state_dict(0, "trained_norm"), latent, pts).numpy().reshape(N, N, N)
    assert np.abs(vol.cpu().numpy() - ref).max() < 3e-4
    vo, fo = mc_oracle.mesh_from_volume(vol.cpu().numpy(), 0.0, (N, N, N), 0.1, "ascent")
    want = Mesh(vo, fo)
    assert np.array_equal(mesh.faces, want.faces) and np.array_equal(mesh.vertices, want.vertices)
    again = U.values2mesh(vol.cpu().numpy(), [-(0.5 + 0.1)] * 3, (1 + 0.2) / N, gradient_direction="ascent", flipxy=True)
    assert np.array_equal(again.faces, want.faces) and np.allclose(again.vertices, want.vertices, rtol=0, atol=1e-12)
    sdf = U.decode_sdf(dec, latent.reshape(1, -1), pts[:100])
    assert np.abs(sdf.cpu().numpy().reshape(-1) - ref.reshape(-1)[:100]).max() < 3e-4


# ---------------------------------------------------------------- latent optimisation (reconstruct_deepsdf.py:17-154)
@pytest.mark.parametrize("kind", ["default", "trained_norm"])
@pytest.mark.parametrize("precision", ["fp32", "fp16x2", "bf16", "fp16"])
def test_latent_gradient_matches_reference_autograd(kind, precision):
    """d loss / d latent of the baseline objective against the reference's own autograd (golden grad0)."""
    from deepjoin_b200 import reconstruct as R
    g = np.load(os.path.join(G, "deepsdf_reconstruct.npz"))
    dec = _decoder(0, kind, precision)
    lat0 = torch.from_numpy(g[kind + "_code0"]).reshape(-1)
    xyz = torch.from_numpy(g[kind + "_batch_pts"][0]).cuda()
    sdf = torch.from_numpy(g[kind + "_batch_sdf"][0]).reshape(-1).cuda()
    grad, terms = R.latent_grad(dec, dec.code192(lat0), xyz, sdf, xyz, clamp_dist=0.1, l2reg=True, code_reg_lambda=1e-4,
                                precision=precision, loss_kind=1)
    grad, terms = grad.cpu().numpy(), terms.cpu().numpy()
    ref = g[kind + "_grad0"].reshape(-1)
    assert grad.shape == (129,) and grad[128] == 0.0  # the stand-in second net contributes nothing
    assert terms[2] == 0.0 and terms[3] == 0.0       # no occupancy / normal terms in this objective
    if precision in ("fp32", "fp16x2"):  # the split-operand tensor-core step is held to the fp32 path's tolerance
        assert abs(terms[0] - float(g[kind + "_loss0"])) < (2e-6 if precision == "fp32" else 2e-5)
        assert np.abs(grad[:128] - ref).max() <= 2e-3 * np.abs(ref).max() + 1e-7
    else:
        # 16-bit forward operands: a rounding error that flips sign(pred - gt) or the clamp gate flips that sample's
        # whole gradient; 512 samples average less of it out than 30,000 do.  IEEE half is 8x finer than bf16.
        cos = float(np.dot(grad[:128], ref) / (np.linalg.norm(grad[:128]) * np.linalg.norm(ref)))
        print(kind, precision, "cos", cos, "max err / max", np.abs(grad[:128] - ref).max() / np.abs(ref).max(), "loss", terms[0], float(g[kind + "_loss0"]))
        assert abs(terms[0] - float(g[kind + "_loss0"])) < (5e-3 if precision == "bf16" else 1e-3)
        assert np.abs(grad[:128] - ref).max() <= (0.15 if precision == "bf16" else 0.08) * np.abs(ref).max()
        assert cos > (0.98 if precision == "bf16" else 0.995)


@pytest.mark.parametrize("precision", ["fp32", "fp16x2"])
@pytest.mark.parametrize("kind", ["default", "trained_norm"])
def test_latent_loop_vs_reference_run(kind, precision):
    """The device loop fed the mini-batches the reference's run drew reproduces its log and its latent."""
    from deepjoin_b200 import reconstruct as R
    g = np.load(os.path.join(G, "deepsdf_reconstruct.npz"))
    dec = _decoder(0, kind, precision)
    bp, bs = g[kind + "_batch_pts"], g[kind + "_batch_sdf"]
    iters, ns = bp.shape[0], bp.shape[1]
    pool_xyz = torch.from_numpy(bp.reshape(-1, 3).copy()).cuda()
    pool_sdf = torch.from_numpy(bs.reshape(-1).copy()).cuda()
    sched = torch.arange(iters * ns, dtype=torch.int32).reshape(iters, ns).cuda()
    log, code = R.optimize_code(dec, dec.code192(torch.from_numpy(g[kind + "_code0"])), pool_xyz, pool_sdf, pool_xyz, sched,
                                lr=5e-3, lambda1=0.0, lambda3=1.0, clamp_dist=0.1, l2reg=True, code_reg_lambda=1e-4,
                                precision=precision, loss_kind=1)
    log = log.cpu().numpy()
    assert log.shape == (2, 8) and list(log[:, 0]) == [0, 10]
    assert np.abs(log[:, 2] - g[kind + "_log_data_loss"]).max() < 2e-5
    assert np.allclose(log[:, 5], g[kind + "_log_reg_loss"], rtol=1e-3, atol=1e-9)
    assert np.allclose(log[:, 6], g[kind + "_log_mag"], atol=2e-4)  # |z| AFTER the step (:134-141)
    assert np.abs(code.cpu().numpy()[:, :128] - g[kind + "_code"]).max() < 2e-3
    assert float(code[0, 128]) == 0.0


def test_reconstruct_deepsdf_dropin_seeded_like_the_reference():
    from deepjoin_b200.reconstruct_deepsdf import reconstruct
    g = np.load(os.path.join(G, "deepsdf_reconstruct.npz"))
    dec = _decoder(0, "default", "fp32")
    np.random.seed(5)
    torch.manual_seed(5)
    loss, lat = reconstruct(dec, 12, 128, (g["xyz"], g["sdf"], {"mask": g["mask"], "broken_gt": None, "percent": 0.0}), 0.01, 0.1,
                            num_samples=512, lr=5e-3, l2reg=True, code_reg_lambda=1e-4, slippage_method="analytical")
    assert set(loss) == {"epoch", "data_loss", "reg_loss", "mag"} and loss["epoch"] == [0, 10]
    assert np.abs(np.asarray(loss["data_loss"]) - g["default_log_data_loss"]).max() < 2e-5
    assert tuple(lat.shape) == (1, 128) and not lat.is_cuda
    assert np.abs(lat.numpy() - g["default_code"]).max() < 2e-3


def test_reconstruct_deepsdf_tensor_core_run_and_errors():
    from deepjoin_b200.reconstruct_deepsdf import reconstruct
    xyz, sdf, _ = synth.make_sample_set(20000, seed=2)
    mask = xyz[:, 1].astype(np.float32) > -0.2
    # Short horizon, small lr: with these synthetic (untrained, large-norm) weights the sign-gated L1 landscape is
    # rough and Adam's lr-sized steps amplify any gradient difference, so long runs of different arithmetic (even
    # the same arithmetic with another atomic order) end in different basins; over 20 small steps the tensor-core
    # run (half forward, bf16 backward) still has to move the latent the way the fp32 run does.
    runs = {}
    for precision in ("fp32", "bf16"):
        dec = _decoder(0, "trained_norm", precision)
        assert dec.latent_prec() == {"fp32": "fp32", "bf16": "fp16"}[precision]
        np.random.seed(1)
        torch.manual_seed(1)
        runs[precision] = reconstruct(dec, 20, 128, (xyz, sdf, {"mask": mask}), 0.01, 0.1, num_samples=8000, lr=5e-4,
                                      l2reg=True, slippage_method="classifier")
    (loss, lat), (loss32, lat32) = runs["bf16"], runs["fp32"]
    assert loss["epoch"] == [0, 10] and torch.isfinite(lat).all()
    torch.manual_seed(1)
    lat0 = torch.ones(1, 128).normal_(mean=0, std=0.01)
    d16, d32 = (lat - lat0).reshape(-1), (lat32 - lat0).reshape(-1)
    cos = float((d16 * d32).sum() / (d16.norm() * d32.norm()))
    print("data_loss tc", loss["data_loss"], "fp32", loss32["data_loss"], "cos of the latent displacement", cos)
    assert np.abs(np.asarray(loss["data_loss"]) - np.asarray(loss32["data_loss"])).max() < 5e-3
    assert cos > 0.8
    with pytest.raises(RuntimeError, match="Unknown slippage method"):
        reconstruct(dec, 4, 128, (xyz, sdf, {"mask": mask}), 0.01, 0.1, num_samples=100, l2reg=True)
    with pytest.raises(NotImplementedError):
        reconstruct(dec, 4, 128, (xyz, sdf, {"percent": 0.1}), 0.01, 0.1, num_samples=100, l2reg=True, slippage_method="analytical")
    with pytest.raises(UnboundLocalError):  # the reference logs reg_loss without having computed it (:139)
        reconstruct(dec, 4, 128, (xyz, sdf, {"mask": mask}), 0.01, 0.1, num_samples=100, slippage_method="analytical")
