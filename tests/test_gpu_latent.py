"""GPU parity of the latent-code optimisation (fused loss + backward-to-code + Adam on the device)
against autograd through the oracle and against the reference's own reconstruct() run (golden)."""
import os

import numpy as np
import pytest
import torch

from oracle import decoder_oracle as O
from oracle import synth
from tests import util

pytestmark = pytest.mark.gpu
G = util.GOLDEN


@pytest.mark.parametrize("kind", ["default", "trained_norm"])
def test_latent_gradient_vs_reference_autograd(kind):
    from deepjoin_b200 import reconstruct as R
    g = np.load(os.path.join(G, "reconstruct.npz"))
    dec = util.make_decoder(0, kind, precision="fp32")
    bp, bs = g[kind + "_batch_pts"], g[kind + "_batch_sdf"]
    code0 = torch.from_numpy(g[kind + "_code0"])
    grad, terms = R.latent_grad(dec, code0.cuda(), torch.from_numpy(bp[0][:, :3]).cuda(),
                                torch.from_numpy(bs[0]).cuda(), torch.from_numpy(bp[0][:, 3:]).cuda())
    ref = g[kind + "_grad0"].reshape(-1)
    got = grad.cpu().numpy()
    assert abs(float(terms[0]) - float(g[kind + "_loss0"])) < 2e-5
    assert np.abs(got - ref).max() <= 2e-3 * np.abs(ref).max() + 1e-7, np.abs(got - ref).max() / np.abs(ref).max()
    # terms add up and match the oracle term by term
    net = util.oracle_net(0, kind)
    _, t = O.latent_grad(net, code0, torch.from_numpy(bp[0][:, :3]), torch.from_numpy(bs[0]), torch.from_numpy(bp[0][:, 3:]))
    assert np.allclose(terms.cpu().numpy(), np.asarray(t), atol=2e-5)


def test_latent_gradient_options():
    """lambda1 = 0 / lambda3 = 0 / no regulariser / other clamp: each switch against the oracle."""
    from deepjoin_b200 import reconstruct as R
    dec = util.make_decoder(1, "trained_norm", precision="fp32")
    net = util.oracle_net(1, "trained_norm")
    xyz, sdf, nrm = synth.make_sample_set(4000, seed=1)
    idx = np.random.default_rng(0).permutation(4000)[:777]
    x = torch.from_numpy(xyz[idx].astype(np.float32))
    s = torch.from_numpy(sdf[idx].astype(np.float32))
    n = torch.from_numpy(nrm[idx].astype(np.float32))
    code = util.trained_code(1)
    for kw in (dict(lambda1=0.0), dict(lambda3=0.0), dict(l2reg=False), dict(clamp_dist=0.03), dict(lambda1=0.5, lambda3=2.0)):
        ref, t = O.latent_grad(net, code, x, s, n, **kw)
        grad, terms = R.latent_grad(dec, code.cuda(), x.cuda(), s.cuda(), n.cuda(), **kw)
        ref = ref.numpy()
        assert np.abs(grad.cpu().numpy() - ref).max() <= 3e-3 * np.abs(ref).max() + 1e-7, kw
        assert np.allclose(terms.cpu().numpy(), np.asarray(t), atol=5e-5), kw


@pytest.mark.parametrize("precision", ["fp32", "fp16x2"])
@pytest.mark.parametrize("kind", ["default", "trained_norm"])
def test_adam_loop_vs_reference_reconstruct(kind, precision):
    """The device loop fed the batches the reference's own run drew reproduces its loss log and code -- in the fp32
    CUDA-core arithmetic and in the split-operand tensor-core step (the default), to the same tolerance."""
    from deepjoin_b200 import reconstruct as R
    g = np.load(os.path.join(G, "reconstruct.npz"))
    dec = util.make_decoder(0, kind, precision=precision)
    bp, bs = g[kind + "_batch_pts"], g[kind + "_batch_sdf"]
    iters, ns = bp.shape[0], bp.shape[1]
    pool_xyz = torch.from_numpy(bp[:, :, :3].reshape(-1, 3).copy()).cuda()
    pool_nrm = torch.from_numpy(bp[:, :, 3:].reshape(-1, 3).copy()).cuda()
    pool_sdf = torch.from_numpy(bs.reshape(-1).copy()).cuda()
    sched = torch.arange(iters * ns, dtype=torch.int32).reshape(iters, ns).cuda()
    log, code = R.optimize_code(dec, torch.from_numpy(g[kind + "_code0"]), pool_xyz, pool_sdf, pool_nrm, sched,
                                lr=5e-3, lambda1=1.0, lambda3=1.0, clamp_dist=0.1, l2reg=True, code_reg_lambda=1e-4,
                                precision=precision)
    log = log.cpu().numpy()
    assert log.shape == (1, 8) and log[0, 0] == 0
    assert abs(log[0, 1] - g[kind + "_log_loss"][0]) < 2e-5
    assert abs(log[0, 6] - g[kind + "_log_mag"][0]) < 1e-5 and abs(log[0, 7] - g[kind + "_log_h_mag"][0]) < 1e-5
    assert np.abs(code.cpu().numpy() - g[kind + "_code"]).max() < 2e-3


@pytest.mark.parametrize("precision", ["fp32", "fp16x2"])
def test_reconstruct_dropin_seeded_like_the_reference(precision):
    """reconstruct(): same seeds -> same initial code and the same mini-batches as the reference run ("fp16x2" is what
    a decoder uses unless told otherwise; the native sampler walks numpy's stream)."""
    from deepjoin_b200.reconstruct import reconstruct
    g = np.load(os.path.join(G, "reconstruct.npz"))
    dec = util.make_decoder(0, "default", precision=precision)
    xyz, sdf, nrm = synth.make_sample_set(4000, seed=3)
    np.random.seed(5)
    torch.manual_seed(5)
    loss, code = reconstruct(dec, 6, 128, 64, [xyz, sdf, nrm], 0.01, 1.0, 0.0, 1.0, 0.1, num_samples=512, lr=5e-3,
                             l2reg=True, code_reg_lambda=1e-4)
    assert set(loss) == {"epoch", "loss", "sdf_data_loss", "occ_data_loss", "norm_loss", "reg_loss", "mag", "h_mag"}
    assert loss["epoch"] == [0]
    assert abs(loss["loss"][0] - g["default_log_loss"][0]) < 2e-5
    assert tuple(code.shape) == (1, 192) and code.is_cuda
    assert np.abs(code.cpu().numpy() - g["default_code"]).max() < 2e-3


def test_longer_run_decreases_loss_and_logs_every_10():
    from deepjoin_b200.reconstruct import reconstruct
    dec = util.make_decoder(0, "trained_norm", precision="fp32")
    xyz, sdf, nrm = synth.make_sample_set(20000, seed=2)
    np.random.seed(1)
    torch.manual_seed(1)
    loss, code = reconstruct(dec, 60, 128, 64, [xyz, sdf, nrm], 0.01, 1.0, 0.0, 1.0, 0.1, num_samples=2000, lr=5e-3,
                             l2reg=True)
    assert loss["epoch"] == [0, 10, 20, 30, 40, 50]
    assert loss["loss"][-1] < loss["loss"][0]
    assert torch.isfinite(code).all()


def test_device_sampler_semantics():
    """dj_sample_schedule: class balance, [pos..., neg...] order, no duplicates inside a class window,
    surface rows + 1/4 of the uniform rows, NoSamplesError on an empty class."""
    from deepjoin_b200 import _lib
    from deepjoin_b200.core import errors
    xyz, sdf, nrm = synth.make_sample_set(40000, seed=0)
    s = torch.from_numpy(sdf.reshape(-1).astype(np.float32)).cuda()
    iters, ns = 7, 3001
    idx = torch.empty((iters, ns), dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().dj_sample_schedule(s.data_ptr(), s.shape[0], iters, ns, 0.2, 1234, idx.data_ptr(), None))
    idx = idx.cpu().numpy()
    sv = sdf.reshape(-1).astype(np.float32)
    assert idx.min() >= 0 and idx.max() < 40000
    for e in range(iters):
        row = idx[e]
        assert (sv[row[: ns // 2]] > 0).all() and (sv[row[ns // 2:]] <= 0).all()
        assert len(np.unique(row)) == ns
    assert not np.array_equal(idx[0], idx[1])
    frac_uniform = (idx >= 20000).mean()
    assert 0.1 < frac_uniform < 0.5
    ones = torch.ones(1000, device="cuda")
    with pytest.raises(errors.NoSamplesError):
        _lib.check(_lib.lib().dj_sample_schedule(ones.data_ptr(), 1000, 2, 100, 0.2, 1, idx_dev_ptr(), None))


def idx_dev_ptr():
    t = torch.empty(4096, dtype=torch.int32, device="cuda")
    idx_dev_ptr.keep = t
    return t.data_ptr()


# ---------------------------------------------------------------------------------------------
# tensor-core latent step (csrc/latent_tc2.cuh): bf16 operands, fp32 accumulation.  Tolerances: the
# gradient carries bf16 rounding noise (north_star: fp32-accumulated bf16), loss terms are those of the
# bf16 forward.
# ---------------------------------------------------------------------------------------------
def _cos(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a * b).sum() / (a.norm() * b.norm()))


@pytest.mark.parametrize("kind,tol_g,tol_l", [("default", 3e-2, 1e-4), ("trained_norm", 6e-2, 2e-2)])
@pytest.mark.parametrize("n", [1, 100, 777, 4000])
def test_tensor_core_latent_gradient_vs_autograd(kind, tol_g, tol_l, n):
    from deepjoin_b200 import reconstruct as R
    dec = util.make_decoder(0, kind, precision="bf16")
    net = util.oracle_net(0, kind)
    xyz, sdf, nrm = synth.make_sample_set(8000, seed=1)
    idx = np.random.default_rng(n).permutation(8000)[:n]
    x, s, m = (torch.from_numpy(a[idx].astype(np.float32)) for a in (xyz, sdf, nrm))
    code = util.code_for(0, kind)
    ref, t = O.latent_grad(net, code, x, s, m)
    grad, terms = R.latent_grad(dec, code.cuda(), x.cuda(), s.cuda(), m.cuda(), precision="bf16")
    grad = grad.cpu()
    assert torch.isfinite(grad).all()
    if n >= 100:
        assert _cos(grad, ref) > 0.998
        # few samples: less averaging of the bf16 rounding noise
        assert float((grad - ref).abs().max()) <= tol_g * (3.0 if n < 1000 else 1.0) * float(ref.abs().max())
    assert np.allclose(terms.cpu().numpy(), np.asarray(t), rtol=tol_l, atol=tol_l)


@pytest.mark.parametrize("kind", ["default", "trained_norm"])
@pytest.mark.parametrize("n", [1, 63, 64, 65, 777, 4000, 30000])
def test_split_operand_latent_gradient_vs_autograd(kind, n):
    """DJ_PREC_FP16X2 (latent_tc2_kernel<true, true>): forward and backward with split operands.  Held to the fp32
    path's own tolerance against autograd: the loss gates (clamp window, sign(pred - gt), B = select(C, T)) are the
    fp32 ones, the gradient is fp32-class (measured 4e-6 .. 1e-4 of the largest component, the fp32 path the same)."""
    from deepjoin_b200 import reconstruct as R
    dec = util.make_decoder(1, kind, precision="fp16x2")
    net = util.oracle_net(1, kind, "float64")
    xyz, sdf, nrm = synth.make_sample_set(40000, seed=1)
    idx = np.random.default_rng(n).permutation(40000)[:n]
    x, s, m = (torch.from_numpy(a[idx].astype(np.float32)) for a in (xyz, sdf, nrm))
    code = util.code_for(1, kind)
    ref, t = O.latent_grad(net, code.double(), x.double(), s.double(), m.double())
    grad, terms = R.latent_grad(dec, code.cuda(), x.cuda(), s.cuda(), m.cuda(), precision="fp16x2")
    grad = grad.cpu().double()
    # a forward difference of 1e-5 can decide one loss gate of one sample the other way (as it can in the fp32 path):
    # that moves O(1/n) of the batch gradient
    assert float((grad - ref).abs().max()) <= (1e-3 + 2.5 / n) * float(ref.abs().max()) + 1e-7
    if n >= 777:
        assert _cos(grad, ref) > 0.999999
    assert np.allclose(terms.cpu().numpy(), np.asarray(t), rtol=0, atol=5e-5)


@pytest.mark.parametrize("precision", ["fp32", "fp16x2"])
def test_occupancy_target_of_the_fused_loss(precision):
    """tensor_sdf_to_occ inside the fused loss (a8; reconstruct.py:169 -> core/reconstruct.py:33-34): a target of exactly
    0 or -0 counts as inside.  The BCE term alone (lambda1 = lambda3 = 0) against the oracle on a batch made of such rows."""
    from deepjoin_b200 import reconstruct as R
    dec = util.make_decoder(0, "trained_norm", precision=precision)
    net = util.oracle_net(0, "trained_norm", "float64")
    xyz, sdf, nrm = synth.make_sample_set(4000, seed=1)
    x, m = (torch.from_numpy(a[:512].astype(np.float32)) for a in (xyz, nrm))
    s = torch.from_numpy(sdf[:512].astype(np.float32)).clone()
    s[0:100] = 0.0
    s[100:200] = -0.0
    s[200:300] = 1e-30
    s[300:400] = -1e-30
    code = util.trained_code(0)
    kw = dict(lambda1=0.0, lambda3=0.0, l2reg=False)
    ref, t = O.latent_grad(net, code.double(), x.double(), s.double(), m.double(), **kw)
    grad, terms = R.latent_grad(dec, code.cuda(), x.cuda(), s.cuda(), m.cuda(), precision=precision, **kw)
    assert abs(float(terms[2]) - t[2]) < 5e-5 and abs(float(terms[0]) - t[2]) < 5e-5 and float(terms[1]) == 0.0
    assert float((grad.cpu().double() - ref).abs().max()) <= 3e-3 * float(ref.abs().max()) + 1e-7
    flipped = s.clone()
    flipped[0:200] = 1e-30  # the same rows counted as outside give a different loss: the target matters
    _, t2 = O.latent_grad(net, code.double(), x.double(), flipped.double(), m.double(), **kw)
    assert abs(t2[2] - t[2]) > 1e-3


def test_split_operand_latent_options():
    from deepjoin_b200 import reconstruct as R
    dec = util.make_decoder(1, "trained_norm", precision="fp16x2")
    net = util.oracle_net(1, "trained_norm", "float64")
    xyz, sdf, nrm = synth.make_sample_set(4000, seed=1)
    x, s, m = (torch.from_numpy(a[:777].astype(np.float32)) for a in (xyz, sdf, nrm))
    code = util.trained_code(1)
    for kw in (dict(lambda1=0.0), dict(lambda3=0.0), dict(l2reg=False), dict(clamp_dist=0.03), dict(lambda1=0.5, lambda3=2.0)):
        ref, t = O.latent_grad(net, code.double(), x.double(), s.double(), m.double(), **kw)
        grad, terms = R.latent_grad(dec, code.cuda(), x.cuda(), s.cuda(), m.cuda(), precision="fp16x2", **kw)
        assert float((grad.cpu().double() - ref).abs().max()) <= 3e-3 * float(ref.abs().max()) + 1e-7, kw
        assert np.allclose(terms.cpu().numpy(), np.asarray(t), atol=5e-5), kw


def test_tensor_core_latent_gradient_options_and_multi_pass():
    """Loss switches, and a batch larger than one pass of the grid (148 CTAs x 128 rows)."""
    from deepjoin_b200 import reconstruct as R
    dec = util.make_decoder(1, "trained_norm", precision="bf16")
    net = util.oracle_net(1, "trained_norm")
    xyz, sdf, nrm = synth.make_sample_set(40000, seed=1)
    code = util.trained_code(1)
    for n, kw in ((777, dict(lambda1=0.0)), (777, dict(lambda3=0.0)), (777, dict(l2reg=False)), (777, dict(clamp_dist=0.03)),
                  (30000, dict(lambda1=0.5, lambda3=2.0))):
        x, s, m = (torch.from_numpy(a[:n].astype(np.float32)) for a in (xyz, sdf, nrm))
        ref, t = O.latent_grad(net, code, x, s, m, **kw)
        grad, terms = R.latent_grad(dec, code.cuda(), x.cuda(), s.cuda(), m.cuda(), precision="bf16", **kw)
        assert _cos(grad.cpu(), ref) > 0.998, kw
        assert float((grad.cpu() - ref).abs().max()) <= 6e-2 * float(ref.abs().max()), kw
        assert np.allclose(terms.cpu().numpy(), np.asarray(t), rtol=2e-2, atol=2e-2), kw


def test_tensor_core_adam_loop_tracks_the_fp32_loop():
    """Same schedule, same start: the bf16 loop's loss curve and final code stay close to the fp32 loop's
    (in-kernel gather of the mini-batch rows vs the fp32 path's gather kernel)."""
    from deepjoin_b200 import reconstruct as R
    dec = util.make_decoder(0, "trained_norm", precision="fp32")
    xyz, sdf, nrm = synth.make_sample_set(20000, seed=2)
    px, ps, pn = (torch.from_numpy(a.astype(np.float32)).cuda() for a in (xyz, sdf.reshape(-1), nrm))
    sched = torch.from_numpy(np.random.default_rng(0).integers(0, 20000, (80, 3000)).astype(np.int32)).cuda()
    code0 = R.init_code(128, 64, 0.01)
    out = {}
    for prec in ("fp32", "bf16"):
        log, code = R.optimize_code(dec, code0, px, ps, pn, sched, lr=5e-3, lambda1=1.0, lambda3=1.0, clamp_dist=0.1,
                                    l2reg=True, precision=prec)
        out[prec] = (log.cpu().numpy(), code.cpu().numpy())
    l32, l16 = out["fp32"][0][:, 1], out["bf16"][0][:, 1]
    assert l16[-1] < l16[0] and np.abs(l16 - l32).max() < 0.05 * np.abs(l32).max()
    assert np.abs(out["bf16"][1] - out["fp32"][1]).max() < 0.05 * np.abs(out["fp32"][1]).max() + 2e-2


def test_reconstruct_in_blocks_equals_one_call():
    """The loop continued block by block (DjLatentCfg.iter_offset / total_iterations: what reconstruct() does while the
    host sampler runs ahead) is the loop in one call: lr step at iterations/2, Adam bias correction, log rows.
    Short horizon: Adam's lr-sized steps amplify the run-to-run order of the atomic accumulations over long runs."""
    from deepjoin_b200 import reconstruct as R
    dec = util.make_decoder(0, "trained_norm", precision="fp32")
    xyz, sdf, nrm = synth.make_sample_set(6000, seed=2)
    px, ps, pn = (torch.from_numpy(a.astype(np.float32)).cuda() for a in (xyz, sdf.reshape(-1), nrm))
    sched = np.random.default_rng(0).integers(0, 6000, (9, 2000)).astype(np.int32)
    code0 = R.init_code(128, 64, 0.01)
    kw = dict(lr=5e-3, lambda1=1.0, lambda3=1.0, clamp_dist=0.1, l2reg=True, log_every=2)
    for prec in ("fp32", "fp16x2"):
        log1, c1 = R.optimize_code(dec, code0, px, ps, pn, torch.from_numpy(sched).cuda(), precision=prec, **kw)
        log2, c2 = R.optimize_code(dec, code0, px, ps, pn, None, precision=prec, total_iterations=9,
                                   schedule_blocks=iter([sched[:3], sched[3:4], sched[4:]]), **kw)
        assert log1.shape == log2.shape == (5, 8)
        assert log1[:, 0].tolist() == [0, 2, 4, 6, 8] == log2[:, 0].tolist()
        assert torch.allclose(log1, log2, rtol=2e-3, atol=1e-4), (prec, (log1 - log2).abs().max())
        assert torch.allclose(c1, c2, atol=2e-3), (prec, (c1 - c2).abs().max())
        # the lr step (x0.1 from iteration int(9/2) = 4) is taken in the right block: a loop without it moves further
        assert float((c1 - code0.cuda()).abs().max()) < 9 * 5e-3
    with pytest.raises(ValueError):
        R.optimize_code(dec, code0, px, ps, pn, None, total_iterations=50, schedule_blocks=iter([sched[:3]]), **kw)


def test_reconstruct_follows_decoder_precision():
    from deepjoin_b200.reconstruct import reconstruct
    dec = util.make_decoder(0, "trained_norm", precision="bf16")
    xyz, sdf, nrm = synth.make_sample_set(20000, seed=2)
    np.random.seed(1)
    torch.manual_seed(1)
    loss, code = reconstruct(dec, 40, 128, 64, [xyz, sdf, nrm], 0.01, 1.0, 0.0, 1.0, 0.1, num_samples=2000, lr=5e-3,
                             l2reg=True, sampling_method="device")
    assert loss["epoch"] == [0, 10, 20, 30] and loss["loss"][-1] < loss["loss"][0]
    assert torch.isfinite(code).all()


@pytest.mark.parametrize("precision", ["bf16", "fp16x2", "fp32"])
def test_batched_shapes_equal_one_by_one(precision):
    """dj_latent_optimize_batch: S shapes in one launch per iteration == the same shapes optimised separately
    (bf16: up to the order of the atomic column sums)."""
    from deepjoin_b200 import reconstruct as R
    dec = util.make_decoder(0, "trained_norm", precision=precision)
    S, iters, ns = 3, 4 if precision == "fp32" else 12, 1000
    pools, scheds, codes0 = [], [], []
    for s in range(S):
        xyz, sdf, nrm = synth.make_sample_set(6000 + 500 * s, seed=10 + s)
        pools.append(tuple(torch.from_numpy(a.astype(np.float32)).cuda() for a in (xyz, sdf.reshape(-1), nrm)))
        scheds.append(torch.from_numpy(np.random.default_rng(s).integers(0, 6000, (iters, ns)).astype(np.int32)))
        torch.manual_seed(s)
        codes0.append(R.init_code(128, 64, 0.01))
    sched = torch.stack(scheds).cuda()
    logs, codes = R.optimize_codes(dec, torch.cat(codes0), pools, sched, 5e-3, 1.0, 1.0, 0.1, l2reg=True, precision=precision)
    assert tuple(logs.shape) == (S, 1 if precision == "fp32" else 2, 8) and tuple(codes.shape) == (S, 192)
    for s in range(S):
        log1, code1 = R.optimize_code(dec, codes0[s], *pools[s], sched[s], 5e-3, 1.0, 1.0, 0.1, l2reg=True, precision=precision)
        a, b = logs[s].cpu().numpy(), log1.cpu().numpy()
        assert np.allclose(a[0], b[0], rtol=1e-4, atol=1e-6), s  # iteration 0: same arithmetic
        # later iterations: the atomic column sums add in a different order, Adam (step = lr per element whatever
        # the gradient's size) amplifies that to a few lr
        tol = 3e-2 if precision == "bf16" else (5e-3 if precision == "fp16x2" else 2e-3)
        assert np.allclose(a, b, rtol=tol, atol=tol), s
        assert np.abs(codes[s].cpu().numpy() - code1.cpu().numpy().reshape(-1)).max() < (2e-2 if precision == "bf16" else 5e-3), s
    assert not np.allclose(codes[0].cpu().numpy(), codes[1].cpu().numpy(), atol=1e-3)  # really different shapes


def test_reconstruct_surface_interior_only_like_the_reference():
    """sampling_method="surface_interior_only" (reconstruct.py:60-98): same seeds -> the reference's log and code."""
    from deepjoin_b200.reconstruct import reconstruct
    g = np.load(os.path.join(G, "reconstruct_surface_interior.npz"))
    for kind in ("default", "trained_norm"):
        dec = util.make_decoder(0, kind, precision="fp32")
        np.random.seed(7)
        torch.manual_seed(7)
        loss, code = reconstruct(dec, 12, 128, 64, [g["xyz"], g["sdf"]], 0.01, 0.0, 0.0, 1.0, 0.1,
                                 sampling_method="surface_interior_only", num_samples=500, lr=5e-3, l2reg=True, code_reg_lambda=1e-4)
        assert loss["epoch"] == [0, 10] and loss["norm_loss"] == [0.0, 0.0]
        assert np.abs(np.asarray(loss["loss"]) - g[kind + "_log_loss"]).max() < 1e-4
        assert np.abs(np.asarray(loss["occ_data_loss"]) - g[kind + "_log_occ_data_loss"]).max() < 1e-4
        assert np.abs(code.cpu().numpy() - g[kind + "_code"]).max() < 2e-3
    with pytest.raises(RuntimeError):  # the reference's normal term cannot be evaluated without normals
        reconstruct(dec, 2, 128, 64, [g["xyz"], g["sdf"]], 0.01, 1.0, 0.0, 1.0, 0.1, sampling_method="surface_interior_only",
                    num_samples=100, l2reg=True)
