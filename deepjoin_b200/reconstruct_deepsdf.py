"""Drop-in for the function ``reconstruct()`` of the reference's baseline driver
``reconstruct_deepsdf.py:17-154``: Adam on the [1, L] latent of the frozen vanilla-DeepSDF decoder,
objective mean |clamp(decoder(z, x), +-d) - clamp(sdf, +-d)| (+ code_reg_lambda * mean(z^2)), lr / 10
from iteration int(iters / 2), mini-batches from ``core.data.select_partial_samples`` (uniform_ratio 0.2)
restricted to the rows a fracture mask keeps.

Same arguments and return value ``(loss_dict, latent[1, L])``; the mini-batch stream consumes numpy's global
RNG exactly as the reference does and the initial latent comes from torch's CPU generator (:41-44), while the
loop itself -- gather, forward, loss, backward to the latent, Adam, logging every 10th iteration -- runs on
the device through the same fused kernels as DeepJoin's (``dj_latent_optimize`` with DJ_LOSS_DEEPSDF: the
single net of the pack, its first head row, no second net).

The fracture masks the reference builds from ground-truth meshes (``core.data.get_fracture_point_mask``,
``partial_sample_*``, reconstruct_deepsdf.py:60-96; trimesh ray casting, a PointNet classifier) are outside this
package: pass the boolean mask they produce as ``sample_kwargs["mask"]``."""
import numpy as np
import torch

from deepjoin_b200 import reconstruct as _rec
from deepjoin_b200.core import data as _data

SLIPPAGE_METHODS = ("fake-classifier", "classifier", "analytical")
LOSS_DEEPSDF = 1


def make_schedule(data_sdf, mask, num_iterations, num_samples, uniform_ratio=0.2):
    """Row indices of every iteration's mini-batch (reconstruct_deepsdf.py:102-108 -> core/data.py:76-119)."""
    sched = np.empty((num_iterations, num_samples), dtype=np.int32)
    vals = _data._f32(np.asarray(data_sdf).reshape(-1))  # widened once, not per iteration
    for e in range(num_iterations):
        sched[e] = _data.select_partial_sample_indices(vals, mask, num_samples, uniform_ratio=uniform_ratio)
    return sched


def reconstruct(decoder, num_iterations, latent_size, test_sdf, stat, clamp_dist, num_samples=30000, lr=5e-4,
                l2reg=False, code_reg_lambda=1e-4, slippage_method=None):
    decoder.eval()
    dev = next(decoder.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("deepjoin_b200: the decoder must live on a CUDA device")
    if latent_size != decoder.latent_size:
        raise ValueError("latent size does not match the decoder")
    # :41-44 (CPU generator, so torch.manual_seed reproduces the reference)
    if type(stat) == type(0.1):
        latent = torch.ones(1, latent_size).normal_(mean=0, std=stat)
    else:
        latent = torch.normal(stat[0].detach(), stat[1].detach())

    pts, data_sdf, sample_kwargs = test_sdf
    if slippage_method not in SLIPPAGE_METHODS:
        raise RuntimeError("Unknown slippage method: {}".format(slippage_method))
    mask = None if sample_kwargs is None else sample_kwargs.get("mask")
    if mask is None:
        raise NotImplementedError(
            "slippage_method=%r: build the fracture mask with the reference's core.data functions "
            "(reconstruct_deepsdf.py:60-96) and pass it as sample_kwargs['mask']" % (slippage_method,))
    if not l2reg:
        # the reference logs reg_loss.item() at iteration 0 whether or not it was computed (:139)
        raise UnboundLocalError("local variable 'reg_loss' referenced before assignment")

    pool_xyz = torch.from_numpy(np.ascontiguousarray(pts, dtype=np.float32)).to(dev)
    pool_sdf = torch.from_numpy(np.ascontiguousarray(np.asarray(data_sdf).reshape(-1), dtype=np.float32)).to(dev)
    sched = torch.from_numpy(make_schedule(data_sdf, mask, num_iterations, num_samples, 0.2)).to(dev)
    log, code = _rec.optimize_code(
        decoder, decoder.code192(latent), pool_xyz, pool_sdf, pool_xyz, sched, lr, 0.0, 1.0, clamp_dist, l2reg=l2reg,
        code_reg_lambda=code_reg_lambda, log_every=10, loss_kind=LOSS_DEEPSDF,
        precision=decoder.latent_prec())
    log = log.cpu().numpy()
    loss_dict = {"epoch": [], "data_loss": [], "reg_loss": [], "mag": []}
    for row in log:  # columns: epoch, loss, sdf, occ, normal, reg, |z|, |z_b|
        loss_dict["epoch"].append(int(row[0]))
        loss_dict["data_loss"].append(float(row[2]))
        loss_dict["reg_loss"].append(float(row[5]))
        loss_dict["mag"].append(float(row[6]))
    return loss_dict, code[:, :latent_size].cpu()
