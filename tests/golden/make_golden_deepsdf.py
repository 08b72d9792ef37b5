"""Golden vectors of the reference's baseline decoder (networks/decoder_deepsdf.py), produced by importing
the reference's own file from /root/reference (build container only) -> tests/golden/deepsdf_forward.npz."""
import importlib
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import synth_data  # noqa: E402

REF_PY = "/root/reference/deepjoin/python"
SPECS = os.path.join(REF_PY, "..", "experiments", "mugs_deepsdf", "specs.json")


def main():
    sys.path.insert(0, REF_PY)
    arch = importlib.import_module("networks.decoder_deepsdf")
    specs = json.load(open(SPECS))
    out = {}
    for kind in ("default", "trained_norm"):
        for seed in (0, 1):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                dec = arch.Decoder(specs["CodeLength"], **specs["NetworkSpecs"]).eval()
            dec.load_state_dict(synth_data.make_deepsdf_state_dict(seed, kind))
            code = synth_data.trained_code(seed)[:128] if kind == "trained_norm" else synth_data.make_code(seed)[:128]
            pts = synth_data.make_points(2048, 20 + seed)
            with torch.no_grad():
                ref = dec(torch.cat([code.reshape(1, -1).expand(pts.shape[0], -1), pts], 1))
            out["%s_%d_code" % (kind, seed)] = code.numpy()
            out["%s_%d_pts" % (kind, seed)] = pts.numpy()
            out["%s_%d_out" % (kind, seed)] = ref.numpy()
            print(kind, seed, "sdf range", float(ref.min()), float(ref.max()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "deepsdf_forward.npz"), **out)
    reconstruct_golden(arch, specs)


def reconstruct_golden(arch, specs):
    """The reference's baseline latent loop (reconstruct_deepsdf.py:17-154) executed in place for a few
    iterations -> tests/golden/deepsdf_reconstruct.npz.  The fracture mask normally comes out of mesh ray casting
    (core.data.partial_sample_analytical, needs trimesh + ground-truth meshes): that one function is replaced by
    a stub returning a fixed synthetic mask; sampling, loss, Adam and logging are the reference's own code."""
    import importlib.util
    from oracle import ref_harness
    ns = ref_harness.load()  # installs the .cuda() shims and the reference's `core` package
    spec = importlib.util.spec_from_file_location("reconstruct_deepsdf", os.path.join(REF_PY, "reconstruct_deepsdf.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    xyz, sdf, _ = synth_data.make_sample_set(4000, seed=3)
    mask = (xyz[:, 1].astype(np.float32) > -0.15)  # "keep the rows above the fracture plane"
    ns.core.loader = lambda f: f
    ns.core_data.partial_sample_analytical = lambda broken_gt, pts, percent: mask
    rec = {"xyz": xyz, "sdf": sdf, "mask": mask}
    # the sampler alone, seeded
    np.random.seed(21)
    p_, v_ = ns.core_data.select_partial_samples(xyz, sdf, mask, num_samples=500, uniform_ratio=0.2)
    rec["sel_pts"], rec["sel_vals"] = p_, v_
    for kind in ("default", "trained_norm"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            dec = arch.Decoder(specs["CodeLength"], **specs["NetworkSpecs"]).eval()
        dec.load_state_dict(synth_data.make_deepsdf_state_dict(0, kind))
        batches = []
        orig = ns.core_data.select_partial_samples

        def recording(pts, values, mask, num_samples=None, uniform_ratio=0.5):
            p, v = orig(pts=pts, values=values, mask=mask, num_samples=num_samples, uniform_ratio=uniform_ratio)
            batches.append((p.copy(), v.copy()))
            return p, v

        ns.core_data.select_partial_samples = recording
        try:
            np.random.seed(5)
            torch.manual_seed(5)
            iters, n_s = 12, 512
            loss, lat = mod.reconstruct(dec, iters, 128, (xyz, sdf, {"broken_gt": None, "percent": 0.0}), 0.01, 0.1,
                                        num_samples=n_s, lr=5e-3, l2reg=True, code_reg_lambda=1e-4,
                                        slippage_method="analytical")
        finally:
            ns.core_data.select_partial_samples = orig
        torch.manual_seed(5)
        lat0 = torch.ones(1, 128).normal_(mean=0, std=0.01)
        rec[kind + "_code0"] = lat0.numpy()
        rec[kind + "_code"] = lat.detach().numpy()
        rec[kind + "_batch_pts"] = np.stack([b[0] for b in batches]).astype(np.float32)
        rec[kind + "_batch_sdf"] = np.stack([b[1] for b in batches]).astype(np.float32)
        for k, val in loss.items():
            rec[kind + "_log_" + k] = np.asarray(val, dtype=np.float64)
        # one explicit autograd gradient at the initial latent on the first batch
        c = lat0.clone().requires_grad_(True)
        pts = torch.from_numpy(rec[kind + "_batch_pts"][0])
        sgt = torch.clamp(torch.from_numpy(rec[kind + "_batch_sdf"][0]), -0.1, 0.1)
        pred = torch.clamp(dec(torch.cat([c.expand(pts.shape[0], -1), pts], 1)), -0.1, 0.1)
        l = torch.nn.L1Loss()(pred, sgt) + 1e-4 * torch.mean(c.pow(2))
        l.backward()
        rec[kind + "_grad0"] = c.grad.numpy()
        rec[kind + "_loss0"] = np.float64(l.item())
        print(kind, "data_loss", loss["data_loss"], "|grad0|", float(c.grad.norm()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "deepsdf_reconstruct.npz"), **rec)


if __name__ == "__main__":
    main()
