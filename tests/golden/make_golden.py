"""Generate the golden fixtures under tests/golden by EXECUTING THE REFERENCE
in place (oracle/ref_harness.py).  Run in the build container only:

    python tests/golden/make_golden.py

The fixtures are what pins ``oracle/decoder_oracle.py`` (and, on the GPU box
where /root/reference is absent, the CUDA path) to the reference's behaviour.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness, synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ns = ref_harness.load()
    torch.set_num_threads(8)

    # 1. shipped fixtures: final per-parameter norms + a few trained latent rows
    logs = torch.load(os.path.join(ref_harness.MUGS, "Logs.pth"), weights_only=False)
    mag = {k: float(v[-1]) for k, v in logs["param_magnitude"].items()}
    with open(os.path.join(OUT, "param_magnitude.json"), "w") as f:
        json.dump(mag, f, indent=1, sort_keys=True)
    codes = {}
    for name in ("2000", "2000_tool", "2000_val", "2000_tool_val"):
        t = torch.load(os.path.join(ref_harness.MUGS, "LatentCodes", name + ".pth"), weights_only=False)
        codes["c" + name] = t["latent_codes"]["weight"][:8].numpy().astype(np.float32)
    np.savez(os.path.join(OUT, "latent_rows.npz"), **codes)

    # 2. decoder forward, every branch
    fwd = {}
    for kind in ("default", "trained_norm"):
        for seed in (0, 1):
            sd = synth.make_state_dict(seed, kind)
            dec = ref_harness.make_decoder(ns, sd)
            if kind == "trained_norm":
                code = torch.from_numpy(np.concatenate([codes["c2000"][seed], codes["c2000_tool"][seed]]))
            else:
                code = synth.make_code(seed)
            pts = synth.make_points(384, seed)
            inp = torch.cat([code.expand(pts.shape[0], -1), pts], 1)
            tag = "%s_s%d" % (kind, seed)
            fwd[tag + "_code"] = code.numpy()
            fwd[tag + "_pts"] = pts.numpy()
            with torch.no_grad():
                outs = dec(inp)
                for nm, o in zip(("c_occ", "b_occ", "r_occ", "t_occ", "c_sdf", "b_sdf", "r_sdf", "t_sdf",
                                  "c_n", "b_n", "r_n", "t_n"), outs):
                    fwd[tag + "_all_" + nm] = o.numpy()
                for u in (0, 1, 2, 3):
                    fwd[tag + "_u%d" % u] = dec(inp, use_net=u).numpy()
                    fwd[tag + "_u%d_occ" % u] = dec(inp, use_net=u, return_occ=True).numpy()
    np.savez_compressed(os.path.join(OUT, "decoder_forward.npz"), **fwd)

    # 3. grid layout: decode_shape / get_query_points (cubic and non-cubic)
    grid = {}
    sd = synth.make_state_dict(0, "trained_norm")
    dec = ref_harness.make_decoder(ns, sd)
    code = torch.from_numpy(np.concatenate([codes["c2000"][0], codes["c2000_tool"][0]]))
    grid["code"] = code.numpy()
    with torch.no_grad():
        for dims in ((12, 12, 12), (6, 9, 11)):
            t = "d%dx%dx%d" % dims
            grid[t + "_pts"] = ns.core_reconstruct.get_query_points(dims, 0.1)
            for u in (0, 1, 2, 3):
                grid[t + "_u%d_flat" % u] = ns.core_reconstruct.decode_shape(
                    code, dec, dims, use_net=u, batch_size=500, padding=0.1, reshape=False, sigmoid=False)
            grid[t + "_u1_reshape_sig"] = ns.core_reconstruct.decode_shape(
                code, dec, dims, use_net=1, batch_size=2 ** 16, padding=0.1, reshape=True, sigmoid=True)
    np.savez_compressed(os.path.join(OUT, "grid.npz"), **grid)

    # 4. sampler: seeded numpy global RNG
    xyz, sdf, nrm = synth.make_sample_set(4000, seed=3)
    samp = {"xyz": xyz, "sdf": sdf, "n": nrm}
    np.random.seed(11)
    p, v = ns.core_data.select_samples(np.hstack((xyz, nrm)), sdf, 600, uniform_ratio=0.2)
    samp["sel_pts"], samp["sel_vals"] = p, v
    np.random.seed(12)
    p, v = ns.core_data.select_samples(np.hstack((xyz, nrm)), sdf, 601, uniform_ratio=0.2)
    samp["sel2_pts"], samp["sel2_vals"] = p, v
    np.savez_compressed(os.path.join(OUT, "sampler.npz"), **samp)

    # 5. latent optimisation: the reference's reconstruct() for a few iterations,
    #    recording the batches its sampler produced so both sides see identical data
    rec = {}
    for kind in ("default", "trained_norm"):
        sd = synth.make_state_dict(0, kind)
        dec = ref_harness.make_decoder(ns, sd)
        batches = []
        orig = ns.core_data.select_samples

        def recording(pts, values, num_samples=None, uniform_ratio=0.5):
            p, v = orig(pts, values, num_samples, uniform_ratio)
            batches.append((p.copy(), v.copy()))
            return p, v

        ns.core_data.select_samples = recording
        try:
            np.random.seed(5)
            torch.manual_seed(5)
            iters, ns_ = 6, 512
            loss, code = ns.reconstruct(dec, iters, 128, 64, [xyz, sdf, nrm], 0.01, 1.0, 0.0, 1.0, 0.1,
                                        num_samples=ns_, lr=5e-3, l2reg=True, code_reg_lambda=1e-4)
        finally:
            ns.core_data.select_samples = orig
        torch.manual_seed(5)
        lat0 = torch.ones(1, 128).normal_(mean=0, std=0.01)
        blat0 = torch.ones(1, 64).normal_(mean=0, std=0.01)
        rec[kind + "_code0"] = torch.cat([lat0, blat0], 1).numpy()
        rec[kind + "_code"] = code.detach().numpy()
        rec[kind + "_batch_pts"] = np.stack([b[0] for b in batches]).astype(np.float32)
        rec[kind + "_batch_sdf"] = np.stack([b[1] for b in batches]).astype(np.float32)
        for k, val in loss.items():
            rec[kind + "_log_" + k] = np.asarray(val, dtype=np.float64)
        # one explicit autograd gradient through the reference decoder
        c = torch.from_numpy(rec[kind + "_code0"]).clone().requires_grad_(True)
        pts = torch.from_numpy(rec[kind + "_batch_pts"][0][:, :3])
        ngt = torch.from_numpy(rec[kind + "_batch_pts"][0][:, 3:])
        sgt = torch.from_numpy(rec[kind + "_batch_sdf"][0])
        outs = dec(torch.cat([c.expand(pts.shape[0], -1), pts], 1))
        b_x = torch.clamp(outs[5], -0.1, 0.1)
        l = (torch.nn.L1Loss()(b_x, sgt) + torch.nn.BCELoss()(outs[1], ns.core_reconstruct.tensor_sdf_to_occ(sgt))
             + torch.nn.MSELoss()(outs[9], ngt)
             + (c[:, :128].pow(2).mean() + c[:, 128:].pow(2).mean()) * 1e-4)
        l.backward()
        rec[kind + "_grad0"] = c.grad.numpy()
        rec[kind + "_loss0"] = np.float64(l.item())
    np.savez_compressed(os.path.join(OUT, "reconstruct.npz"), **rec)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
