"""Stage-by-stage GPU diagnostics, each stage in its own subprocess (a device trap in one stage must
not poison the others).  Writes gpurun_out/diag.json.  Usage: python tools/gpu_diag.py [stage ...]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


# the dj_debug_* entry points live in the diagnostics build only (python -m deepjoin_b200.build --debug)
os.environ.setdefault("DEEPJOIN_B200_LIB", os.path.join(ROOT, "deepjoin_b200", "lib", "libdeepjoin_b200_debug.so"))


def bf16_round(a):
    import numpy as np
    u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def stage_probe():
    import ctypes
    import numpy as np
    from deepjoin_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(0)
    res = {}
    for K in (64, 256):
        A = rng.standard_normal((128, K)).astype(np.float32)
        W = rng.standard_normal((128, K)).astype(np.float32)
        ref = bf16_round(A).astype(np.float64) @ bf16_round(W).astype(np.float64).T
        for bulk in (0, 1, 3):
            D = np.zeros((128, 128), np.float32)
            rc = L.dj_debug_umma_probe(A.ctypes.data, W.ctypes.data, K, bulk, 0, 0, 0, D.ctypes.data)
            key = "K%d_bulk%d" % (K, bulk)
            if rc:
                res[key] = {"rc": rc, "err": _lib.last_error()}
                continue
            err = np.abs(D - ref)
            res[key] = {"rc": 0, "max_err": float(err.max()), "ref_absmax": float(np.abs(ref).max()),
                        "frac_close": float((err < 1e-2 * max(1.0, np.abs(ref).max())).mean()),
                        "nan": int(np.isnan(D).sum())}
            if err.max() > 1e-2:
                # pattern hints: which rows / columns are right
                ok = err < 1e-2
                res[key]["rows_ok"] = int(ok.all(axis=1).sum())
                res[key]["cols_ok"] = int(ok.all(axis=0).sum())
                res[key]["D00"] = [float(x) for x in D[0, :4]]
                res[key]["ref00"] = [float(x) for x in ref[0, :4]]
    return res


def _fwd_errors(precision, kind, n=8192):
    import numpy as np
    import torch
    from oracle import decoder_oracle as O, synth
    from tests import util
    dec = util.make_decoder(0, kind, precision=precision)
    net = util.oracle_net(0, kind, "float64")
    code = util.code_for(0, kind)
    pts = synth.make_points(n, 1)
    out = {}
    for u in (0, 3, 1, 2):
        ref = O.forward(net, code.double(), pts.double(), u).numpy()
        t0 = time.time()
        got = dec.query(code.cuda(), pts.cuda(), use_net=u).cpu().numpy()
        e = np.abs(got - ref)
        out["u%d" % u] = {"max": float(e.max()), "mean": float(e.mean()), "nan": int(np.isnan(got).sum()),
                          "ref_std": float(ref.std()), "t": time.time() - t0}
    ref = [o.numpy() for o in O.forward_all(net, code.double(), pts.double())]
    got = [o.cpu().numpy() for o in dec.query_all(code.cuda(), pts.cuda())]
    out["all20_max"] = [float(np.abs(g - r).max()) for g, r in zip(got, ref)]
    return out


def stage_fp32():
    return {k: _fwd_errors("fp32", k) for k in ("default", "trained_norm")}


def stage_tc():
    return {k: _fwd_errors("bf16", k) for k in ("default", "trained_norm")}


def stage_fp16():
    """IEEE-half operands on the pair kernel (DJ_PREC_FP16) vs bf16: error against the fp64 oracle."""
    return {"fp16": {k: _fwd_errors("fp16", k) for k in ("default", "trained_norm")},
            "bf16": {k: _fwd_errors("bf16", k) for k in ("default", "trained_norm")}}


def stage_mc():
    import numpy as np
    import torch
    from oracle import mc_oracle
    from deepjoin_b200.core import reconstruct as R
    res = {}
    rng = np.random.default_rng(0)
    for name, vol in (("noise", rng.uniform(-1, 1, (20, 21, 22)).astype(np.float32)),
                      ("sphere", (np.sqrt(((np.indices((48, 48, 48)) - 23.5) ** 2).sum(0)) - 15).astype(np.float32))):
        vo, fo = mc_oracle.marching_cubes(vol, 0.0, "ascent")
        v, f = R.marching_cubes(torch.from_numpy(vol).cuda(), 0.0, "ascent")
        v, f = v.cpu().numpy(), f.cpu().numpy()
        res[name] = {"nv": [len(v), len(vo)], "nf": [len(f), len(fo)],
                     "faces_equal": bool(f.shape == fo.shape and np.array_equal(f, fo)),
                     "verts_equal": bool(v.shape == vo.shape and np.array_equal(v.view(np.uint32), vo.view(np.uint32)))}
        if v.shape == vo.shape and not res[name]["verts_equal"]:
            res[name]["vert_max_diff"] = float(np.abs(v - vo).max())
    return res


def stage_latent():
    import numpy as np
    import torch
    from oracle import decoder_oracle as O, synth
    from tests import util
    from deepjoin_b200 import reconstruct as RR
    res = {}
    xyz, sdf, nrm = synth.make_sample_set(4000, seed=1)
    x = torch.from_numpy(xyz[:1000].astype(np.float32))
    s = torch.from_numpy(sdf[:1000].astype(np.float32))
    n = torch.from_numpy(nrm[:1000].astype(np.float32))
    for kind in ("default", "trained_norm"):
        dec = util.make_decoder(0, kind, precision="fp32")
        code = util.code_for(0, kind)
        g, t = RR.latent_grad(dec, code.cuda(), x.cuda(), s.cuda(), n.cuda())
        gr, tr = O.latent_grad(util.oracle_net(0, kind), code, x, s, n)
        res[kind] = {"rel_err": float((g.cpu() - gr).abs().max() / gr.abs().max()), "terms": [float(v) for v in t.cpu()],
                     "terms_ref": tr}
    return res


def stage_timing():
    """First timings: grid query at 128^3 / 256^3 per precision, MC, latent step."""
    import torch
    from tests import util
    from deepjoin_b200.core import reconstruct as R
    res = {}
    code = util.trained_code(0)
    for precision, dimlist in (("bf16", (128, 256)), ("fp32", (64,))):
        dec = util.make_decoder(0, "trained_norm", precision=precision)
        for d in dimlist:
            for u in (1, 0):
                R.decode_shape_device(code, dec, (d, d, d), use_net=u, padding=0.1, sigmoid=False)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                vol = R.decode_shape_device(code, dec, (d, d, d), use_net=u, padding=0.1, sigmoid=False)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                res["%s_grid%d_u%d" % (precision, d, u)] = {"ms": ms, "q_per_s": d ** 3 / ms * 1e3}
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            R.marching_cubes(vol.reshape(d, d, d), 0.0, "ascent")
            e0.record()
            v, f = R.marching_cubes(vol.reshape(d, d, d), 0.0, "ascent")
            e1.record()
            torch.cuda.synchronize()
            res["%s_mc%d" % (precision, d)] = {"ms": e0.elapsed_time(e1), "nv": int(v.shape[0]), "nf": int(f.shape[0])}
    return res


def stage_latent_tc():
    """Tensor-core fused forward+backward (latent_tc2.cuh) against the fp32 path and the autograd oracle."""
    import numpy as np
    import torch
    from oracle import decoder_oracle as O, synth
    from tests import util
    from deepjoin_b200 import reconstruct as RR
    res = {}
    xyz, sdf, nrm = synth.make_sample_set(40000, seed=1)
    for kind in ("default", "trained_norm"):
        for n in (100, 1000, 30000):
            x = torch.from_numpy(xyz[:n].astype(np.float32)).cuda()
            s = torch.from_numpy(sdf[:n].astype(np.float32)).cuda()
            m = torch.from_numpy(nrm[:n].astype(np.float32)).cuda()
            dec = util.make_decoder(0, kind, precision="fp32")
            code = util.code_for(0, kind)
            g32, t32 = RR.latent_grad(dec, code.cuda(), x, s, m, precision="fp32")
            g16, t16 = RR.latent_grad(dec, code.cuda(), x, s, m, precision="bf16")
            torch.cuda.synchronize()
            g32, g16 = g32.cpu().double(), g16.cpu().double()
            r = {"rel_err_vs_fp32": float((g16 - g32).abs().max() / g32.abs().max()),
                 "cos": float((g16 * g32).sum() / (g16.norm() * g32.norm())),
                 "nan": int(torch.isnan(g16).sum()), "gnorm32": float(g32.norm()), "gnorm16": float(g16.norm()),
                 "terms32": [float(v) for v in t32.cpu()], "terms16": [float(v) for v in t16.cpu()]}
            if n <= 1000:
                gr, tr = O.latent_grad(util.oracle_net(0, kind), code, x.cpu(), s.cpu(), m.cpu())
                r["rel_err_vs_oracle"] = float((g16 - gr.double()).abs().max() / gr.abs().max())
            res["%s_n%d" % (kind, n)] = r
    return res


def stage_latent_timing():
    """Config 3 per-shape cost: 30,000 samples / iteration, device-resident loop (dj_latent_optimize)."""
    import numpy as np
    import torch
    from oracle import synth
    from tests import util
    from deepjoin_b200 import reconstruct as RR
    xyz, sdf, nrm = synth.make_sample_set(200000, seed=1)
    dec = util.make_decoder(0, "trained_norm", precision="fp32")
    lat_prec = os.environ.get("LATENT_PRECISION", "fp32")
    dev = torch.device("cuda", 0)
    px = torch.from_numpy(xyz.astype(np.float32)).to(dev)
    ps = torch.from_numpy(sdf.astype(np.float32).reshape(-1)).to(dev)
    pn = torch.from_numpy(nrm.astype(np.float32)).to(dev)
    res = {}
    for iters, ns in ((20, 30000), (100, 30000), (100, 8000)):
        sched = torch.randint(0, px.shape[0], (iters, ns), device=dev, dtype=torch.int32)
        code0 = RR.init_code(128, 64, 0.01)
        RR.optimize_code(dec, code0, px, ps, pn, sched[:2], 5e-3, 1.0, 1.0, 0.1, l2reg=True, precision=lat_prec)
        torch.cuda.synchronize()
        t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        log, code = RR.optimize_code(dec, code0, px, ps, pn, sched, 5e-3, 1.0, 1.0, 0.1, l2reg=True, precision=lat_prec)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        res["it%d_n%d" % (iters, ns)] = {"ms_per_iter": ms / iters, "wall_ms_per_iter": 1e3 * (time.time() - t0) / iters,
                                         "tflops_alg": ns * 11.03e6 / (ms / iters * 1e-3) / 1e12,
                                         "loss_first_last": [float(log[0, 1]), float(log[-1, 1])]}
    return res


def stage_prof():
    """In-kernel cycle counters of the CTA-pair forward (DEEPJOIN_B200_DEBUG |= 8): where each warp role waits."""
    import numpy as np
    import torch
    from tests import util
    from deepjoin_b200 import _lib
    from deepjoin_b200.core import reconstruct as R
    dbg = int(os.environ.get("DEEPJOIN_B200_DEBUG", "0")) | 8
    os.environ["DEEPJOIN_B200_DEBUG"] = str(dbg)
    d = int(os.environ.get("PROF_DIMS", "256"))
    dec = util.make_decoder(0, "trained_norm", precision=os.environ.get("PROF_PREC", "bf16"))
    code = util.trained_code(0)
    for _ in range(2):
        R.decode_shape_device(code, dec, (d, d, d), use_net=1, padding=0.1, sigmoid=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    R.decode_shape_device(code, dec, (d, d, d), use_net=1, padding=0.1, sigmoid=False)
    e1.record()
    torch.cuda.synchronize()
    pack = dec.pack(torch.device("cuda", 0))
    nsm = torch.cuda.get_device_properties(0).multi_processor_count
    buf = np.zeros((nsm, 16), np.uint64)
    _lib.check(_lib.lib().dj_debug_read_prof(pack.handle, buf.ctypes.data, buf.size))
    names = ["mma_total", "mma_wait_acc_empty", "mma_wait_a_ready", "mma_wait_w_full", "epi_total", "epi_wait_acc_full",
             "epi_wait_bar", "epi_l0", "epi_wait_head", "prod_total", "prod_wait_w_empty", "relay_total", "relay_wait", "epi_w4_wait_acc_full"]
    # three-pass fp16x2 program: epi_wait_bar = R_FREE(1) wait of warp 2, epi_l0 = R_FREE(0) wait of warp 4
    lead, peer = buf[0::2].astype(np.float64), buf[1::2].astype(np.float64)
    out = {"ms": e0.elapsed_time(e1), "debug": dbg}
    for i, nm in enumerate(names):
        out[nm] = {"leader_mean": float(lead[:, i].mean()), "peer_mean": float(peer[:, i].mean()),
                   "leader_max": float(lead[:, i].max())}
    return out


STAGES = {"fp16": stage_fp16, "prof": stage_prof, "latent_timing": stage_latent_timing, "latent_tc": stage_latent_tc, "probe": stage_probe, "fp32": stage_fp32, "mc": stage_mc, "latent": stage_latent, "tc": stage_tc,
          "timing": stage_timing}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        try:
            print("DIAG_RESULT " + json.dumps(STAGES[sys.argv[2]]()))
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            print("DIAG_RESULT " + json.dumps({"exception": repr(e)}))
        sys.exit(0)
    os.makedirs(OUT, exist_ok=True)
    stages = sys.argv[1:] or list(STAGES)
    results = {}
    for st in stages:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", st], capture_output=True, text=True,
                               timeout=300, cwd=ROOT)
            line = [l for l in p.stdout.splitlines() if l.startswith("DIAG_RESULT ")]
            results[st] = json.loads(line[-1][12:]) if line else {"rc": p.returncode, "stderr": p.stderr[-2000:], "stdout": p.stdout[-500:]}
            if line and p.stderr.strip() and "exception" in results[st]:
                results[st]["stderr"] = p.stderr[-2000:]
        except subprocess.TimeoutExpired:
            results[st] = {"timeout": True}
        results[st]["_sec"] = round(time.time() - t0, 1)
        print(st, json.dumps(results[st])[:1500], flush=True)
        json.dump(results, open(os.path.join(OUT, "diag.json"), "w"), indent=1)
