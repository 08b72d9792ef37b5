#!/usr/bin/env python
"""Benchmark of the DeepJoin inference hot path on B200 (contract: one JSON line on rank 0).

Metric (BASELINE.json): decoder queries/s at a 256^3 grid.
Workload at every N: BASELINE configs[1] per GPU -- one shape, dense 256^3 grid query of the joint
decoder (use_net=1: B = max(C, -T), both MLPs) + marching cubes; with N > 1 every rank runs its own
shape (independent shapes shard with no data-path collective -> weak scaling).
A step = one grid query + one marching-cubes extraction.

  value     queries/s with everything resident in HBM (device-timed, max over ranks)
  e2e       the same through core.reconstruct.create_mesh (host code in, host vertices/faces out)
  roofline  tensor-core bound of the fused decoder kernel: algorithmic FLOP (5,518,336 / query,
            SURVEY.md 8d) / CUDA-event time of the grid-query launches
  cpu_baseline  the oracle's torch-fp32 restatement of the reference decoder on the host cores,
            bounded sample (rank 0, N=1)

`--impl reference` times that CPU path alone (the reference is pure Python/PyTorch; its CPU
implementation of the path is what oracle/decoder_oracle.py restates and pins).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_ALG = {0: 3412992, 3: 2105344, 1: 5518336, 2: 5518336}
METRIC = "decoder_queries_per_sec_256cubed_grid"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


SPECS = {  # experiments/mugs/specs.json of the reference
    "NetworkSpecs": {"dims": [512] * 8, "dropout": list(range(8)), "dropout_prob": 0.2, "norm_layers": list(range(8)),
                     "latent_in": [4], "xyz_in_all": False, "use_tanh": False, "latent_dropout": False,
                     "weight_norm": True},
    "SubnetSpecs": {"subnet_dims": [512] * 5, "subnet_dropout": list(range(5)), "subnet_norm": list(range(5)),
                    "subnet_xyz": True},
}


def make_decoder(seed, device, precision):
    """Product decoder with synthetic trained-norm weights (synth_data.py; no oracle code involved)."""
    import synth_data
    from deepjoin_b200.networks.decoder_z_lb_both_leaky_n_sdf import Decoder
    dec = Decoder(latent_size=128, tool_latent_size=64, num_dims=3, do_code_regularization=True,
                  **SPECS["NetworkSpecs"], **SPECS["SubnetSpecs"])
    dec.load_state_dict(synth_data.make_state_dict(seed, "trained_norm"))
    dec.precision = precision
    return dec.to(device).eval()


def cpu_decoder_rate(n_points, use_net=1, threads=None):
    """THE ONLY PLACE bench.py touches oracle/: the oracle port of the reference decoder (torch fp32 on the
    host cores), timed as the CPU baseline / reference arm."""
    import torch
    import synth_data
    from oracle import decoder_oracle as O
    if threads:
        torch.set_num_threads(threads)
    net = O.fold(synth_data.make_state_dict(0, "trained_norm"), dtype=torch.float32)
    code = synth_data.trained_code(0)
    pts = synth_data.make_points(min(n_points, 1 << 16), 0)
    with torch.no_grad():
        O.forward(net, code, pts[:4096], use_net)  # warm-up
        done, t0 = 0, time.perf_counter()
        while done < n_points:
            O.forward(net, code, pts, use_net)
            done += pts.shape[0]
        dt = time.perf_counter() - t0
    return done / dt, done, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    # all host threads, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank)
    torch.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    sample = 1 << 17
    rates = []
    for _ in range(args.warmup):
        cpu_decoder_rate(1 << 14)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        r, done, dt, thr = cpu_decoder_rate(sample)
        rates.append((done, dt))
    tot = sum(d for d, _ in rates)
    sec = sum(t for _, t in rates)
    val = tot / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: 256^3 grid query, use_net=1 (both MLPs), bounded sample of %d points per step "
                               "on the host cores" % sample},
        "cpu_baseline": {"value": val, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "%d grid-range points per step, torch fp32 restatement of Decoder.forward "
                                   "(oracle/decoder_oracle.py, pinned to the reference by tests/golden)" % sample},
        "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dims", type=int, default=256)
    ap.add_argument("--use-net", type=int, default=1)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--query-only", action="store_true", help="diagnostics: time the grid query alone (not a bench line)")
    ap.add_argument("--latent-shapes", type=int, default=8,
                    help="shapes per GPU of the second half of the metric (BASELINE configs[2]: 800 Adam iterations x 30,000 "
                         "samples per shape, then one 256^3 restoration mesh); 0 disables")
    ap.add_argument("--slab-dims", type=int, default=0,
                    help="also time BASELINE configs[3]: ONE grid of this size sharded by axis-0 slabs over the ranks, "
                         "fragments gathered to rank 0 over NCCL (adds a 'slab' object to the JSON line)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from deepjoin_b200 import _lib
    import synth_data
    from deepjoin_b200.core import reconstruct as R

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    d = args.dims
    dims = (d, d, d)
    npts = d ** 3
    u = args.use_net
    # every rank its own shape: synthetic weights seed r are centred on the shipped latent row r (zero level set exists)
    dec = make_decoder(rank % 8, dev, args.precision)
    code = synth_data.trained_code(rank % 8)
    code_dev = code.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def device_step(ev=None):
        flush.fill_(1)
        if ev:
            ev[0].record()
        vol = R.decode_shape_device(code_dev, dec, dims, use_net=u, padding=0.1, sigmoid=False)
        if ev:
            ev[1].record()
        if args.query_only:
            if ev:
                ev[2].record()
            return vol, vol
        v, f = R.marching_cubes(vol.reshape(dims), 0.0, "ascent")
        if ev:
            ev[2].record()
        return v, f

    for _ in range(max(args.warmup, 3)):
        v, f = device_step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    _lib.lib().dj_launch_count(1)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for s in range(args.steps):
        v, f = device_step(evs[s])
    t1.record()
    barrier()
    launches = int(_lib.lib().dj_launch_count(0))
    ms = t0.elapsed_time(t1)
    q_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    mc_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    nv, nf = int(v.shape[0]), int(f.shape[0])

    if args.query_only:
        if rank == 0:
            print(json.dumps({"diagnostic": "query-only", "grid_query_ms": q_ms, "queries_per_s": npts / (q_ms * 1e-3),
                              "tflops_alg": npts * FLOP_ALG[u] / (q_ms * 1e-3) / 1e12, "clocks": sampler.stop()}))
        return
    # ---- end to end through the public API: host code in, host mesh out
    for _ in range(2):
        mesh = R.create_mesh(code, dec, u, sigmoid=False, level=0.0, gradient_direction="ascent", dims=dims, padding=0.1)
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        mesh = R.create_mesh(code, dec, u, sigmoid=False, level=0.0, gradient_direction="ascent", dims=dims, padding=0.1)
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - w0)
    clocks = sampler.stop() if sampler else None

    latent = None
    if args.latent_shapes and not args.query_only:
        from deepjoin_b200 import reconstruct as RR
        n_sh, iters_l, ns_l = args.latent_shapes, 800, 30000
        pools = []
        for sh in range(n_sh):  # every shape its own synthetic fractured-mug sample set (resident in HBM)
            xyz_s, sdf_s, nrm_s = synth_data.make_sample_set(100000, seed=rank * n_sh + sh)
            pools.append((torch.from_numpy(xyz_s.astype(np.float32)).to(dev), torch.from_numpy(sdf_s.astype(np.float32).reshape(-1)).to(dev),
                          torch.from_numpy(nrm_s.astype(np.float32)).to(dev)))
        sched = torch.empty((n_sh, iters_l, ns_l), device=dev, dtype=torch.int32)

        def run_shapes(n_it):
            """configs[2] on this rank: draw every shape's mini-batch schedule on the device, optimise all codes with
            one fused launch per Adam iteration, then extract each shape's restoration mesh."""
            w_s = time.perf_counter()
            for sh in range(n_sh):
                _lib.check(_lib.lib().dj_sample_schedule(pools[sh][1].data_ptr(), pools[sh][1].shape[0], n_it, ns_l, 0.2,
                                                         1000 + rank * n_sh + sh, sched[sh].data_ptr(), _lib.stream_ptr()))
            torch.cuda.synchronize()
            run_shapes.sched_ms = 1e3 * (time.perf_counter() - w_s)
            codes0 = []
            for sh in range(n_sh):
                torch.manual_seed(rank * n_sh + sh)
                codes0.append(RR.init_code(128, 64, 0.01))
            t_a = torch.cuda.Event(enable_timing=True)
            t_b = torch.cuda.Event(enable_timing=True)
            t_a.record()
            logs, codes = RR.optimize_codes(dec, torch.cat(codes0), pools, sched[:, :n_it], 5e-3, 1.0, 1.0, 0.1, l2reg=True,
                                            code_reg_lambda=1e-4, precision=args.precision)
            t_b.record()
            torch.cuda.synchronize()
            run_shapes.opt_ms = t_a.elapsed_time(t_b)
            n_ok = 0
            w_m = time.perf_counter()
            for sh in range(n_sh):
                try:
                    R.create_mesh(codes[sh], dec, 2, sigmoid=False, level=0.0, gradient_direction="ascent", dims=dims, padding=0.1)
                    n_ok += 1
                except Exception:  # IsosurfaceExtractionError -> empty mesh, as the reference's mesh_chunk does
                    pass
            run_shapes.mesh_ms = 1e3 * (time.perf_counter() - w_m)
            return logs, n_ok

        run_shapes(10)
        barrier()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wl = time.perf_counter()
        l0.record()
        logs, n_mesh = run_shapes(iters_l)
        l1.record()
        torch.cuda.synchronize()
        tl = torch.tensor([l0.elapsed_time(l1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        if rank == 0:
            latent = {"workload": "configs[2]: %d shapes per GPU x %d Adam iterations x %d samples (device sampler, all shapes of a "
                                  "rank in one fused forward+backward tensor-core launch per iteration), then one %d^3 restoration "
                                  "mesh per shape" % (n_sh, iters_l, ns_l, d),
                      "shapes_per_s": world * n_sh / (float(tl[0]) * 1e-3), "ms_per_shape": float(tl[0]) / n_sh,
                      "wall_ms_per_shape": 1e3 * (time.perf_counter() - wl) / n_sh, "meshes_with_surface": n_mesh,
                      "adam_loop_ms_per_shape": run_shapes.opt_ms / n_sh, "schedule_ms_per_shape": run_shapes.sched_ms / n_sh,
                      "mesh_ms_per_shape": run_shapes.mesh_ms / n_sh,
                      "adam_loop_tflops_alg": n_sh * iters_l * ns_l * 11.03e6 / (run_shapes.opt_ms * 1e-3) / 1e12,
                      "loss_first_last_shape0": [float(logs[0, 0, 1]), float(logs[0, -1, 1])], "precision": args.precision}
    slab = None
    if args.slab_dims:
        from deepjoin_b200 import parallel as Par
        sd = (args.slab_dims,) * 3
        dec0 = make_decoder(0, dev, args.precision)  # same weights on every rank
        code0 = synth_data.trained_code(0).to(dev)
        for _ in range(2):
            m = Par.create_mesh_sharded(code0, dec0, u, sigmoid=False, level=0.0, gradient_direction="ascent", dims=sd, padding=0.1)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        s0.record()
        reps = 3
        for _ in range(reps):
            m = Par.create_mesh_sharded(code0, dec0, u, sigmoid=False, level=0.0, gradient_direction="ascent", dims=sd, padding=0.1)
            barrier()
        s1.record()
        torch.cuda.synchronize()
        ts = torch.tensor([s0.elapsed_time(s1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        if rank == 0:
            slab = {"workload": "configs[3]: %d^3 grid sharded by axis-0 slabs over %d GPU(s), NCCL gather of mesh fragments, "
                                "mesh finished on rank 0 (host Mesh out)" % (args.slab_dims, world),
                    "ms_per_mesh": float(ts[0]), "wall_ms_per_mesh": 1e3 * (time.perf_counter() - w0) / reps,
                    "queries_per_s": args.slab_dims ** 3 / (float(ts[0]) * 1e-3),
                    "vertices": int(len(m.vertices)), "faces": int(len(m.faces))}
    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    if rank == 0:
        burst, sustained, hbm, which = peaks()
        value = world * npts * args.steps / (ms * 1e-3)
        achieved = npts * FLOP_ALG[u] / (q_ms * 1e-3) / 1e12
        peak = sustained
        line = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"bf16": "bf16", "fp16": "f16", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": "configs[1]: single-shape dense query + marching cubes at %d^3 grid per GPU, "
                                   "use_net=%d, padding 0.1, level 0 ascent; synthetic trained-norm weights" % (d, u),
                       "l2": "256 MB flush buffer written before every step (inside the timed bracket)",
                       "grid_query_ms": q_ms, "marching_cubes_ms": mc_ms, "vertices": nv, "faces": nf,
                       "precision": args.precision},
            "e2e": {"value": world * npts * args.steps / (e2e_ms * 1e-3), "unit": "queries/s",
                    "h2d_bytes_per_step": int(code.numel() * 4), "d2h_bytes_per_step": int(mesh.vertices.nbytes + mesh.faces.nbytes),
                    "ms_per_step": e2e_ms / args.steps, "api": "deepjoin_b200.core.reconstruct.create_mesh"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one 256^3 launch (ncu --set full,
                         # profiles/r01_tc2_forward_256_ncu_full.txt): 6.3 MB of weights + 14.0 MB of the 67 MB
                         # output volume (the rest is still in L2 when the kernel ends)
                         "traffic": 20361728 if (args.precision != "fp32" and d == 256) else None,
                         "kernel": "tc2_forward_kernel" if args.precision != "fp32" else "sgemm_kernel",
                         "peak_source": "bf16_tflops_sustained of %s (kernel timed inside a long step); burst %.1f" % (which, burst),
                         "flop_per_query": FLOP_ALG[u]},
            "clocks": clocks,
        }
        if latent:
            line["latent"] = latent
        if slab:
            line["slab"] = slab
        if not args.no_cpu_baseline and world == 1:
            rate, done, dt, thr = cpu_decoder_rate(1 << 19, u)
            line["cpu_baseline"] = {"value": rate, "unit": "queries/s", "cores": thr, "kind": "port",
                                    "sample": "%d points (%.1f s), torch fp32 restatement of the reference Decoder.forward, "
                                              "use_net=%d" % (done, dt, u)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
