"""Stage times of the slab-sharded create_mesh (BASELINE configs[3]) on every rank: slab query + marching cubes,
the statistics all-reduce, the fragment exchange, the id merge on rank 0, the mesh finish (vertex merge + host
copy).  Run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/slab_breakdown.py --dims 512 --precision bf16

Each stage is bracketed by a device synchronisation, so the sum is an upper bound of the pipelined call; the
last line is the un-instrumented call for comparison."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, default=512)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bench
    import synth_data
    from deepjoin_b200 import parallel as Par

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dec = bench.make_decoder(0, dev, args.precision)
    code = synth_data.trained_code(0).to(dev)
    dims = (args.dims,) * 3
    kw = dict(sigmoid=False, level=0.0, gradient_direction="ascent", dims=dims, padding=0.1)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return time.perf_counter()

    for _ in range(2):
        m = None
        m = Par.create_mesh_sharded(code, dec, 1, **kw)
    for rep in range(args.reps):
        m = None
        t = [sync()]
        z_lo, z_hi = Par.slab_ranges(dims[0], world)[rank]
        verts, faces, top, (vmin, vmax) = Par.slab_fragment(code, dec, 1, False, 0.0, dims, 0.1, z_lo, z_hi,
                                                            seam=rank > 0 and z_lo > 0)
        t.append(sync())
        stats = torch.stack([-vmin, vmax, torch.tensor(float(verts.shape[0]), device=dev)]).to(torch.float64)
        if world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        stats.tolist()
        t.append(sync())
        got = Par.gather_fragments(verts, faces, top, merged=True)
        t.append(sync())
        if got is not None:
            v, f = Par.merge_gathered(*got)
        t.append(sync())
        if got is not None:
            m = Par.finish_mesh(v, f, dims, 0.1, "ascent")
        t.append(sync())
        if rank == 0:
            names = ["query+mc", "stats", "gather", "merge ids", "finish"]
            print("rep %d: " % rep + ", ".join("%s %.2f ms" % (n, 1e3 * (b - a)) for n, a, b in zip(names, t, t[1:]))
                  + " | sum %.2f ms; %d vertices" % (1e3 * (t[-1] - t[0]), len(m.vertices)), flush=True)
        del verts, faces, top, got
    for rep in range(args.reps):
        m = None
        t0 = sync()
        m = Par.create_mesh_sharded(code, dec, 1, **kw)
        t1 = sync()
        if rank == 0:
            print("whole call: %.2f ms" % (1e3 * (t1 - t0)), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
