"""Sharded path on the CPU: shard arithmetic, fragment merge, and the
torch.distributed gather with world_size 2 and 3 over gloo.  Fragments come from the C oracle
re-expressed in the GPU kernels' seam convention (tests/slab_util.py)."""
import os
import socket

import numpy as np
import pytest
import torch

from deepjoin_b200 import parallel as P
from oracle import mc_oracle
from tests import slab_util


def _volume(shape=(13, 9, 11), seed=3):
    return np.random.default_rng(seed).uniform(-1, 1, shape).astype(np.float32)


def test_shard_arithmetic():
    assert P.slab_ranges(512, 8)[0] == (0, 64) and P.slab_ranges(512, 8)[-1] == (448, 511)
    for n0 in (2, 5, 33, 256):
        for world in (1, 2, 3, 8):
            r = P.slab_ranges(n0, world)
            assert r[0][0] == 0 and r[-1][1] == n0 - 1 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    assert [list(P.shard_items(64, 8, k)) for k in (0, 7)] == [list(range(8)), list(range(56, 64))]
    assert sum(len(P.shard_items(10, 4, k)) for k in range(4)) == 10


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_fragment_merge_reproduces_single_sweep(world):
    vol = _volume()
    vo, fo = mc_oracle.marching_cubes(vol, 0.0, "ascent")
    frags = [slab_util.oracle_slab_fragment(vol, lo, hi, seam=k > 0 and lo > 0)
             for k, (lo, hi) in enumerate(P.slab_ranges(vol.shape[0], world))]
    v, f = P.merge_slab_meshes([(a, b) for a, b, _ in frags], [c for _, _, c in frags])
    # ids / order exactly; coordinates to rounding (the CPU stand-in adds the slab offset after the fact,
    # the GPU kernels interpolate in global index units and are bit-exact: tests/test_gpu_mc.py)
    assert np.array_equal(f, fo) and np.allclose(v, vo, rtol=0, atol=2e-6)
    tv, tf = P.merge_fragments([tuple(torch.from_numpy(x) for x in fr) for fr in frags])
    assert np.array_equal(tf.numpy(), fo) and np.array_equal(tv.numpy(), v)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vol = _volume()
        lo, hi = P.slab_ranges(vol.shape[0], world)[rank]
        v, f, top = slab_util.oracle_slab_fragment(vol, lo, hi, seam=rank > 0 and lo > 0)
        frags = P.gather_fragments(torch.from_numpy(v), torch.from_numpy(f), torch.from_numpy(top), dst=0)
        # shape sharding: every rank's share of 10 items, gathered as in the batch driver
        mine = torch.tensor(list(P.shard_items(10, world, rank)), dtype=torch.int64)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([len(mine)]))
        got = P.gather_fragments(torch.from_numpy(v), torch.from_numpy(f), torch.from_numpy(top), dst=0, merged=True)
        if rank == 0:
            assert frags is not None and len(frags) == world
            mv, mf = P.merge_fragments([(a, b, c if c is not None else torch.zeros(0, dtype=torch.int32)) for a, b, c in frags])
            gv, gf = P.merge_gathered(*got)  # the path create_mesh_sharded takes: merged in place
            assert torch.equal(gv, mv) and torch.equal(gf, mf)
            assert got[3][-1] is None  # the last slab's top-plane table is never sent
            mesh = P.finish_mesh(mv, mf, vol.shape, 0.1, "descent")
            np.savez(out_path, v=mv.numpy(), f=mf.numpy(), mesh_v=mesh.vertices, mesh_f=mesh.faces,
                     sizes=np.array([int(s) for s in sizes]))
        else:
            assert frags is None and got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_gather_of_mesh_fragments(world, tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)
    vol = _volume()
    vo, fo = mc_oracle.marching_cubes(vol, 0.0, "ascent")
    assert np.array_equal(got["f"], fo) and np.allclose(got["v"], vo, rtol=0, atol=2e-6)
    assert got["sizes"].sum() == 10
    # the finished mesh equals the oracle's create_mesh post-processing of the single sweep
    from deepjoin_b200.mesh import Mesh
    rv, rf = mc_oracle.mesh_from_volume(vol, 0.0, vol.shape, 0.1, "descent")
    ref = Mesh(rv, rf)
    assert np.array_equal(got["mesh_f"], ref.faces) and np.allclose(got["mesh_v"], ref.vertices, rtol=0, atol=1e-6)
