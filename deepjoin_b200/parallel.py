"""Multi-GPU sharding of the hot path: independent shapes per rank, or axis-0
slabs of one grid per rank with a final gather of mesh fragments.  No collective runs during
compute; the only exchange is the gather at the end (torch.distributed: NCCL on GPUs, gloo in
the CPU tests)."""
import numpy as np


def slab_ranges(n0, world):
    """Split the n0-1 cell layers of axis 0 into `world` contiguous slabs.
    Returns [(z_lo, z_hi)] sample-plane ranges, inclusive: slab k holds planes z_lo..z_hi and
    shares plane z_hi with slab k+1 (recomputed locally, never exchanged)."""
    cells = n0 - 1
    per = -(-cells // world)
    out = []
    for r in range(world):
        lo = min(r * per, cells)
        hi = min(lo + per, cells)
        out.append((lo, hi))
    return out


def shard_items(n_items, world, rank):
    """Contiguous shape shards: ceil(n/world) per rank (reference reconstruct.py:642-660 chunking)."""
    per = -(-n_items // world)
    return range(min(rank * per, n_items), min((rank + 1) * per, n_items))


def merge_slab_meshes(fragments, top_planes):
    """fragments[k] = (verts [V_k,3], faces [F_k,3]) of slab k with slab-local ids; faces of slabs
    k>0 hold -2-slot for vertices on the seam plane; top_planes[k] = ids (local to slab k) of the
    edges in slab k's last plane, 3 per sample.  Concatenation in slab order with the exclusive
    scan of vertex counts reproduces the single-sweep ids exactly."""
    verts, faces = [], []
    offset = 0
    prev_top, prev_offset = None, 0
    for k, (v, f) in enumerate(fragments):
        f = np.asarray(f, dtype=np.int64).copy()
        if k > 0:
            seam = f < -1
            slots = -2 - f[seam]
            ids = np.asarray(prev_top, dtype=np.int64)[slots]
            if (ids < 0).any():
                raise RuntimeError("slab %d references a seam edge the slab below never created" % k)
            f[seam] = ids + prev_offset
            f[~seam] += offset
        verts.append(np.asarray(v))
        faces.append(f)
        prev_top, prev_offset = top_planes[k], offset
        offset += len(v)
    return np.concatenate(verts, 0), np.concatenate(faces, 0).astype(np.int32)
