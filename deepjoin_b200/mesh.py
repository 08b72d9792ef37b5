"""Minimal mesh container standing in for ``trimesh.Trimesh`` at the create_mesh boundary
(reference core/reconstruct.py:189-191, core/utils_multiprocessing.py:97,107).

trimesh is not a dependency of the hot path; what its callers use is ``.vertices``,
``.faces``, ``.export(path)`` and emptiness.  ``process=True`` (trimesh's default) merges
vertices that coincide after rounding to 1e-8 (``trimesh.constants.tol.merge``) and drops
NaN/Inf vertices; the merge here keeps first-occurrence order (trimesh's own order comes
from a hashed sort and is not reproduced -- SURVEY.md §8 f3)."""
import os
import struct

import numpy as np


class Mesh(object):
    def __init__(self, vertices=None, faces=None, process=True):
        if vertices is None:
            vertices = np.zeros((0, 3), np.float64)
        if faces is None:
            faces = np.zeros((0, 3), np.int64)
        self.vertices = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
        self.faces = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
        if process and len(self.vertices):
            self.remove_infinite_values()
            self.merge_vertices()

    @property
    def is_empty(self):
        return len(self.vertices) == 0 or len(self.faces) == 0

    def remove_infinite_values(self):
        ok = np.isfinite(self.vertices).all(axis=1)
        if ok.all():
            return
        remap = np.cumsum(ok) - 1
        keep = ok[self.faces].all(axis=1)
        self.faces = remap[self.faces[keep]]
        self.vertices = self.vertices[ok]

    def merge_vertices(self, digits=8):
        key = np.round(self.vertices, digits)
        _, first, inverse = np.unique(key, axis=0, return_index=True, return_inverse=True)
        inverse = inverse.reshape(-1)
        if len(first) == len(self.vertices):
            return
        order = np.argsort(first)              # unique ids in first-occurrence order
        rank = np.empty_like(order)
        rank[order] = np.arange(len(order))
        self.vertices = self.vertices[first[order]]
        self.faces = rank[inverse][self.faces]

    def export(self, path):
        ext = os.path.splitext(path)[1].lower()
        if ext == ".ply":
            self._export_ply(path)
        elif ext == ".obj":
            self._export_obj(path)
        else:
            raise ValueError("unsupported mesh format: %s" % ext)
        return path

    def _native(self, fn_name, path):
        """Export through the library's host writers (dj_host_write_ply / _obj): same bytes as the numpy / Python
        writers below, 10x / 15x faster -- the batch meshing worker exports every mesh it extracts."""
        from deepjoin_b200 import _lib
        v = np.ascontiguousarray(self.vertices, dtype=np.float64)
        f = np.ascontiguousarray(self.faces, dtype=np.int64)
        _lib.check(getattr(_lib.lib(), fn_name)(os.fsencode(path), v.ctypes.data, len(v), f.ctypes.data, len(f)))

    def _export_ply(self, path):
        self._native("dj_host_write_ply", path)

    def _export_obj(self, path):
        self._native("dj_host_write_obj", path)

    def _export_ply_numpy(self, path):
        v = self.vertices.astype("<f4")
        f = self.faces.astype("<i4")
        header = ("ply\nformat binary_little_endian 1.0\ncomment deepjoin_b200\n"
                  "element vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                  "element face %d\nproperty list uchar int vertex_indices\nend_header\n" % (len(v), len(f)))
        rec = np.empty(len(f), dtype=[("n", "u1"), ("i", "<i4", (3,))])
        rec["n"] = 3
        rec["i"] = f
        with open(path, "wb") as fh:
            fh.write(header.encode("ascii"))
            fh.write(v.tobytes())
            fh.write(rec.tobytes())

    def _export_obj_python(self, path):
        with open(path, "w") as fh:
            for p in self.vertices:
                fh.write("v %.8f %.8f %.8f\n" % tuple(p))
            for t in self.faces + 1:
                fh.write("f %d %d %d\n" % tuple(t))


def load_ply(path):
    """Reader for the binary PLY written above (round-trip tests)."""
    with open(path, "rb") as fh:
        nv = nf = 0
        while True:
            line = fh.readline().decode("ascii").strip()
            if line.startswith("element vertex"):
                nv = int(line.split()[-1])
            elif line.startswith("element face"):
                nf = int(line.split()[-1])
            elif line == "end_header":
                break
        v = np.frombuffer(fh.read(nv * 12), dtype="<f4").reshape(-1, 3)
        rec = np.frombuffer(fh.read(nf * 13), dtype=[("n", "u1"), ("i", "<i4", (3,))])
    return Mesh(v, rec["i"], process=False)


def load_obj(path):
    """Reader for Wavefront OBJ triangle meshes (v / f records; f entries may carry /vt/vn suffixes)."""
    v, f = [], []
    with open(path, "r") as fh:
        for line in fh:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                v.append([float(x) for x in t[1:4]])
            elif t[0] == "f":
                ids = [int(x.split("/")[0]) for x in t[1:]]
                for k in range(1, len(ids) - 1):  # fan-triangulate polygons
                    f.append([ids[0], ids[k], ids[k + 1]])
    v = np.asarray(v, dtype=np.float64).reshape(-1, 3)
    f = np.asarray(f, dtype=np.int64).reshape(-1, 3)
    f = np.where(f > 0, f - 1, f + len(v))  # 1-based, negative = relative to the end
    return Mesh(v, f, process=False)
