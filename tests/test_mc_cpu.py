"""Marching cubes, CPU side: table invariants, the C oracle on analytic fields, and a pure-Python
emulation of the GPU's parallel formulation (ownership masks + exclusive scans) checked against
the oracle's literal sequential first-use sweep."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_mc_tables as T  # noqa: E402
from oracle import mc_oracle  # noqa: E402


def test_generated_header_is_current(tmp_path):
    out = tmp_path / "mc_tables.h"
    T.emit(str(out))
    assert out.read_text() == open(os.path.join(ROOT, "deepjoin_b200", "csrc", "mc_tables.h")).read()


def test_table_invariants():
    var_base, n_amb, amb_faces, variants = T.build()
    assert len(variants) == sum(1 << k for k in n_amb)
    for cfg in range(256):
        active = {e for e, (a, b) in enumerate(T.EDGES) if ((cfg >> a) & 1) != ((cfg >> b) & 1)}
        for v in range(1 << n_amb[cfg]):
            tris, order = variants[var_base[cfg] + v]
            used = {e for t in tris for e in t}
            assert used - {T.CENTRE} == active and set(order) == used and len(order) == len(used)
            # no chord inside a cube face: the neighbour could reuse it (non-manifold edge)
            segs = set()
            for f in range(6):
                joined = f in amb_faces[cfg][: n_amb[cfg]] and bool((v >> amb_faces[cfg][: n_amb[cfg]].index(f)) & 1)
                for a, b in T.face_segments(cfg, f, joined):
                    segs.add((a, b))
            for t in tris:
                for i in range(3):
                    a, b = t[i], t[(i + 1) % 3]
                    if T.CENTRE in (a, b) or (a, b) in segs:
                        continue
                    assert not T._coplanar(a, b), (cfg, v, t)
            # every directed edge of the patch boundary lies on a cube face; interior edges pair up
            directed = [(t[i], t[(i + 1) % 3]) for t in tris for i in range(3)]
            assert len(set(directed)) == len(directed)
    # complementary configurations use the same edges
    for cfg in range(256):
        a = {e for t in variants[var_base[cfg]][0] for e in t} - {T.CENTRE}
        b = {e for t in variants[var_base[255 - cfg]][0] for e in t} - {T.CENTRE}
        assert a == b


def _parse_header_array(src, name):
    import re
    m = re.search(r"%s(?:\[[^\]]*\])+ = \{(.*?)\};" % name, src, flags=re.S)
    return [int(x) for x in re.findall(r"-?\d+", m.group(1))]


def test_product_tables_equal_the_oracles_own_derivation():
    """The oracle derives its case tables itself (oracle/mc_case_tables.py: its own statement of the Lorensen-Cline
    table, a marching-squares table per face, its own loop tracer) and shares no source with the product; a wrong
    entry on either side shows up here (the oracle used to #include the product's header)."""
    from oracle import mc_case_tables as O
    assert "mc_tables.h" not in open(os.path.join(ROOT, "oracle", "mc_oracle.c")).read().split("*/", 1)[1]
    t = O.build()
    src = open(os.path.join(ROOT, "deepjoin_b200", "csrc", "mc_tables.h")).read()
    nvar = len(t["n_tri"])
    assert _parse_header_array(src, "DJ_MC_VARBASE") == t["var_base"]
    assert _parse_header_array(src, "DJ_MC_NAMB") == t["n_amb"]
    assert _parse_header_array(src, "DJ_MC_AMBFACE") == t["amb_face"]
    assert _parse_header_array(src, "DJ_MC_NTRI") == t["n_tri"]
    assert _parse_header_array(src, "DJ_MC_FACE") == t["face_corners"]
    tri = _parse_header_array(src, "DJ_MC_TRI")
    maxt = len(tri) // nvar
    for v in range(nvar):
        assert tri[v * maxt: v * maxt + 3 * t["n_tri"][v]] == t["tri"][v * 36: v * 36 + 3 * t["n_tri"][v]], v
    assert tuple(tuple(r) for r in T.CLASSIC) == O.TRITABLE


def test_classic_rows_are_the_unresolved_variants():
    """Variant 0 of every configuration -- all 136 unambiguous configurations, and the separated resolution of the
    120 ambiguous ones -- is its row of the Lorensen-Cline / Bourke table; every row triangulates exactly the traced
    contour loops (edges, orientation, pairing of interior sides)."""
    from oracle import mc_case_tables as O
    var_base, n_amb, amb_faces, variants = T.build()
    n_unamb = 0
    for cfg in range(256):
        row = O.TRITABLE[cfg]
        tris = [tuple(row[i:i + 3]) for i in range(0, len(row), 3)]
        assert variants[var_base[cfg]][0] == tris
        n_unamb += n_amb[cfg] == 0
        sides = [(t[i], t[(i + 1) % 3]) for t in tris for i in range(3)]
        assert len(set(sides)) == len(sides)
        boundary = sorted(s for s in sides if (s[1], s[0]) not in sides)
        loops = O.contour_loops(cfg, 0)
        assert boundary == sorted((l[i], l[(i + 1) % len(l)]) for l in loops for i in range(len(l)))
        assert len(tris) == sum(len(l) - 2 for l in loops)
    assert n_unamb == 136
    # spot values every marching-cubes text quotes
    assert O.TRITABLE[1] == (0, 8, 3) and O.TRITABLE[3] == (1, 8, 3, 9, 8, 1) and O.TRITABLE[254] == (0, 3, 8)
    assert O.TRITABLE[15] == (9, 8, 10, 10, 8, 11)


def test_face_test_dead_band():
    """Lewiner's test_face: |AC - BD| < FLT_EPSILON -> the corners above the level are joined (his `return face >= 0`)."""
    lib = mc_oracle.lib()
    import ctypes
    eps = float(np.finfo(np.float32).eps)
    # configuration 5 (corners 0 and 2 above the level): its one ambiguous face is the bottom (0,3,2,1)
    for d, joined in ((0.0, 1), (0.9 * eps, 1), (-0.9 * eps, 1), (1e-3, 1), (-1e-3, 0)):
        # A = f0 = 1, C = f2 = 1 + d, B = f3 = -1, D = f1 = -1:  AC - BD = d
        f = (ctypes.c_double * 8)(1.0, -1.0, 1.0 + d, -1.0, -1.0, -1.0, -1.0, -1.0)
        var_base, n_amb, amb_faces, variants = T.build()
        assert n_amb[5] == 1
        assert lib.dj_oracle_mc_variant(f) - var_base[5] == joined, d


def test_skimage_comparison_when_installed():
    """Runs wherever scikit-image is importable (it is not in the offline image: marching-cubes parity against the
    library stays UNPINNED there).  Vertex sets must agree; face arrays are reported, not asserted, for the joined
    ambiguous variants (Lewiner's own tilings are not restated)."""
    pytest.importorskip("skimage")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import compare_with_skimage as C
    rows = C.compare(32, verbose=True)
    for r in rows:
        assert r["skimage"][0] == r["ours"][0], r  # same vertices: one per intersected grid edge (+ centre vertices)
        if r["volume"] in ("sphere", "torus"):
            assert r["skimage"][1] == r["ours"][1], r


def _closed_manifold(v, f):
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]).astype(np.int64)
    key = e[:, 0] * (len(v) + 1) + e[:, 1]
    rkey = e[:, 1] * (len(v) + 1) + e[:, 0]
    return len(np.unique(key)) == len(key) and np.isin(rkey, key).all()


def _fields(n):
    ax = np.linspace(-1, 1, n)
    z, y, x = np.meshgrid(ax, ax, ax, indexing="ij")
    sphere = np.sqrt(x * x + y * y + z * z) - 0.7
    torus = np.sqrt((np.sqrt(x * x + y * y) - 0.55) ** 2 + z * z) - 0.25
    gyroid = np.sin(4 * x) * np.cos(4 * y) + np.sin(4 * y) * np.cos(4 * z) + np.sin(4 * z) * np.cos(4 * x)
    return {"sphere": sphere, "torus": torus, "gyroid": gyroid}


@pytest.mark.parametrize("name", ["sphere", "torus"])
def test_oracle_closed_surfaces(name):
    vol = _fields(40)[name].astype(np.float32)
    v, f = mc_oracle.marching_cubes(vol, 0.0, "ascent")
    assert _closed_manifold(v, f)
    chi = len(v) - (3 * len(f)) // 2 + len(f)
    assert chi == (2 if name == "sphere" else 0)
    # ascent on an SDF: normals point outwards (towards larger values) in the (x,y,z)=(col2,col1,col0) frame
    p = v[:, ::-1].astype(np.float64)
    vol6 = np.einsum("ij,ij->i", p[f[:, 0]], np.cross(p[f[:, 1]], p[f[:, 2]])).sum()
    assert vol6 > 0
    vd, fd = mc_oracle.marching_cubes(vol, 0.0, "descent")
    assert np.array_equal(vd, v) and np.array_equal(fd, f[:, ::-1])


def test_oracle_crack_free_on_ambiguous_field():
    rng = np.random.default_rng(0)
    vol = (_fields(24)["gyroid"] + 0.3 * rng.uniform(-1, 1, (24, 24, 24))).astype(np.float32)
    v, f = mc_oracle.marching_cubes(vol, 0.0, "ascent")
    # open at the volume boundary only: every interior directed edge has its reverse
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]).astype(np.int64)
    key = e[:, 0] * (len(v) + 1) + e[:, 1]
    rkey = e[:, 1] * (len(v) + 1) + e[:, 0]
    assert len(np.unique(key)) == len(key)
    lonely = e[~np.isin(rkey, key)]
    on_border = lambda p: ((p <= 1e-6) | (p >= 23 - 1e-6)).any(axis=1)
    assert on_border(v[lonely[:, 0]]).all() and on_border(v[lonely[:, 1]]).all()


def test_oracle_errors():
    vol = _fields(12)["sphere"].astype(np.float32)
    with pytest.raises(ValueError):
        mc_oracle.marching_cubes(vol, 10.0)
    with pytest.raises(ValueError):
        mc_oracle.marching_cubes(vol, float(vol.min()) - 1.0)
    with pytest.raises(ValueError):
        mc_oracle.marching_cubes(vol, 0.0, "sideways")
    flat = np.zeros((4, 4, 4), np.float32)
    with pytest.raises(RuntimeError):  # level == min == max: inside the range but no surface
        mc_oracle.marching_cubes(flat, 0.0)


def test_vertex_interpolation_formula():
    vol = np.zeros((2, 2, 2), np.float32)
    vol[1, :, :] = 3.0
    vol[0, :, :] = -1.0
    v, f = mc_oracle.marching_cubes(vol, 0.0, "ascent")
    eps = float(np.finfo(np.float32).eps)
    w1, w2 = 1 / (eps + 1.0), 1 / (eps + 3.0)
    assert np.allclose(v[:, 0], np.float32(w2 / (w1 + w2)), rtol=0, atol=0)
    assert len(v) == 4 and len(f) == 2


# ---- emulation of the GPU algorithm (deepjoin_b200/csrc/mc.cuh) --------------------------------
MX, MY, MZ = 0x988, 0x311, 0x00F


def _parallel_mc(vol, level, seam=False):
    var_base, n_amb, amb_faces, variants = T.build()
    n0, n1, n2 = vol.shape
    cells = []
    for z in range(n0 - 1):
        for y in range(n1 - 1):
            for x in range(n2 - 1):
                f = [float(vol[z + c[2], y + c[1], x + c[0]]) - level for c in T.CORNERS]
                cfg = sum(1 << i for i in range(8) if f[i] > 0)
                var = var_base[cfg]
                for a in range(n_amb[cfg]):
                    c = T.FACES[amb_faces[cfg][a]]
                    det = f[c[0]] * f[c[2]] - f[c[1]] * f[c[3]]
                    joined = abs(det) < float(np.finfo(np.float32).eps) or ((det > 0) if (cfg >> c[0]) & 1 else (det < 0))
                    var += (1 << a) if joined else 0
                own = 0x1FFF & ~((MZ if (z or seam) else 0) | (MY if y else 0) | (MX if x else 0))
                cells.append((z, y, x, var, own, f))
    edge_id, verts, faces = {}, [], []

    def slot(z, y, x, e):
        if e == T.CENTRE:
            return (z, y, x, 3)
        a, b = T.EDGES[e]
        ca, cb = T.CORNERS[a], T.CORNERS[b]
        d = 0 if ca[0] != cb[0] else (1 if ca[1] != cb[1] else 2)
        return (z + min(ca[2], cb[2]), y + min(ca[1], cb[1]), x + min(ca[0], cb[0]), d)

    # pass "emit vertices": ids by exclusive scan of owned counts in sweep order
    nid = 0
    for z, y, x, var, own, f in cells:
        tris, order = variants[var]
        for e in order:
            if (own >> e) & 1:
                edge_id[slot(z, y, x, e)] = nid
                nid += 1
    for z, y, x, var, own, f in cells:
        for t in variants[var][0]:
            faces.append([edge_id.get(slot(z, y, x, e), -1) for e in t])
    _parallel_mc.last_edge_id = edge_id
    return nid, np.asarray(faces, dtype=np.int64).reshape(-1, 3)


@pytest.mark.parametrize("shape,seed", [((7, 6, 5), 0), ((5, 9, 4), 1), ((6, 6, 6), 2)])
def test_parallel_formulation_equals_sequential_sweep(shape, seed):
    rng = np.random.default_rng(seed)
    vol = rng.uniform(-1, 1, shape).astype(np.float32)  # noise: every case, many ambiguous faces
    v, f = mc_oracle.marching_cubes(vol, 0.0, "ascent")
    nv, pf = _parallel_mc(vol, 0.0)
    assert nv == len(v)
    assert np.array_equal(pf, f.astype(np.int64))


def test_seam_mode_leaves_only_plane0_references_open():
    rng = np.random.default_rng(5)
    vol = rng.uniform(-1, 1, (5, 6, 7)).astype(np.float32)
    nv, pf = _parallel_mc(vol, 0.0, seam=True)
    nv_full, pf_full = _parallel_mc(vol, 0.0, seam=False)
    seam_ids = {i for (z, y, x, d), i in _parallel_mc.last_edge_id.items() if z == 0 and d in (0, 1)}
    assert nv == nv_full - len(seam_ids) and pf.shape == pf_full.shape
    assert ((pf == -1) == np.isin(pf_full, list(seam_ids))).all()


def test_kernel_shortcut_constants_match_the_tables():
    """classify_kernel skips the table look-ups for cells the surface does not cut: its two hard-wired variant ids
    must be the tables' (and triangle-free)."""
    import re
    var_base, n_amb, amb_faces, variants = T.build()
    src = open(os.path.join(ROOT, "deepjoin_b200", "csrc", "mc.cuh")).read()
    m = re.search(r"DJ_MC_VAR_EMPTY0 = (\d+), DJ_MC_VAR_EMPTY255 = (\d+)", src)
    assert (int(m.group(1)), int(m.group(2))) == (var_base[0], var_base[255])
    assert len(variants[var_base[0]][0]) == 0 and len(variants[var_base[255]][0]) == 0 and n_amb[0] == 0 and n_amb[255] == 0
