"""Generate the marching-cubes case tables (deepjoin_b200/csrc/mc_tables.h).

Why generated: the reference extracts meshes with scikit-image 0.19.3's Lewiner
marching cubes (core/reconstruct.py:169-174).  scikit-image and its look-up
tables are not present in this offline image, so the tables are derived here
from first principles and validated by invariants (tests/test_mc_tables.py):

* cube corner / edge numbering and the `value - level > 0` index bit follow the
  Lewiner convention used by skimage;
* every face contour is traced on the cube surface; ambiguous faces (two
  diagonal corners above the level) are resolved PER FACE by the asymptotic
  decider on the actual corner values (the same quantity Lewiner's test_face
  evaluates), so each sign configuration has 2^k variants for its k ambiguous
  faces.  Because a face's resolution depends only on that face's four values,
  adjacent cubes always agree -> crack-free output;
* contours are linked into closed loops and fan-triangulated; winding is such
  that face normals point towards corners ABOVE the level.

* polygons are triangulated without chords that lie inside a cube face (the neighbouring cell could
  reuse such a chord -> non-manifold edge); where no such triangulation exists the loop is fanned
  around a cell-centre vertex (pseudo edge id 12), the role Lewiner's 13th vertex plays.

Not reproduced: Lewiner's interior (tunnel) tests and his exact tilings -- see DESIGN.md.
"""
import itertools
import os
import sys

CORNERS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]  # (x,y,z)
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
EDGE_OF = {}
for _e, (_a, _b) in enumerate(EDGES):
    EDGE_OF[(_a, _b)] = _e
    EDGE_OF[(_b, _a)] = _e


def _faces():
    """6 faces, corners in counter-clockwise order seen from OUTSIDE the cube."""
    faces = []
    for axis in range(3):
        for side in (0, 1):
            cs = [i for i, c in enumerate(CORNERS) if c[axis] == side]
            u, v = [a for a in range(3) if a != axis]
            cen = (0.5, 0.5)
            import math
            cs.sort(key=lambda i: math.atan2(CORNERS[i][v] - cen[1], CORNERS[i][u] - cen[0]))
            # orientation: cross(c1-c0, c2-c1) must point along the outward normal
            p = [CORNERS[i] for i in cs]
            e1 = [p[1][k] - p[0][k] for k in range(3)]
            e2 = [p[2][k] - p[1][k] for k in range(3)]
            cr = [e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]]
            outward = 1 if side == 1 else -1
            if cr[axis] * outward < 0:
                cs.reverse()
            # canonical rotation: start at the smallest corner id
            k = cs.index(min(cs))
            faces.append(cs[k:] + cs[:k])
    return faces


FACES = _faces()


def face_segments(cfg, f, joined):
    """Directed contour segments (edge_from, edge_to) on face f; positive region on the left."""
    cs = FACES[f]
    pos = [(cfg >> c) & 1 for c in cs]
    down = [i for i in range(4) if pos[i] and not pos[(i + 1) % 4]]  # + -> - walking CCW
    up = [i for i in range(4) if not pos[i] and pos[(i + 1) % 4]]    # - -> +
    fe = [EDGE_OF[(cs[i], cs[(i + 1) % 4])] for i in range(4)]
    if len(down) == 0:
        return []
    if len(down) == 1:
        return [(fe[down[0]], fe[up[0]])]
    segs = []
    for i in range(4):
        if joined:
            # cut off each NEGATIVE corner c_i: from edge (i-1 -> i) [+ -> -] to edge (i -> i+1) [- -> +]
            if not pos[i]:
                segs.append((fe[(i - 1) % 4], fe[i]))
        else:
            # circle each POSITIVE corner c_i: from edge (i -> i+1) [+ -> -] to edge (i-1 -> i) [- -> +]
            if pos[i]:
                segs.append((fe[i], fe[(i - 1) % 4]))
    return segs


def ambiguous_faces(cfg):
    out = []
    for f, cs in enumerate(FACES):
        pos = [(cfg >> c) & 1 for c in cs]
        if pos[0] == pos[2] and pos[1] == pos[3] and pos[0] != pos[1]:
            out.append(f)
    return out


EDGE_FACES = {e: set() for e in range(12)}
for _f, _cs in enumerate(FACES):
    for _i in range(4):
        EDGE_FACES[EDGE_OF[(_cs[_i], _cs[(_i + 1) % 4])]].add(_f)

CENTRE = 12  # pseudo edge id of the cell-centre vertex (mean of the cell's edge vertices)


def _coplanar(a, b):
    """A chord between the vertices of edges a and b would lie inside a cube face: forbidden,
    the neighbouring cell may use the same chord and the surface would not be a 2-manifold."""
    return bool(EDGE_FACES[a] & EDGE_FACES[b])


def _triangulate_polygon(poly):
    """First triangulation (deterministic order) of the closed polygon without coplanar chords."""
    n = len(poly)
    if n == 3:
        return [tuple(poly)]
    for k in range(1, n - 1):
        if k > 1 and _coplanar(poly[0], poly[k]):
            continue
        if k < n - 2 and _coplanar(poly[k], poly[-1]):
            continue
        left = _triangulate_polygon(poly[: k + 1]) if k > 1 else []
        if left is None:
            continue
        right = _triangulate_polygon(poly[k:]) if k < n - 2 else []
        if right is None:
            continue
        return left + [(poly[0], poly[k], poly[-1])] + right
    return None


def triangulate(cfg, variant):
    amb = ambiguous_faces(cfg)
    nxt = {}
    for f in range(6):
        joined = False
        if f in amb:
            joined = bool((variant >> amb.index(f)) & 1)
        for a, b in face_segments(cfg, f, joined):
            assert a not in nxt
            nxt[a] = b
    tris, seen = [], set()
    for start in sorted(nxt):
        if start in seen:
            continue
        loop, e = [], start
        while e not in seen:
            seen.add(e)
            loop.append(e)
            e = nxt[e]
        assert e == start and len(loop) >= 3
        got = None
        for r in range(len(loop)):
            got = _triangulate_polygon(loop[r:] + loop[:r])
            if got is not None:
                break
        if got is None:  # fan around the cell-centre vertex (the role of Lewiner's 13th vertex)
            got = [(CENTRE, loop[i], loop[(i + 1) % len(loop)]) for i in range(len(loop))]
        tris.extend(got)
    active = [e for e, (a, b) in enumerate(EDGES) if ((cfg >> a) & 1) != ((cfg >> b) & 1)]
    assert sorted(seen) == active
    return tris


def build():
    var_base, n_amb, amb_faces = [], [], []
    variants = []
    for cfg in range(256):
        amb = ambiguous_faces(cfg)
        var_base.append(len(variants))
        n_amb.append(len(amb))
        amb_faces.append(amb + [0] * (6 - len(amb)))
        for v in range(1 << len(amb)):
            tris = triangulate(cfg, v)
            order = []
            for t in tris:
                for e in t:
                    if e not in order:
                        order.append(e)
            variants.append((tris, order))
    return var_base, n_amb, amb_faces, variants


def emit(path):
    var_base, n_amb, amb_faces, variants = build()
    maxt = max(len(t) for t, _ in variants)
    L = []
    L.append("// GENERATED by tools/gen_mc_tables.py -- do not edit.")
    L.append("// Marching-cubes case tables: Lewiner/skimage corner+edge numbering, per-face")
    L.append("// asymptotic-decider variants, fan-triangulated contour loops.")
    L.append("#ifndef DJ_MC_TABLES_H")
    L.append("#define DJ_MC_TABLES_H")
    L.append("#ifndef DJ_MC_QUAL")
    L.append("#define DJ_MC_QUAL static const")
    L.append("#endif")
    L.append("#define DJ_MC_NVARIANTS %d" % len(variants))
    L.append("#define DJ_MC_MAXTRI %d" % maxt)
    L.append("// corner i -> (dx,dy,dz); x = fastest volume axis (axis 2), z = axis 0")
    L.append("DJ_MC_QUAL unsigned char DJ_MC_CORNER[8][3] = {%s};" % ",".join("{%d,%d,%d}" % c for c in CORNERS))
    L.append("DJ_MC_QUAL unsigned char DJ_MC_EDGE[12][2] = {%s};" % ",".join("{%d,%d}" % e for e in EDGES))
    L.append("// face f -> 4 corners, CCW seen from outside")
    L.append("DJ_MC_QUAL unsigned char DJ_MC_FACE[6][4] = {%s};" % ",".join("{%d,%d,%d,%d}" % tuple(f) for f in FACES))
    L.append("DJ_MC_QUAL unsigned short DJ_MC_VARBASE[256] = {%s};" % ",".join(map(str, var_base)))
    L.append("DJ_MC_QUAL unsigned char DJ_MC_NAMB[256] = {%s};" % ",".join(map(str, n_amb)))
    L.append("DJ_MC_QUAL unsigned char DJ_MC_AMBFACE[256][6] = {%s};" % ",".join(
        "{%s}" % ",".join(map(str, a)) for a in amb_faces))
    L.append("// per variant: number of triangles, triangle edge triples, edges in first-appearance order")
    L.append("DJ_MC_QUAL unsigned char DJ_MC_NTRI[DJ_MC_NVARIANTS] = {%s};" % ",".join(str(len(t)) for t, _ in variants))
    rows = []
    for t, _ in variants:
        flat = [e for tri in t for e in tri]
        flat += [255] * (maxt * 3 - len(flat))
        rows.append("{%s}" % ",".join(map(str, flat)))
    L.append("DJ_MC_QUAL unsigned char DJ_MC_TRI[DJ_MC_NVARIANTS][DJ_MC_MAXTRI*3] = {\n%s};" % ",\n".join(rows))
    L.append("DJ_MC_QUAL unsigned char DJ_MC_NEDGE[DJ_MC_NVARIANTS] = {%s};" % ",".join(str(len(o)) for _, o in variants))
    rows = []
    for _, o in variants:
        rows.append("{%s}" % ",".join(map(str, o + [255] * (13 - len(o)))))
    L.append("DJ_MC_QUAL unsigned char DJ_MC_ORDER[DJ_MC_NVARIANTS][13] = {\n%s};" % ",\n".join(rows))
    L.append("#endif")
    with open(path, "w") as f:
        f.write("\n".join(L) + "\n")
    return len(variants), maxt


if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "deepjoin_b200", "csrc", "mc_tables.h")
    n, m = emit(out)
    print("wrote %s: %d variants, max %d triangles/cell" % (out, n, m))
