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

* variant 0 of every configuration (no ambiguous face resolved as "joined": all unambiguous configurations, and
  the classic resolution of the ambiguous ones) is the row of the published Lorensen-Cline / Bourke table
  (`CLASSIC` below, restated here and checked row by row against the traced face contours): these are the
  triangle lists scikit-image's CASESCLASSIC and the Lewiner tilings of the unambiguous cases encode;
* ambiguous faces use Lewiner's `test_face`: positive corners joined iff |AC - BD| < FLT_EPSILON or the
  face-centre saddle value is positive.

Not reproduced: Lewiner's interior (tunnel) tests and his tilings of the joined variants -- see DESIGN.md.
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

# The Lorensen-Cline / Bourke triangle table, index bit i = corner i above the level (row = 3 edge ids per triangle).
CLASSIC = [
[],
[0, 8, 3],
[0, 1, 9],
[1, 8, 3, 9, 8, 1],
[1, 2, 10],
[0, 8, 3, 1, 2, 10],
[9, 2, 10, 0, 2, 9],
[2, 8, 3, 2, 10, 8, 10, 9, 8],
[3, 11, 2],
[0, 11, 2, 8, 11, 0],
[1, 9, 0, 2, 3, 11],
[1, 11, 2, 1, 9, 11, 9, 8, 11],
[3, 10, 1, 11, 10, 3],
[0, 10, 1, 0, 8, 10, 8, 11, 10],
[3, 9, 0, 3, 11, 9, 11, 10, 9],
[9, 8, 10, 10, 8, 11],
[4, 7, 8],
[4, 3, 0, 7, 3, 4],
[0, 1, 9, 8, 4, 7],
[4, 1, 9, 4, 7, 1, 7, 3, 1],
[1, 2, 10, 8, 4, 7],
[3, 4, 7, 3, 0, 4, 1, 2, 10],
[9, 2, 10, 9, 0, 2, 8, 4, 7],
[2, 10, 9, 2, 9, 7, 2, 7, 3, 7, 9, 4],
[8, 4, 7, 3, 11, 2],
[11, 4, 7, 11, 2, 4, 2, 0, 4],
[9, 0, 1, 8, 4, 7, 2, 3, 11],
[4, 7, 11, 9, 4, 11, 9, 11, 2, 9, 2, 1],
[3, 10, 1, 3, 11, 10, 7, 8, 4],
[1, 11, 10, 1, 4, 11, 1, 0, 4, 7, 11, 4],
[4, 7, 8, 9, 0, 11, 9, 11, 10, 11, 0, 3],
[4, 7, 11, 4, 11, 9, 9, 11, 10],
[9, 5, 4],
[9, 5, 4, 0, 8, 3],
[0, 5, 4, 1, 5, 0],
[8, 5, 4, 8, 3, 5, 3, 1, 5],
[1, 2, 10, 9, 5, 4],
[3, 0, 8, 1, 2, 10, 4, 9, 5],
[5, 2, 10, 5, 4, 2, 4, 0, 2],
[2, 10, 5, 3, 2, 5, 3, 5, 4, 3, 4, 8],
[9, 5, 4, 2, 3, 11],
[0, 11, 2, 0, 8, 11, 4, 9, 5],
[0, 5, 4, 0, 1, 5, 2, 3, 11],
[2, 1, 5, 2, 5, 8, 2, 8, 11, 4, 8, 5],
[10, 3, 11, 10, 1, 3, 9, 5, 4],
[4, 9, 5, 0, 8, 1, 8, 10, 1, 8, 11, 10],
[5, 4, 0, 5, 0, 11, 5, 11, 10, 11, 0, 3],
[5, 4, 8, 5, 8, 10, 10, 8, 11],
[9, 7, 8, 5, 7, 9],
[9, 3, 0, 9, 5, 3, 5, 7, 3],
[0, 7, 8, 0, 1, 7, 1, 5, 7],
[1, 5, 3, 3, 5, 7],
[9, 7, 8, 9, 5, 7, 10, 1, 2],
[10, 1, 2, 9, 5, 0, 5, 3, 0, 5, 7, 3],
[8, 0, 2, 8, 2, 5, 8, 5, 7, 10, 5, 2],
[2, 10, 5, 2, 5, 3, 3, 5, 7],
[7, 9, 5, 7, 8, 9, 3, 11, 2],
[9, 5, 7, 9, 7, 2, 9, 2, 0, 2, 7, 11],
[2, 3, 11, 0, 1, 8, 1, 7, 8, 1, 5, 7],
[11, 2, 1, 11, 1, 7, 7, 1, 5],
[9, 5, 8, 8, 5, 7, 10, 1, 3, 10, 3, 11],
[5, 7, 0, 5, 0, 9, 7, 11, 0, 1, 0, 10, 11, 10, 0],
[11, 10, 0, 11, 0, 3, 10, 5, 0, 8, 0, 7, 5, 7, 0],
[11, 10, 5, 7, 11, 5],
[10, 6, 5],
[0, 8, 3, 5, 10, 6],
[9, 0, 1, 5, 10, 6],
[1, 8, 3, 1, 9, 8, 5, 10, 6],
[1, 6, 5, 2, 6, 1],
[1, 6, 5, 1, 2, 6, 3, 0, 8],
[9, 6, 5, 9, 0, 6, 0, 2, 6],
[5, 9, 8, 5, 8, 2, 5, 2, 6, 3, 2, 8],
[2, 3, 11, 10, 6, 5],
[11, 0, 8, 11, 2, 0, 10, 6, 5],
[0, 1, 9, 2, 3, 11, 5, 10, 6],
[5, 10, 6, 1, 9, 2, 9, 11, 2, 9, 8, 11],
[6, 3, 11, 6, 5, 3, 5, 1, 3],
[0, 8, 11, 0, 11, 5, 0, 5, 1, 5, 11, 6],
[3, 11, 6, 0, 3, 6, 0, 6, 5, 0, 5, 9],
[6, 5, 9, 6, 9, 11, 11, 9, 8],
[5, 10, 6, 4, 7, 8],
[4, 3, 0, 4, 7, 3, 6, 5, 10],
[1, 9, 0, 5, 10, 6, 8, 4, 7],
[10, 6, 5, 1, 9, 7, 1, 7, 3, 7, 9, 4],
[6, 1, 2, 6, 5, 1, 4, 7, 8],
[1, 2, 5, 5, 2, 6, 3, 0, 4, 3, 4, 7],
[8, 4, 7, 9, 0, 5, 0, 6, 5, 0, 2, 6],
[7, 3, 9, 7, 9, 4, 3, 2, 9, 5, 9, 6, 2, 6, 9],
[3, 11, 2, 7, 8, 4, 10, 6, 5],
[5, 10, 6, 4, 7, 2, 4, 2, 0, 2, 7, 11],
[0, 1, 9, 4, 7, 8, 2, 3, 11, 5, 10, 6],
[9, 2, 1, 9, 11, 2, 9, 4, 11, 7, 11, 4, 5, 10, 6],
[8, 4, 7, 3, 11, 5, 3, 5, 1, 5, 11, 6],
[5, 1, 11, 5, 11, 6, 1, 0, 11, 7, 11, 4, 0, 4, 11],
[0, 5, 9, 0, 6, 5, 0, 3, 6, 11, 6, 3, 8, 4, 7],
[6, 5, 9, 6, 9, 11, 4, 7, 9, 7, 11, 9],
[10, 4, 9, 6, 4, 10],
[4, 10, 6, 4, 9, 10, 0, 8, 3],
[10, 0, 1, 10, 6, 0, 6, 4, 0],
[8, 3, 1, 8, 1, 6, 8, 6, 4, 6, 1, 10],
[1, 4, 9, 1, 2, 4, 2, 6, 4],
[3, 0, 8, 1, 2, 9, 2, 4, 9, 2, 6, 4],
[0, 2, 4, 4, 2, 6],
[8, 3, 2, 8, 2, 4, 4, 2, 6],
[10, 4, 9, 10, 6, 4, 11, 2, 3],
[0, 8, 2, 2, 8, 11, 4, 9, 10, 4, 10, 6],
[3, 11, 2, 0, 1, 6, 0, 6, 4, 6, 1, 10],
[6, 4, 1, 6, 1, 10, 4, 8, 1, 2, 1, 11, 8, 11, 1],
[9, 6, 4, 9, 3, 6, 9, 1, 3, 11, 6, 3],
[8, 11, 1, 8, 1, 0, 11, 6, 1, 9, 1, 4, 6, 4, 1],
[3, 11, 6, 3, 6, 0, 0, 6, 4],
[6, 4, 8, 11, 6, 8],
[7, 10, 6, 7, 8, 10, 8, 9, 10],
[0, 7, 3, 0, 10, 7, 0, 9, 10, 6, 7, 10],
[10, 6, 7, 1, 10, 7, 1, 7, 8, 1, 8, 0],
[10, 6, 7, 10, 7, 1, 1, 7, 3],
[1, 2, 6, 1, 6, 8, 1, 8, 9, 8, 6, 7],
[2, 6, 9, 2, 9, 1, 6, 7, 9, 0, 9, 3, 7, 3, 9],
[7, 8, 0, 7, 0, 6, 6, 0, 2],
[7, 3, 2, 6, 7, 2],
[2, 3, 11, 10, 6, 8, 10, 8, 9, 8, 6, 7],
[2, 0, 7, 2, 7, 11, 0, 9, 7, 6, 7, 10, 9, 10, 7],
[1, 8, 0, 1, 7, 8, 1, 10, 7, 6, 7, 10, 2, 3, 11],
[11, 2, 1, 11, 1, 7, 10, 6, 1, 6, 7, 1],
[8, 9, 6, 8, 6, 7, 9, 1, 6, 11, 6, 3, 1, 3, 6],
[0, 9, 1, 11, 6, 7],
[7, 8, 0, 7, 0, 6, 3, 11, 0, 11, 6, 0],
[7, 11, 6],
[7, 6, 11],
[3, 0, 8, 11, 7, 6],
[0, 1, 9, 11, 7, 6],
[8, 1, 9, 8, 3, 1, 11, 7, 6],
[10, 1, 2, 6, 11, 7],
[1, 2, 10, 3, 0, 8, 6, 11, 7],
[2, 9, 0, 2, 10, 9, 6, 11, 7],
[6, 11, 7, 2, 10, 3, 10, 8, 3, 10, 9, 8],
[7, 2, 3, 6, 2, 7],
[7, 0, 8, 7, 6, 0, 6, 2, 0],
[2, 7, 6, 2, 3, 7, 0, 1, 9],
[1, 6, 2, 1, 8, 6, 1, 9, 8, 8, 7, 6],
[10, 7, 6, 10, 1, 7, 1, 3, 7],
[10, 7, 6, 1, 7, 10, 1, 8, 7, 1, 0, 8],
[0, 3, 7, 0, 7, 10, 0, 10, 9, 6, 10, 7],
[7, 6, 10, 7, 10, 8, 8, 10, 9],
[6, 8, 4, 11, 8, 6],
[3, 6, 11, 3, 0, 6, 0, 4, 6],
[8, 6, 11, 8, 4, 6, 9, 0, 1],
[9, 4, 6, 9, 6, 3, 9, 3, 1, 11, 3, 6],
[6, 8, 4, 6, 11, 8, 2, 10, 1],
[1, 2, 10, 3, 0, 11, 0, 6, 11, 0, 4, 6],
[4, 11, 8, 4, 6, 11, 0, 2, 9, 2, 10, 9],
[10, 9, 3, 10, 3, 2, 9, 4, 3, 11, 3, 6, 4, 6, 3],
[8, 2, 3, 8, 4, 2, 4, 6, 2],
[0, 4, 2, 4, 6, 2],
[1, 9, 0, 2, 3, 4, 2, 4, 6, 4, 3, 8],
[1, 9, 4, 1, 4, 2, 2, 4, 6],
[8, 1, 3, 8, 6, 1, 8, 4, 6, 6, 10, 1],
[10, 1, 0, 10, 0, 6, 6, 0, 4],
[4, 6, 3, 4, 3, 8, 6, 10, 3, 0, 3, 9, 10, 9, 3],
[10, 9, 4, 6, 10, 4],
[4, 9, 5, 7, 6, 11],
[0, 8, 3, 4, 9, 5, 11, 7, 6],
[5, 0, 1, 5, 4, 0, 7, 6, 11],
[11, 7, 6, 8, 3, 4, 3, 5, 4, 3, 1, 5],
[9, 5, 4, 10, 1, 2, 7, 6, 11],
[6, 11, 7, 1, 2, 10, 0, 8, 3, 4, 9, 5],
[7, 6, 11, 5, 4, 10, 4, 2, 10, 4, 0, 2],
[3, 4, 8, 3, 5, 4, 3, 2, 5, 10, 5, 2, 11, 7, 6],
[7, 2, 3, 7, 6, 2, 5, 4, 9],
[9, 5, 4, 0, 8, 6, 0, 6, 2, 6, 8, 7],
[3, 6, 2, 3, 7, 6, 1, 5, 0, 5, 4, 0],
[6, 2, 8, 6, 8, 7, 2, 1, 8, 4, 8, 5, 1, 5, 8],
[9, 5, 4, 10, 1, 6, 1, 7, 6, 1, 3, 7],
[1, 6, 10, 1, 7, 6, 1, 0, 7, 8, 7, 0, 9, 5, 4],
[4, 0, 10, 4, 10, 5, 0, 3, 10, 6, 10, 7, 3, 7, 10],
[7, 6, 10, 7, 10, 8, 5, 4, 10, 4, 8, 10],
[6, 9, 5, 6, 11, 9, 11, 8, 9],
[3, 6, 11, 0, 6, 3, 0, 5, 6, 0, 9, 5],
[0, 11, 8, 0, 5, 11, 0, 1, 5, 5, 6, 11],
[6, 11, 3, 6, 3, 5, 5, 3, 1],
[1, 2, 10, 9, 5, 11, 9, 11, 8, 11, 5, 6],
[0, 11, 3, 0, 6, 11, 0, 9, 6, 5, 6, 9, 1, 2, 10],
[11, 8, 5, 11, 5, 6, 8, 0, 5, 10, 5, 2, 0, 2, 5],
[6, 11, 3, 6, 3, 5, 2, 10, 3, 10, 5, 3],
[5, 8, 9, 5, 2, 8, 5, 6, 2, 3, 8, 2],
[9, 5, 6, 9, 6, 0, 0, 6, 2],
[1, 5, 8, 1, 8, 0, 5, 6, 8, 3, 8, 2, 6, 2, 8],
[1, 5, 6, 2, 1, 6],
[1, 3, 6, 1, 6, 10, 3, 8, 6, 5, 6, 9, 8, 9, 6],
[10, 1, 0, 10, 0, 6, 9, 5, 0, 5, 6, 0],
[0, 3, 8, 5, 6, 10],
[10, 5, 6],
[11, 5, 10, 7, 5, 11],
[11, 5, 10, 11, 7, 5, 8, 3, 0],
[5, 11, 7, 5, 10, 11, 1, 9, 0],
[10, 7, 5, 10, 11, 7, 9, 8, 1, 8, 3, 1],
[11, 1, 2, 11, 7, 1, 7, 5, 1],
[0, 8, 3, 1, 2, 7, 1, 7, 5, 7, 2, 11],
[9, 7, 5, 9, 2, 7, 9, 0, 2, 2, 11, 7],
[7, 5, 2, 7, 2, 11, 5, 9, 2, 3, 2, 8, 9, 8, 2],
[2, 5, 10, 2, 3, 5, 3, 7, 5],
[8, 2, 0, 8, 5, 2, 8, 7, 5, 10, 2, 5],
[9, 0, 1, 5, 10, 3, 5, 3, 7, 3, 10, 2],
[9, 8, 2, 9, 2, 1, 8, 7, 2, 10, 2, 5, 7, 5, 2],
[1, 3, 5, 3, 7, 5],
[0, 8, 7, 0, 7, 1, 1, 7, 5],
[9, 0, 3, 9, 3, 5, 5, 3, 7],
[9, 8, 7, 5, 9, 7],
[5, 8, 4, 5, 10, 8, 10, 11, 8],
[5, 0, 4, 5, 11, 0, 5, 10, 11, 11, 3, 0],
[0, 1, 9, 8, 4, 10, 8, 10, 11, 10, 4, 5],
[10, 11, 4, 10, 4, 5, 11, 3, 4, 9, 4, 1, 3, 1, 4],
[2, 5, 1, 2, 8, 5, 2, 11, 8, 4, 5, 8],
[0, 4, 11, 0, 11, 3, 4, 5, 11, 2, 11, 1, 5, 1, 11],
[0, 2, 5, 0, 5, 9, 2, 11, 5, 4, 5, 8, 11, 8, 5],
[9, 4, 5, 2, 11, 3],
[2, 5, 10, 3, 5, 2, 3, 4, 5, 3, 8, 4],
[5, 10, 2, 5, 2, 4, 4, 2, 0],
[3, 10, 2, 3, 5, 10, 3, 8, 5, 4, 5, 8, 0, 1, 9],
[5, 10, 2, 5, 2, 4, 1, 9, 2, 9, 4, 2],
[8, 4, 5, 8, 5, 3, 3, 5, 1],
[0, 4, 5, 1, 0, 5],
[8, 4, 5, 8, 5, 3, 9, 0, 5, 0, 3, 5],
[9, 4, 5],
[4, 11, 7, 4, 9, 11, 9, 10, 11],
[0, 8, 3, 4, 9, 7, 9, 11, 7, 9, 10, 11],
[1, 10, 11, 1, 11, 4, 1, 4, 0, 7, 4, 11],
[3, 1, 4, 3, 4, 8, 1, 10, 4, 7, 4, 11, 10, 11, 4],
[4, 11, 7, 9, 11, 4, 9, 2, 11, 9, 1, 2],
[9, 7, 4, 9, 11, 7, 9, 1, 11, 2, 11, 1, 0, 8, 3],
[11, 7, 4, 11, 4, 2, 2, 4, 0],
[11, 7, 4, 11, 4, 2, 8, 3, 4, 3, 2, 4],
[2, 9, 10, 2, 7, 9, 2, 3, 7, 7, 4, 9],
[9, 10, 7, 9, 7, 4, 10, 2, 7, 8, 7, 0, 2, 0, 7],
[3, 7, 10, 3, 10, 2, 7, 4, 10, 1, 10, 0, 4, 0, 10],
[1, 10, 2, 8, 7, 4],
[4, 9, 1, 4, 1, 7, 7, 1, 3],
[4, 9, 1, 4, 1, 7, 0, 8, 1, 8, 7, 1],
[4, 0, 3, 7, 4, 3],
[4, 8, 7],
[9, 10, 8, 10, 11, 8],
[3, 0, 9, 3, 9, 11, 11, 9, 10],
[0, 1, 10, 0, 10, 8, 8, 10, 11],
[3, 1, 10, 11, 3, 10],
[1, 2, 11, 1, 11, 9, 9, 11, 8],
[3, 0, 9, 3, 9, 11, 1, 2, 9, 2, 11, 9],
[0, 2, 11, 8, 0, 11],
[3, 2, 11],
[2, 3, 8, 2, 8, 10, 10, 8, 9],
[9, 10, 2, 0, 9, 2],
[2, 3, 8, 2, 8, 10, 0, 1, 8, 1, 10, 8],
[1, 10, 2],
[1, 3, 8, 9, 1, 8],
[0, 9, 1],
[0, 3, 8],
[],
]


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


def classic_triangles(cfg):
    row = CLASSIC[cfg]
    return [tuple(row[i:i + 3]) for i in range(0, len(row), 3)]


def patch_boundary(tris):
    """Directed sides of a triangle list that are not cancelled by their reverse (None: a side used twice)."""
    sides = {}
    for t in tris:
        for i in range(3):
            k = (t[i], t[(i + 1) % 3])
            if k in sides:
                return None
            sides[k] = 1
    return sorted(k for k in sides if (k[1], k[0]) not in sides)


def check_classic():
    """Every CLASSIC row triangulates exactly the contour loops of variant 0 (same edges, same orientation, interior
    sides paired, no interior side inside a cube face)."""
    for cfg in range(256):
        tris = classic_triangles(cfg)
        segs = sorted(sg for f in range(6) for sg in face_segments(cfg, f, False))
        assert all(len(set(t)) == 3 for t in tris), cfg
        assert patch_boundary(tris) == segs, cfg
        n_loops = len(_loops(cfg, 0))
        assert len(tris) == len(segs) - 2 * n_loops, cfg
        bd = set(segs)
        for t in tris:
            for i in range(3):
                a, b = t[i], t[(i + 1) % 3]
                assert (a, b) in bd or not _coplanar(a, b), cfg


def _loops(cfg, variant):
    amb = ambiguous_faces(cfg)
    nxt = {}
    for f in range(6):
        joined = f in amb and bool((variant >> amb.index(f)) & 1)
        for a, b in face_segments(cfg, f, joined):
            nxt[a] = b
    loops, seen = [], set()
    for start in sorted(nxt):
        if start in seen:
            continue
        loop, e = [], start
        while e not in seen:
            seen.add(e)
            loop.append(e)
            e = nxt[e]
        loops.append(loop)
    return loops


def build():
    check_classic()
    var_base, n_amb, amb_faces = [], [], []
    variants = []
    for cfg in range(256):
        amb = ambiguous_faces(cfg)
        var_base.append(len(variants))
        n_amb.append(len(amb))
        amb_faces.append(amb + [0] * (6 - len(amb)))
        for v in range(1 << len(amb)):
            tris = classic_triangles(cfg) if v == 0 else triangulate(cfg, v)
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
    L.append("// Marching-cubes case tables: Lewiner/skimage corner+edge numbering; variant 0 of a configuration = its")
    L.append("// row of the classic Lorensen-Cline / Bourke table, further variants = per-face test_face outcomes.")
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
    L.append("// per variant: bit e set = edge e (12 = the cell-centre vertex) carries a vertex of this variant")
    L.append("DJ_MC_QUAL unsigned short DJ_MC_EMASK[DJ_MC_NVARIANTS] = {%s};" % ",".join(str(sum(1 << e for e in o)) for _, o in variants))
    L.append("#endif")
    with open(path, "w") as f:
        f.write("\n".join(L) + "\n")
    return len(variants), maxt


if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "deepjoin_b200", "csrc", "mc_tables.h")
    n, m = emit(out)
    print("wrote %s: %d variants, max %d triangles/cell" % (out, n, m))
