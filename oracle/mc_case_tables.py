"""Marching-cubes case tables of the ORACLE, derived here independently of the product's generator
(tools/gen_mc_tables.py -> deepjoin_b200/csrc/mc_tables.h).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED against scikit-image 0.19.3 (`_marching_cubes_lewiner_luts.py` is not
available offline); what is restated:

* corner / edge numbering of Lewiner's implementation (the one skimage uses): corners 0-3 on the z = 0 face
  counter-clockwise from the origin, 4-7 above them; edges 0-3 on the bottom, 4-7 on top, 8-11 vertical;
  pseudo-edge 12 = cell-centre vertex;
* cube index bit i = (v_i - level > 0);
* UNAMBIGUOUS configurations, and the "separated" resolution of ambiguous ones: the published Lorensen-Cline /
  Bourke triangle table (`TRITABLE`, written out below);
* ambiguous faces: Lewiner's test_face outcome per face (the oracle's C sweep evaluates it); a configuration with k
  ambiguous faces has 2^k variants, variant bit j = "the corners above the level are joined across the j-th ambiguous
  face (faces in the order of `FACE_CORNERS`)";
* joined variants: the contour loops are traced with a marching-squares table per face and triangulated by the rule
  documented in `triangulate_loop` (first triangulation without a chord inside a cube face, else a fan around the
  cell-centre vertex, the role of Lewiner's 13th vertex).  NOT restated: Lewiner's interior tests (sub-cases 4.2,
  6.1.2, 7.4.2, 10.1.2, 12.1.2, 13.5.2 -- a tunnel through the cell) and his exact tilings of the joined sub-cases.

The product's tables must equal these (tests/test_mc_cpu.py); the C sweep (oracle/mc_oracle.c) receives these arrays
at run time and shares no source with the product.
"""

CORNER_XYZ = ((0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1))
EDGE_ENDS = ((0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7))
# the six faces, corners counter-clockwise seen from outside the cube, starting at the smallest corner id
FACE_CORNERS = ((0, 4, 7, 3), (1, 2, 6, 5), (0, 1, 5, 4), (2, 3, 7, 6), (0, 3, 2, 1), (4, 5, 6, 7))
CENTRE = 12

# marching squares on one face: pattern bit i = local corner i above the level, local edge i joins local corners
# i and i+1.  Directed segments (from local edge, to local edge) with the region above the level on the left, seen
# from outside the cube.  Patterns 5 and 10 (diagonal corners) have two resolutions: [separated, joined].
SQUARES = {
    0: [], 15: [],
    1: [(0, 3)], 2: [(1, 0)], 4: [(2, 1)], 8: [(3, 2)],
    3: [(1, 3)], 6: [(2, 0)], 12: [(3, 1)], 9: [(0, 2)],
    7: [(2, 3)], 11: [(1, 2)], 13: [(0, 1)], 14: [(3, 0)],
}
SQUARES_AMBIGUOUS = {
    5: ([(0, 3), (2, 1)], [(0, 1), (2, 3)]),
    10: ([(1, 0), (3, 2)], [(3, 0), (1, 2)]),
}

TRITABLE = (
    (), (0, 8, 3), (0, 1, 9), (1, 8, 3, 9, 8, 1), (1, 2, 10), (0, 8, 3, 1, 2, 10), (9, 2, 10, 0, 2, 9),
    (2, 8, 3, 2, 10, 8, 10, 9, 8), (3, 11, 2), (0, 11, 2, 8, 11, 0), (1, 9, 0, 2, 3, 11),
    (1, 11, 2, 1, 9, 11, 9, 8, 11), (3, 10, 1, 11, 10, 3), (0, 10, 1, 0, 8, 10, 8, 11, 10),
    (3, 9, 0, 3, 11, 9, 11, 10, 9), (9, 8, 10, 10, 8, 11), (4, 7, 8), (4, 3, 0, 7, 3, 4), (0, 1, 9, 8, 4, 7),
    (4, 1, 9, 4, 7, 1, 7, 3, 1), (1, 2, 10, 8, 4, 7), (3, 4, 7, 3, 0, 4, 1, 2, 10), (9, 2, 10, 9, 0, 2, 8, 4, 7),
    (2, 10, 9, 2, 9, 7, 2, 7, 3, 7, 9, 4), (8, 4, 7, 3, 11, 2), (11, 4, 7, 11, 2, 4, 2, 0, 4),
    (9, 0, 1, 8, 4, 7, 2, 3, 11), (4, 7, 11, 9, 4, 11, 9, 11, 2, 9, 2, 1), (3, 10, 1, 3, 11, 10, 7, 8, 4),
    (1, 11, 10, 1, 4, 11, 1, 0, 4, 7, 11, 4), (4, 7, 8, 9, 0, 11, 9, 11, 10, 11, 0, 3),
    (4, 7, 11, 4, 11, 9, 9, 11, 10), (9, 5, 4), (9, 5, 4, 0, 8, 3), (0, 5, 4, 1, 5, 0),
    (8, 5, 4, 8, 3, 5, 3, 1, 5), (1, 2, 10, 9, 5, 4), (3, 0, 8, 1, 2, 10, 4, 9, 5), (5, 2, 10, 5, 4, 2, 4, 0, 2),
    (2, 10, 5, 3, 2, 5, 3, 5, 4, 3, 4, 8), (9, 5, 4, 2, 3, 11), (0, 11, 2, 0, 8, 11, 4, 9, 5),
    (0, 5, 4, 0, 1, 5, 2, 3, 11), (2, 1, 5, 2, 5, 8, 2, 8, 11, 4, 8, 5), (10, 3, 11, 10, 1, 3, 9, 5, 4),
    (4, 9, 5, 0, 8, 1, 8, 10, 1, 8, 11, 10), (5, 4, 0, 5, 0, 11, 5, 11, 10, 11, 0, 3),
    (5, 4, 8, 5, 8, 10, 10, 8, 11), (9, 7, 8, 5, 7, 9), (9, 3, 0, 9, 5, 3, 5, 7, 3), (0, 7, 8, 0, 1, 7, 1, 5, 7),
    (1, 5, 3, 3, 5, 7), (9, 7, 8, 9, 5, 7, 10, 1, 2), (10, 1, 2, 9, 5, 0, 5, 3, 0, 5, 7, 3),
    (8, 0, 2, 8, 2, 5, 8, 5, 7, 10, 5, 2), (2, 10, 5, 2, 5, 3, 3, 5, 7), (7, 9, 5, 7, 8, 9, 3, 11, 2),
    (9, 5, 7, 9, 7, 2, 9, 2, 0, 2, 7, 11), (2, 3, 11, 0, 1, 8, 1, 7, 8, 1, 5, 7), (11, 2, 1, 11, 1, 7, 7, 1, 5),
    (9, 5, 8, 8, 5, 7, 10, 1, 3, 10, 3, 11), (5, 7, 0, 5, 0, 9, 7, 11, 0, 1, 0, 10, 11, 10, 0),
    (11, 10, 0, 11, 0, 3, 10, 5, 0, 8, 0, 7, 5, 7, 0), (11, 10, 5, 7, 11, 5), (10, 6, 5), (0, 8, 3, 5, 10, 6),
    (9, 0, 1, 5, 10, 6), (1, 8, 3, 1, 9, 8, 5, 10, 6), (1, 6, 5, 2, 6, 1), (1, 6, 5, 1, 2, 6, 3, 0, 8),
    (9, 6, 5, 9, 0, 6, 0, 2, 6), (5, 9, 8, 5, 8, 2, 5, 2, 6, 3, 2, 8), (2, 3, 11, 10, 6, 5),
    (11, 0, 8, 11, 2, 0, 10, 6, 5), (0, 1, 9, 2, 3, 11, 5, 10, 6), (5, 10, 6, 1, 9, 2, 9, 11, 2, 9, 8, 11),
    (6, 3, 11, 6, 5, 3, 5, 1, 3), (0, 8, 11, 0, 11, 5, 0, 5, 1, 5, 11, 6), (3, 11, 6, 0, 3, 6, 0, 6, 5, 0, 5, 9),
    (6, 5, 9, 6, 9, 11, 11, 9, 8), (5, 10, 6, 4, 7, 8), (4, 3, 0, 4, 7, 3, 6, 5, 10), (1, 9, 0, 5, 10, 6, 8, 4, 7),
    (10, 6, 5, 1, 9, 7, 1, 7, 3, 7, 9, 4), (6, 1, 2, 6, 5, 1, 4, 7, 8), (1, 2, 5, 5, 2, 6, 3, 0, 4, 3, 4, 7),
    (8, 4, 7, 9, 0, 5, 0, 6, 5, 0, 2, 6), (7, 3, 9, 7, 9, 4, 3, 2, 9, 5, 9, 6, 2, 6, 9),
    (3, 11, 2, 7, 8, 4, 10, 6, 5), (5, 10, 6, 4, 7, 2, 4, 2, 0, 2, 7, 11), (0, 1, 9, 4, 7, 8, 2, 3, 11, 5, 10, 6),
    (9, 2, 1, 9, 11, 2, 9, 4, 11, 7, 11, 4, 5, 10, 6), (8, 4, 7, 3, 11, 5, 3, 5, 1, 5, 11, 6),
    (5, 1, 11, 5, 11, 6, 1, 0, 11, 7, 11, 4, 0, 4, 11), (0, 5, 9, 0, 6, 5, 0, 3, 6, 11, 6, 3, 8, 4, 7),
    (6, 5, 9, 6, 9, 11, 4, 7, 9, 7, 11, 9), (10, 4, 9, 6, 4, 10), (4, 10, 6, 4, 9, 10, 0, 8, 3),
    (10, 0, 1, 10, 6, 0, 6, 4, 0), (8, 3, 1, 8, 1, 6, 8, 6, 4, 6, 1, 10), (1, 4, 9, 1, 2, 4, 2, 6, 4),
    (3, 0, 8, 1, 2, 9, 2, 4, 9, 2, 6, 4), (0, 2, 4, 4, 2, 6), (8, 3, 2, 8, 2, 4, 4, 2, 6),
    (10, 4, 9, 10, 6, 4, 11, 2, 3), (0, 8, 2, 2, 8, 11, 4, 9, 10, 4, 10, 6), (3, 11, 2, 0, 1, 6, 0, 6, 4, 6, 1, 10),
    (6, 4, 1, 6, 1, 10, 4, 8, 1, 2, 1, 11, 8, 11, 1), (9, 6, 4, 9, 3, 6, 9, 1, 3, 11, 6, 3),
    (8, 11, 1, 8, 1, 0, 11, 6, 1, 9, 1, 4, 6, 4, 1), (3, 11, 6, 3, 6, 0, 0, 6, 4), (6, 4, 8, 11, 6, 8),
    (7, 10, 6, 7, 8, 10, 8, 9, 10), (0, 7, 3, 0, 10, 7, 0, 9, 10, 6, 7, 10), (10, 6, 7, 1, 10, 7, 1, 7, 8, 1, 8, 0),
    (10, 6, 7, 10, 7, 1, 1, 7, 3), (1, 2, 6, 1, 6, 8, 1, 8, 9, 8, 6, 7), (2, 6, 9, 2, 9, 1, 6, 7, 9, 0, 9, 3, 7, 3, 9),
    (7, 8, 0, 7, 0, 6, 6, 0, 2), (7, 3, 2, 6, 7, 2), (2, 3, 11, 10, 6, 8, 10, 8, 9, 8, 6, 7),
    (2, 0, 7, 2, 7, 11, 0, 9, 7, 6, 7, 10, 9, 10, 7), (1, 8, 0, 1, 7, 8, 1, 10, 7, 6, 7, 10, 2, 3, 11),
    (11, 2, 1, 11, 1, 7, 10, 6, 1, 6, 7, 1), (8, 9, 6, 8, 6, 7, 9, 1, 6, 11, 6, 3, 1, 3, 6), (0, 9, 1, 11, 6, 7),
    (7, 8, 0, 7, 0, 6, 3, 11, 0, 11, 6, 0), (7, 11, 6), (7, 6, 11), (3, 0, 8, 11, 7, 6), (0, 1, 9, 11, 7, 6),
    (8, 1, 9, 8, 3, 1, 11, 7, 6), (10, 1, 2, 6, 11, 7), (1, 2, 10, 3, 0, 8, 6, 11, 7), (2, 9, 0, 2, 10, 9, 6, 11, 7),
    (6, 11, 7, 2, 10, 3, 10, 8, 3, 10, 9, 8), (7, 2, 3, 6, 2, 7), (7, 0, 8, 7, 6, 0, 6, 2, 0),
    (2, 7, 6, 2, 3, 7, 0, 1, 9), (1, 6, 2, 1, 8, 6, 1, 9, 8, 8, 7, 6), (10, 7, 6, 10, 1, 7, 1, 3, 7),
    (10, 7, 6, 1, 7, 10, 1, 8, 7, 1, 0, 8), (0, 3, 7, 0, 7, 10, 0, 10, 9, 6, 10, 7), (7, 6, 10, 7, 10, 8, 8, 10, 9),
    (6, 8, 4, 11, 8, 6), (3, 6, 11, 3, 0, 6, 0, 4, 6), (8, 6, 11, 8, 4, 6, 9, 0, 1),
    (9, 4, 6, 9, 6, 3, 9, 3, 1, 11, 3, 6), (6, 8, 4, 6, 11, 8, 2, 10, 1), (1, 2, 10, 3, 0, 11, 0, 6, 11, 0, 4, 6),
    (4, 11, 8, 4, 6, 11, 0, 2, 9, 2, 10, 9), (10, 9, 3, 10, 3, 2, 9, 4, 3, 11, 3, 6, 4, 6, 3),
    (8, 2, 3, 8, 4, 2, 4, 6, 2), (0, 4, 2, 4, 6, 2), (1, 9, 0, 2, 3, 4, 2, 4, 6, 4, 3, 8), (1, 9, 4, 1, 4, 2, 2, 4, 6),
    (8, 1, 3, 8, 6, 1, 8, 4, 6, 6, 10, 1), (10, 1, 0, 10, 0, 6, 6, 0, 4),
    (4, 6, 3, 4, 3, 8, 6, 10, 3, 0, 3, 9, 10, 9, 3), (10, 9, 4, 6, 10, 4), (4, 9, 5, 7, 6, 11),
    (0, 8, 3, 4, 9, 5, 11, 7, 6), (5, 0, 1, 5, 4, 0, 7, 6, 11), (11, 7, 6, 8, 3, 4, 3, 5, 4, 3, 1, 5),
    (9, 5, 4, 10, 1, 2, 7, 6, 11), (6, 11, 7, 1, 2, 10, 0, 8, 3, 4, 9, 5), (7, 6, 11, 5, 4, 10, 4, 2, 10, 4, 0, 2),
    (3, 4, 8, 3, 5, 4, 3, 2, 5, 10, 5, 2, 11, 7, 6), (7, 2, 3, 7, 6, 2, 5, 4, 9), (9, 5, 4, 0, 8, 6, 0, 6, 2, 6, 8, 7),
    (3, 6, 2, 3, 7, 6, 1, 5, 0, 5, 4, 0), (6, 2, 8, 6, 8, 7, 2, 1, 8, 4, 8, 5, 1, 5, 8),
    (9, 5, 4, 10, 1, 6, 1, 7, 6, 1, 3, 7), (1, 6, 10, 1, 7, 6, 1, 0, 7, 8, 7, 0, 9, 5, 4),
    (4, 0, 10, 4, 10, 5, 0, 3, 10, 6, 10, 7, 3, 7, 10), (7, 6, 10, 7, 10, 8, 5, 4, 10, 4, 8, 10),
    (6, 9, 5, 6, 11, 9, 11, 8, 9), (3, 6, 11, 0, 6, 3, 0, 5, 6, 0, 9, 5), (0, 11, 8, 0, 5, 11, 0, 1, 5, 5, 6, 11),
    (6, 11, 3, 6, 3, 5, 5, 3, 1), (1, 2, 10, 9, 5, 11, 9, 11, 8, 11, 5, 6),
    (0, 11, 3, 0, 6, 11, 0, 9, 6, 5, 6, 9, 1, 2, 10), (11, 8, 5, 11, 5, 6, 8, 0, 5, 10, 5, 2, 0, 2, 5),
    (6, 11, 3, 6, 3, 5, 2, 10, 3, 10, 5, 3), (5, 8, 9, 5, 2, 8, 5, 6, 2, 3, 8, 2), (9, 5, 6, 9, 6, 0, 0, 6, 2),
    (1, 5, 8, 1, 8, 0, 5, 6, 8, 3, 8, 2, 6, 2, 8), (1, 5, 6, 2, 1, 6), (1, 3, 6, 1, 6, 10, 3, 8, 6, 5, 6, 9, 8, 9, 6),
    (10, 1, 0, 10, 0, 6, 9, 5, 0, 5, 6, 0), (0, 3, 8, 5, 6, 10), (10, 5, 6), (11, 5, 10, 7, 5, 11),
    (11, 5, 10, 11, 7, 5, 8, 3, 0), (5, 11, 7, 5, 10, 11, 1, 9, 0), (10, 7, 5, 10, 11, 7, 9, 8, 1, 8, 3, 1),
    (11, 1, 2, 11, 7, 1, 7, 5, 1), (0, 8, 3, 1, 2, 7, 1, 7, 5, 7, 2, 11), (9, 7, 5, 9, 2, 7, 9, 0, 2, 2, 11, 7),
    (7, 5, 2, 7, 2, 11, 5, 9, 2, 3, 2, 8, 9, 8, 2), (2, 5, 10, 2, 3, 5, 3, 7, 5), (8, 2, 0, 8, 5, 2, 8, 7, 5, 10, 2, 5),
    (9, 0, 1, 5, 10, 3, 5, 3, 7, 3, 10, 2), (9, 8, 2, 9, 2, 1, 8, 7, 2, 10, 2, 5, 7, 5, 2), (1, 3, 5, 3, 7, 5),
    (0, 8, 7, 0, 7, 1, 1, 7, 5), (9, 0, 3, 9, 3, 5, 5, 3, 7), (9, 8, 7, 5, 9, 7), (5, 8, 4, 5, 10, 8, 10, 11, 8),
    (5, 0, 4, 5, 11, 0, 5, 10, 11, 11, 3, 0), (0, 1, 9, 8, 4, 10, 8, 10, 11, 10, 4, 5),
    (10, 11, 4, 10, 4, 5, 11, 3, 4, 9, 4, 1, 3, 1, 4), (2, 5, 1, 2, 8, 5, 2, 11, 8, 4, 5, 8),
    (0, 4, 11, 0, 11, 3, 4, 5, 11, 2, 11, 1, 5, 1, 11), (0, 2, 5, 0, 5, 9, 2, 11, 5, 4, 5, 8, 11, 8, 5),
    (9, 4, 5, 2, 11, 3), (2, 5, 10, 3, 5, 2, 3, 4, 5, 3, 8, 4), (5, 10, 2, 5, 2, 4, 4, 2, 0),
    (3, 10, 2, 3, 5, 10, 3, 8, 5, 4, 5, 8, 0, 1, 9), (5, 10, 2, 5, 2, 4, 1, 9, 2, 9, 4, 2), (8, 4, 5, 8, 5, 3, 3, 5, 1),
    (0, 4, 5, 1, 0, 5), (8, 4, 5, 8, 5, 3, 9, 0, 5, 0, 3, 5), (9, 4, 5), (4, 11, 7, 4, 9, 11, 9, 10, 11),
    (0, 8, 3, 4, 9, 7, 9, 11, 7, 9, 10, 11), (1, 10, 11, 1, 11, 4, 1, 4, 0, 7, 4, 11),
    (3, 1, 4, 3, 4, 8, 1, 10, 4, 7, 4, 11, 10, 11, 4), (4, 11, 7, 9, 11, 4, 9, 2, 11, 9, 1, 2),
    (9, 7, 4, 9, 11, 7, 9, 1, 11, 2, 11, 1, 0, 8, 3), (11, 7, 4, 11, 4, 2, 2, 4, 0),
    (11, 7, 4, 11, 4, 2, 8, 3, 4, 3, 2, 4), (2, 9, 10, 2, 7, 9, 2, 3, 7, 7, 4, 9),
    (9, 10, 7, 9, 7, 4, 10, 2, 7, 8, 7, 0, 2, 0, 7), (3, 7, 10, 3, 10, 2, 7, 4, 10, 1, 10, 0, 4, 0, 10),
    (1, 10, 2, 8, 7, 4), (4, 9, 1, 4, 1, 7, 7, 1, 3), (4, 9, 1, 4, 1, 7, 0, 8, 1, 8, 7, 1), (4, 0, 3, 7, 4, 3),
    (4, 8, 7), (9, 10, 8, 10, 11, 8), (3, 0, 9, 3, 9, 11, 11, 9, 10), (0, 1, 10, 0, 10, 8, 8, 10, 11),
    (3, 1, 10, 11, 3, 10), (1, 2, 11, 1, 11, 9, 9, 11, 8), (3, 0, 9, 3, 9, 11, 1, 2, 9, 2, 11, 9), (0, 2, 11, 8, 0, 11),
    (3, 2, 11), (2, 3, 8, 2, 8, 10, 10, 8, 9), (9, 10, 2, 0, 9, 2), (2, 3, 8, 2, 8, 10, 0, 1, 8, 1, 10, 8), (1, 10, 2),
    (1, 3, 8, 9, 1, 8), (0, 9, 1), (0, 3, 8), (),
)
assert len(TRITABLE) == 256


def _edge_between(a, b):
    for e, (p, q) in enumerate(EDGE_ENDS):
        if (p, q) == (a, b) or (q, p) == (a, b):
            return e
    raise KeyError((a, b))


# cube faces each grid edge lies in (a chord between two edge vertices of one face would lie inside that face: the
# neighbouring cell may use the same chord, so the surface would stop being a 2-manifold there)
_FACES_OF_EDGE = [set() for _ in range(12)]
for _f, _c in enumerate(FACE_CORNERS):
    for _i in range(4):
        _FACES_OF_EDGE[_edge_between(_c[_i], _c[(_i + 1) % 4])].add(_f)


def chord_in_a_face(a, b):
    return len(_FACES_OF_EDGE[a] & _FACES_OF_EDGE[b]) > 0


def face_pattern(cfg, f):
    return sum(((cfg >> c) & 1) << i for i, c in enumerate(FACE_CORNERS[f]))


def ambiguous_faces(cfg):
    return [f for f in range(6) if face_pattern(cfg, f) in SQUARES_AMBIGUOUS]


def contour_loops(cfg, joined_bits):
    """Closed loops of intersected grid edges (each a list starting at its smallest edge id, loops ordered by that
    id), directed with the region above the level on the left seen from outside."""
    amb = ambiguous_faces(cfg)
    follow = {}
    for f, corners in enumerate(FACE_CORNERS):
        local_edge = [_edge_between(corners[i], corners[(i + 1) % 4]) for i in range(4)]
        pat = face_pattern(cfg, f)
        if pat in SQUARES_AMBIGUOUS:
            segs = SQUARES_AMBIGUOUS[pat][(joined_bits >> amb.index(f)) & 1]
        else:
            segs = SQUARES[pat]
        for a, b in segs:
            follow[local_edge[a]] = local_edge[b]
    loops, done = [], set()
    for start in sorted(follow):
        if start in done:
            continue
        loop, e = [], start
        while e not in done:
            done.add(e)
            loop.append(e)
            e = follow[e]
        loops.append(loop)
    return loops


def _first_fit(poly):
    """Triangulation rule of the joined variants: split at the first k (1 <= k <= n-2) such that neither chord
    (poly[0], poly[k]) nor (poly[k], poly[-1]) lies inside a cube face and both sides can be triangulated by the
    same rule; the triangle (poly[0], poly[k], poly[-1]) goes between the two sides' triangles."""
    n = len(poly)
    if n == 3:
        return [tuple(poly)]
    for k in range(1, n - 1):
        if (k > 1 and chord_in_a_face(poly[0], poly[k])) or (k < n - 2 and chord_in_a_face(poly[k], poly[-1])):
            continue
        left = _first_fit(poly[:k + 1]) if k > 1 else []
        right = _first_fit(poly[k:]) if k < n - 2 else []
        if left is None or right is None:
            continue
        return left + [(poly[0], poly[k], poly[-1])] + right
    return None


def triangulate_loop(loop):
    for r in range(len(loop)):
        t = _first_fit(loop[r:] + loop[:r])
        if t is not None:
            return t
    return [(CENTRE, loop[i], loop[(i + 1) % len(loop)]) for i in range(len(loop))]


def variant_triangles(cfg, joined_bits):
    if joined_bits == 0:
        row = TRITABLE[cfg]
        return [tuple(row[i:i + 3]) for i in range(0, len(row), 3)]
    out = []
    for loop in contour_loops(cfg, joined_bits):
        out += triangulate_loop(loop)
    return out


def build():
    """-> dict of flat lists the C sweep takes: var_base[256], n_amb[256], amb_face[256*6], n_tri[V], tri[V*36]
    (255-padded), face_corners[24]."""
    var_base, n_amb, amb_face, n_tri, tri = [], [], [], [], []
    for cfg in range(256):
        amb = ambiguous_faces(cfg)
        var_base.append(len(n_tri))
        n_amb.append(len(amb))
        amb_face += amb + [0] * (6 - len(amb))
        for bits in range(1 << len(amb)):
            t = variant_triangles(cfg, bits)
            assert len(t) <= 12
            n_tri.append(len(t))
            flat = [e for x in t for e in x]
            tri += flat + [255] * (36 - len(flat))
    return dict(var_base=var_base, n_amb=n_amb, amb_face=amb_face, n_tri=n_tri, tri=tri,
                face_corners=[c for f in FACE_CORNERS for c in f])
