/* CPU restatement of the marching-cubes step of create_mesh()
 * (reference core/reconstruct.py:162-176 -> skimage.measure.marching_cubes,
 *  scikit-image 0.19.3 `_marching_cubes_lewiner_cy.pyx`, NOT present offline).
 *
 * TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED against scikit-image: what is
 * restated from its published behaviour is
 *   - the sequential sweep z(axis0) -> y(axis1) -> x(axis2, fastest),
 *   - cube index bit i = (v_i - level > 0) with Lewiner corner numbering,
 *   - one vertex per intersected grid edge, created on FIRST USE (ids in
 *     first-use order), faces appended cell by cell in table order,
 *   - vertex position = inverse-distance weights 1/(FLT_EPSILON+|v-level|) in
 *     double, stored as float32, columns (axis0, axis1, axis2) in index units,
 *   - 'descent' reverses each face, level outside [min,max] / no surface errors.
 *   - ambiguous faces: Lewiner's test_face -- corners above the level joined across a face iff
 *     |AC - BD| < FLT_EPSILON or the face-centre saddle value is positive.
 * The case tables are NOT the product's: they are derived by oracle/mc_case_tables.py (its own statement of the
 * Lorensen-Cline table and of the joined variants) and handed to this sweep at run time (dj_oracle_mc_set_tables).
 * Not restated: Lewiner's interior tests (tunnel sub-cases) and his tilings of the joined sub-cases.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (oracle/build_oracle.py).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Lewiner numbering, written out here (nothing is shared with the product's sources) */
static const unsigned char DJ_MC_CORNER[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
static const unsigned char DJ_MC_EDGE[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
/* case tables, set by dj_oracle_mc_set_tables (copied) */
static int *T_VARBASE, *T_NAMB, *T_AMBFACE, *T_NTRI, *T_TRI, *T_FACE;
#define DJ_MC_VARBASE T_VARBASE
#define DJ_MC_NAMB T_NAMB
#define DJ_MC_AMBFACE(cfg, a) T_AMBFACE[(cfg) * 6 + (a)]
#define DJ_MC_NTRI T_NTRI
#define DJ_MC_TRI(var, i) T_TRI[(var) * 36 + (i)]

static int *dup_ints(const int *src, int n) {
  int *d = (int *)malloc(sizeof(int) * (size_t)n);
  memcpy(d, src, sizeof(int) * (size_t)n);
  return d;
}
void dj_oracle_mc_set_tables(const int *var_base, const int *n_amb, const int *amb_face, const int *n_tri,
                             const int *tri, int n_variants, const int *face_corners) {
  free(T_VARBASE); free(T_NAMB); free(T_AMBFACE); free(T_NTRI); free(T_TRI); free(T_FACE);
  T_VARBASE = dup_ints(var_base, 256);
  T_NAMB = dup_ints(n_amb, 256);
  T_AMBFACE = dup_ints(amb_face, 256 * 6);
  T_NTRI = dup_ints(n_tri, n_variants);
  T_TRI = dup_ints(tri, n_variants * 36);
  T_FACE = dup_ints(face_corners, 24);
}

typedef struct { float *v; int32_t *f; int64_t nv, nf, capv, capf; } Mesh;

static void push_v(Mesh *m, float a, float b, float c) {
  if (m->nv == m->capv) { m->capv = m->capv ? m->capv * 2 : 4096; m->v = (float *)realloc(m->v, m->capv * 3 * sizeof(float)); }
  m->v[m->nv * 3 + 0] = a; m->v[m->nv * 3 + 1] = b; m->v[m->nv * 3 + 2] = c; m->nv++;
}
static void push_f(Mesh *m, int32_t a, int32_t b, int32_t c) {
  if (m->nf == m->capf) { m->capf = m->capf ? m->capf * 2 : 8192; m->f = (int32_t *)realloc(m->f, m->capf * 3 * sizeof(int32_t)); }
  m->f[m->nf * 3 + 0] = a; m->f[m->nf * 3 + 1] = b; m->f[m->nf * 3 + 2] = c; m->nf++;
}

/* joined (positive diagonal connected through the face centre)? products of the
 * two diagonals, each rounded once; no fused multiply-add. */
static int face_joined(const double *f, int face, int cfg) {
  const int *c = T_FACE + 4 * face;
  double p02 = f[c[0]] * f[c[2]];
  double p13 = f[c[1]] * f[c[3]];
  double det = p02 - p13;
  if (fabs(det) < (double)FLT_EPSILON) return 1; /* Lewiner's dead band: test_face(+f) true, test_face(-f) false */
  if ((cfg >> c[0]) & 1) return det > 0; /* corners 0,2 are the diagonal above the level */
  return det < 0;
}

/* position of the vertex on edge e relative to the cell origin, (x,y,z) order:
 * inverse-distance weights 1/(FLT_EPSILON+|v-level|); out[d] = (accumulate ? out[d] : 0) + frac */
static void edge_frac(const double *f, int e, double eps, double *out, int accumulate) {
  int a = DJ_MC_EDGE[e][0], b = DJ_MC_EDGE[e][1];
  const unsigned char *ca = DJ_MC_CORNER[a], *cb = DJ_MC_CORNER[b];
  double w1 = 1.0 / (eps + fabs(f[a])), w2 = 1.0 / (eps + fabs(f[b]));
  double ws = w1 + w2;
  for (int d = 0; d < 3; d++) {
    double fr = (ca[d] == cb[d]) ? (double)ca[d] : (cb[d] ? w2 : w1) / ws;
    out[d] = accumulate ? out[d] + fr : fr;
  }
}

int dj_oracle_mc_variant(const double *f) {
  int cfg = 0;
  for (int i = 0; i < 8; i++) if (f[i] > 0) cfg |= 1 << i;
  int var = DJ_MC_VARBASE[cfg], bits = 0;
  for (int a = 0; a < DJ_MC_NAMB[cfg]; a++)
    if (face_joined(f, DJ_MC_AMBFACE(cfg, a), cfg)) bits |= 1 << a;
  return var + bits;
}

/* returns 0 ok, 4 no surface, 5 level outside data range, 1 bad args */
int dj_oracle_mc(const float *vol, int n0, int n1, int n2, double level, int descent,
                 float **verts, int32_t **faces, int64_t *nV, int64_t *nF) {
  *verts = 0; *faces = 0; *nV = 0; *nF = 0;
  if (n0 < 2 || n1 < 2 || n2 < 2 || !T_TRI) return 1;
  {
    float mn = vol[0], mx = vol[0];
    for (int64_t i = 0; i < (int64_t)n0 * n1 * n2; i++) { if (vol[i] < mn) mn = vol[i]; if (vol[i] > mx) mx = vol[i]; }
    if (level < (double)mn || level > (double)mx) return 5;
  }
  Mesh m; memset(&m, 0, sizeof m);
  const int64_t plane = (int64_t)n1 * n2;
  /* edge -> vertex id, two z-planes for x/y edges, one layer for z edges */
  int32_t *xe[2], *ye[2], *ze;
  for (int k = 0; k < 2; k++) { xe[k] = (int32_t *)malloc(plane * 4); ye[k] = (int32_t *)malloc(plane * 4); }
  ze = (int32_t *)malloc(plane * 4);
  memset(xe[1], 0xff, plane * 4); memset(ye[1], 0xff, plane * 4);
  const double eps = (double)FLT_EPSILON;
  for (int z = 0; z < n0 - 1; z++) {
    { int32_t *t = xe[0]; xe[0] = xe[1]; xe[1] = t; t = ye[0]; ye[0] = ye[1]; ye[1] = t; }
    memset(xe[1], 0xff, plane * 4); memset(ye[1], 0xff, plane * 4); memset(ze, 0xff, plane * 4);
    for (int y = 0; y < n1 - 1; y++) for (int x = 0; x < n2 - 1; x++) {
      double f[8];
      for (int i = 0; i < 8; i++) {
        const unsigned char *c = DJ_MC_CORNER[i];
        f[i] = (double)vol[((int64_t)(z + c[2]) * n1 + (y + c[1])) * n2 + (x + c[0])] - level;
      }
      int var = dj_oracle_mc_variant(f);
      int nt = DJ_MC_NTRI[var];
      int32_t centre_id = -1;
      for (int t = 0; t < nt; t++) {
        int32_t id[3];
        for (int k = 0; k < 3; k++) {
          int e = DJ_MC_TRI(var, t * 3 + k);
          int base[3] = {x, y, z};
          if (e == 12) { /* cell-centre vertex: mean of the cell's edge vertices (edge id order) */
            if (centre_id < 0) {
              double sum[3] = {0, 0, 0};
              int cnt = 0;
              for (int ee = 0; ee < 12; ee++) {
                int a = DJ_MC_EDGE[ee][0], b = DJ_MC_EDGE[ee][1];
                if ((f[a] > 0) == (f[b] > 0)) continue;
                edge_frac(f, ee, eps, sum, 1);
                cnt++;
              }
              centre_id = (int32_t)m.nv;
              push_v(&m, (float)((double)base[2] + sum[2] / (double)cnt), (float)((double)base[1] + sum[1] / (double)cnt),
                     (float)((double)base[0] + sum[0] / (double)cnt));
            }
            id[k] = centre_id;
            continue;
          }
          int a = DJ_MC_EDGE[e][0], b = DJ_MC_EDGE[e][1];
          const unsigned char *ca = DJ_MC_CORNER[a], *cb = DJ_MC_CORNER[b];
          int32_t *slot;
          if (ca[0] != cb[0]) slot = &xe[ca[2]][(int64_t)(y + ca[1]) * n2 + x];
          else if (ca[1] != cb[1]) slot = &ye[ca[2]][(int64_t)y * n2 + (x + ca[0])];
          else slot = &ze[(int64_t)(y + ca[1]) * n2 + (x + ca[0])];
          if (*slot < 0) {
            double p[3] = {0, 0, 0};
            edge_frac(f, e, eps, p, 0);
            *slot = (int32_t)m.nv;
            push_v(&m, (float)((double)base[2] + p[2]), (float)((double)base[1] + p[1]),
                   (float)((double)base[0] + p[0])); /* (axis0, axis1, axis2) */
          }
          id[k] = *slot;
        }
        if (descent) push_f(&m, id[2], id[1], id[0]); else push_f(&m, id[0], id[1], id[2]);
      }
    }
  }
  for (int k = 0; k < 2; k++) { free(xe[k]); free(ye[k]); }
  free(ze);
  if (m.nv == 0) { free(m.v); free(m.f); return 4; }
  *verts = m.v; *faces = m.f; *nV = m.nv; *nF = m.nf;
  return 0;
}

void dj_oracle_free(void *p) { free(p); }
