// Marching cubes on the device, output identical (ids and order) to the sequential sweep.
//
// replaces: skimage.measure.marching_cubes as called by create_mesh (core/reconstruct.py:168-176).
// The sequential algorithm creates a vertex the first time a cell uses a grid edge and appends
// faces cell by cell.  In parallel form:
//   * every intersected edge is used by every cell that contains it, so the creating cell is the
//     lowest-index cell adjacent to the edge -> a 12-bit ownership mask from (x==0, y==0, z==0);
//   * vertex id = exclusive scan over cells (sweep order) of owned-vertex counts + rank of the
//     edge among the cell's owned edges in first-appearance order of its triangle list;
//   * face offset = exclusive scan of triangle counts.
// Kernels (round 2): the volume is read ONCE.
//   classify_kernel   the unit of work and of the scan is a warp's 31 consecutive cells of one row (x fastest; lane 31
//                     only carries the samples at x + 1 of lane 30's cell); warps never synchronise with each other.
//                     A lane loads its samples at x and takes what it needs of x + 1 from its neighbour (one shuffle
//                     of the packed "above the level" flags); a warp pass covers 4 rows (10 coalesced reads in flight).
//                     ~97 % of the cells are not cut: they cost 8 compares and nothing is stored for them.  A cut
//                     cell gets a RECORD (cell, variant, its exclusive (vertex, triangle) prefix inside the unit,
//                     the 8 corner values) in a record list; the unit reserves its run with one atomic on one of
//                     NARENA cursors (unit index mod NARENA: a single cursor would serialise ~100,000 returning
//                     atomics on one L2 address).  Per live unit: counts.
//   scan_*            two-level exclusive scan of the per-unit (vertex, triangle) counts in sweep order; totals for
//                     the host.  (The volume's min / max -- skimage's level-range check -- is only computed when no
//                     cell is cut: a cut cell proves the level is inside the range.)
//   emit_vertices /   dense over the records, in the order they were stored (NOT sweep order: a record carries its
//   emit_faces        block and prefix, so its global ids are block offset + prefix whatever the storage order):
//                     one thread per cut cell, 1.5 owned vertices / 2.8 triangles on average.
// HBM-bound: 4 B/sample read once + 12 B per vertex + 12 B per face written; the edge -> id table (4 int per sample,
// touched only at intersected edges) is the only scattered traffic.
#pragma once
#include "common.cuh"

#define DJ_MC_QUAL __device__ const
#include "mc_tables.h"

namespace dj {
namespace mc {

constexpr int CB = 256;  // cells per block (consecutive in x, one row)
// variants of the two configurations without a surface (DJ_MC_VARBASE[0], DJ_MC_VARBASE[255]; no ambiguous faces)
constexpr int DJ_MC_VAR_EMPTY0 = 0, DJ_MC_VAR_EMPTY255 = 655;
// one cut cell: what the emit kernels need, so that they never touch the volume
struct __align__(16) Record {
  int cell;     // (zr * c1 + y) * c2 + x within the slab
  int unit;     // sweep-order index of the cell's 32-cell unit (its scanned offsets are the base of the ids)
  int var;      // case-table variant
  int pre;      // exclusive prefix inside the unit: owned vertices | triangles << 16
  float f[8];   // corner values (float32 as stored in the volume)
};
constexpr int NARENA = 1024;  // the record list is NARENA regions of arena_cap records, one cursor each
// edges that are owned only on the x==0 / y==0 / z==0 boundary of the slab
constexpr unsigned MX = 0x988, MY = 0x311, MZ = 0x00F;
constexpr double FLT_EPS_D = 1.1920928955078125e-07;

struct Dims {
  int n0, n1, n2;   // samples of this slab (n0 planes along axis 0)
  int z_off;        // global index of the slab's first plane (added to axis-0 coordinates)
  int seam;         // 1: upper slab of a sharded sweep -- plane-0 edges belong to the slab below
  int c1, c2;       // cells along axis 1 / 2
  int wpr;          // units (UC = 31 cells) per row: ceil(c2 / 31)
  int64_t ncells;
  double level;
  float level_lo;   // the largest float32 <= level: for a float32 sample v, ((double)v - level > 0) == (v > level_lo) exactly
};


// variant of a cut cell: configuration + Lewiner's test_face on its ambiguous faces.  v: float32 corner values;
// the differences to the level are formed in float64 (as the sweep does) only for the corners of ambiguous faces
__device__ __forceinline__ int cube_variant(int cfg, const float (&v)[8], double level) {
  int var = DJ_MC_VARBASE[cfg];
  const int na = DJ_MC_NAMB[cfg];
  for (int a = 0; a < na; a++) {
    const unsigned char* c = DJ_MC_FACE[DJ_MC_AMBFACE[cfg][a]];
    // with A, C one diagonal and B, D the other, the corners above the level are joined through the face when
    // |AC - BD| < FLT_EPSILON (Lewiner's dead band) or the face-centre saddle (AC - BD) / (A + C - B - D) is
    // positive; products and difference rounded separately (no fused multiply-add), as the C library computes them
    const double f0 = __dsub_rn((double)v[c[0]], level), f1 = __dsub_rn((double)v[c[1]], level);
    const double f2 = __dsub_rn((double)v[c[2]], level), f3 = __dsub_rn((double)v[c[3]], level);
    const double det = __dsub_rn(__dmul_rn(f0, f2), __dmul_rn(f1, f3));
    const bool joined = (fabs(det) < FLT_EPS_D) || (((cfg >> c[0]) & 1) ? (det > 0.0) : (det < 0.0));
    var += joined ? (1 << a) : 0;
  }
  return var;
}

// bit 12 = the cell-centre vertex, always created by its own cell
__device__ __forceinline__ unsigned owned_mask(const Dims& D, int zr, int y, int x) {
  return 0x1FFFu & ~(((zr | D.seam) ? MZ : 0u) | (y ? MY : 0u) | (x ? MX : 0u));
}

__device__ __forceinline__ int owned_count(int var, unsigned own) { return __popc(own & (unsigned)DJ_MC_EMASK[var]); }

// ---- pass 1: classify every cell, record the cut ones, per-unit counts, min / max.
// The unit of the scan is a WARP's 31 cells of one row (sweep order: layer, row, chunk): warps never
// synchronise with each other.  A warp pass handles ROWS consecutive rows at once: ROWS + 1 sample rows of two planes,
// 2 (ROWS + 1) independent coalesced 128-byte reads in flight per warp, each sample row shared by the two cell rows
// that touch it; the samples at x + 1 come from the neighbouring lane.
constexpr int ROWS = 4;
constexpr int UC = 31;  // cells of a unit: lane 31 of the warp only carries the samples at x + 1 of lane 30's cell
__device__ __forceinline__ unsigned pack_counts(int n_rec, int nv, int nt) { return (unsigned)n_rec | ((unsigned)nv << 8) | ((unsigned)nt << 20); }
__device__ __forceinline__ int unpack_nv(unsigned p) { return (int)((p >> 8) & 0xFFFu); }
__device__ __forceinline__ int unpack_nt(unsigned p) { return (int)(p >> 20); }
__device__ __forceinline__ unsigned arena_of(int unit) { return ((unsigned)unit * 2654435761u) >> 22; }  // 10 bits: NARENA = 1024

// min / max of the volume for skimage's level-range ValueError: only needed when NO cell is cut (a cut cell has
// samples on both sides of the level, so the level is inside the range), i.e. on the error path of dj_mc_count
__global__ void minmax_kernel(const float* __restrict__ vol, int64_t n, int* __restrict__ mm /*[2] ordered ints*/) {
  float lo = INFINITY, hi = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = __ldg(vol + i);
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    auto ord = [](float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; };
    atomicMin(mm, ord(lo));
    atomicMax(mm + 1, ord(hi));
  }
}

__global__ void __launch_bounds__(CB) classify_kernel(Dims D, const float* __restrict__ vol, uint2* __restrict__ units /* zeroed */,
                                                      Record* __restrict__ records, int arena_cap, int* __restrict__ cursors) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // grid: x = blocks along the row (8 units each), y = groups of ROWS rows, z = cell layer
  const int zr = (int)blockIdx.z;
  const int y0 = (int)blockIdx.y * ROWS;
  const int xw = (int)blockIdx.x * (CB / 32) + w;  // unit of the row
  if (xw >= D.wpr) return;
  const int x = xw * UC + lane;
  const bool has_sample = x < D.n2, has_cell = lane < UC && x < D.c2;
  const int64_t sy = D.n2, sz = (int64_t)D.n1 * D.n2;
  const float* p = vol + ((int64_t)zr * D.n1 + y0) * D.n2 + x;
  // samples at x: rows y0 .. y0 + ROWS of planes z (k = 0) and z + 1 (k = 1)
  float a[ROWS + 1][2];
#pragma unroll
  for (int r = 0; r <= ROWS; r++) {
    a[r][0] = 0.f; a[r][1] = 0.f;
    if (has_sample && y0 + r < D.n1) { a[r][0] = __ldg(p + r * sy); a[r][1] = __ldg(p + sz + r * sy); }
  }
  // "above the level" flags of my samples: v > level_lo  <=>  sign bit of (level_lo - v) (v == level_lo gives +0).
  // Bit 2 r + k of `mine`; the flags of the samples at x + 1 come from the next lane in ONE shuffle.
  unsigned mine = 0u;
#pragma unroll
  for (int r = ROWS; r >= 0; r--) {
    mine = __funnelshift_l(__float_as_uint(D.level_lo - a[r][1]), mine, 1);
    mine = __funnelshift_l(__float_as_uint(D.level_lo - a[r][0]), mine, 1);
  }
  const unsigned next = __shfl_down_sync(0xffffffffu, mine, 1);
#pragma unroll
  for (int r = 0; r < ROWS; r++) {
    const int y = y0 + r;
    if (y >= D.c1) break;  // warp-uniform
    // flags of the cell's 8 corners: m = mine >> 2r -> bit0 (y, z) bit1 (y, z+1) bit2 (y+1, z) bit3 (y+1, z+1) at x,
    // n = next >> 2r the same at x + 1.  The cell is cut iff they are not all equal; ~85 % of the units have no
    // cut cell and stop here.
    const unsigned m = (mine >> (2 * r)) & 15u, n = (next >> (2 * r)) & 15u;
    const unsigned all8 = m | (n << 4);
    const bool cut = has_cell && all8 != 0u && all8 != 255u;
    const unsigned cutmask = __ballot_sync(0xffffffffu, cut);
    if (cutmask == 0u) continue;
    // ---- a unit with cut cells: configuration in Lewiner's corner order (0,0,0) (1,0,0) (1,1,0) (0,1,0) and the
    // same at z + 1, and the values at x + 1
    const int cfg = (int)((m & 1u) | ((n & 1u) << 1) | (((n >> 2) & 1u) << 2) | (((m >> 2) & 1u) << 3) |
                          (((m >> 1) & 1u) << 4) | (((n >> 1) & 1u) << 5) | (((n >> 3) & 1u) << 6) | (((m >> 3) & 1u) << 7));
    const float b0 = __shfl_down_sync(0xffffffffu, a[r][0], 1), b1 = __shfl_down_sync(0xffffffffu, a[r][1], 1);
    const float b2 = __shfl_down_sync(0xffffffffu, a[r + 1][0], 1), b3 = __shfl_down_sync(0xffffffffu, a[r + 1][1], 1);
    const float v[8] = {a[r][0], b0, b2, a[r + 1][0], a[r][1], b1, b3, a[r + 1][1]};
    int var = 0, nv = 0, nt = 0;
    if (cut) {
      var = cube_variant(cfg, v, D.level);
      nt = DJ_MC_NTRI[var];
      nv = owned_count(var, owned_mask(D, zr, y, x));
    }
    const int packed = nv | (nt << 16);
    int inc = packed;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    const int tot = __shfl_sync(0xffffffffu, inc, 31);
    const int n_rec = __popc(cutmask);
    const int unit = (zr * D.c1 + y) * D.wpr + xw;
    const unsigned arena = arena_of(unit);
    int base = 0;
    if (lane == 0) {
      base = atomicAdd(cursors + arena, n_rec);  // this unit's run inside its arena
      units[unit] = make_uint2(pack_counts(n_rec, tot & 0xFFFF, tot >> 16), (unsigned)base);
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (cut) {
      const int k = base + __popc(cutmask & ((1u << lane) - 1u));
      if (k < arena_cap) {  // otherwise the host grows the arenas and classifies again (dj_mc_count)
        Record rec;
        rec.cell = (zr * D.c1 + y) * D.c2 + x;
        rec.unit = unit;
        rec.var = var;
        rec.pre = inc - packed;
#pragma unroll
        for (int i = 0; i < 8; i++) rec.f[i] = v[i];
        records[(int64_t)arena * arena_cap + k] = rec;
      }
    }
  }
}

// ---- pass 2: exclusive scan of the per-unit (vertex, triangle) counts in two levels, sweep order.
// level 1: every CTA scans 1024 units (-> offsets within its group) and writes the group totals; level 2: one CTA
// scans the group totals and writes the grand totals for the host.  Consumers add offsets[u] + group_off[u >> 10].
constexpr int SG = 1024;
__device__ __forceinline__ void scan1024(long long& a, long long& b, long long (*wsum)[32]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long a0 = a, b0 = b;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long ta = __shfl_up_sync(0xffffffffu, a, o), tb = __shfl_up_sync(0xffffffffu, b, o);
    if (lane >= o) { a += ta; b += tb; }
  }
  if (lane == 31) { wsum[0][w] = a; wsum[1][w] = b; }
  __syncthreads();
  if (w == 0) {
    long long sa = wsum[0][lane], sb = wsum[1][lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long ta = __shfl_up_sync(0xffffffffu, sa, o), tb = __shfl_up_sync(0xffffffffu, sb, o);
      if (lane >= o) { sa += ta; sb += tb; }
    }
    wsum[0][lane] = sa;
    wsum[1][lane] = sb;
  }
  __syncthreads();
  a = (w ? wsum[0][w - 1] : 0) + a - a0;  // exclusive
  b = (w ? wsum[1][w - 1] : 0) + b - b0;
}
__global__ void __launch_bounds__(SG) scan_groups_kernel(const uint2* __restrict__ units, int64_t nunits, int2* __restrict__ offsets,
                                                         longlong2* __restrict__ group_tot) {
  __shared__ long long wsum[2][32];
  const int64_t i = (int64_t)blockIdx.x * SG + threadIdx.x;
  const unsigned pc = (i < nunits) ? units[i].x : 0u;
  const long long cv = unpack_nv(pc), ct = unpack_nt(pc);
  long long a = cv, b = ct;
  scan1024(a, b, wsum);
  if (i < nunits) offsets[i] = make_int2((int)a, (int)b);
  if (threadIdx.x == SG - 1) group_tot[blockIdx.x] = make_longlong2(a + cv, b + ct);
}
__global__ void __launch_bounds__(SG) scan_totals_kernel(longlong2* __restrict__ group_tot, int64_t ngroups,
                                                         long long* __restrict__ totals) {
  __shared__ long long wsum[2][32];
  __shared__ long long carry[2];
  if (threadIdx.x == 0) { carry[0] = 0; carry[1] = 0; }
  __syncthreads();
  for (int64_t base = 0; base < ngroups; base += SG) {
    const int64_t i = base + threadIdx.x;
    const longlong2 c = (i < ngroups) ? group_tot[i] : make_longlong2(0, 0);
    long long a = c.x, b = c.y;
    scan1024(a, b, wsum);
    if (i < ngroups) group_tot[i] = make_longlong2(carry[0] + a, carry[1] + b);
    __syncthreads();
    if (threadIdx.x == SG - 1) { carry[0] += a + c.x; carry[1] += b + c.y; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { totals[0] = carry[0]; totals[1] = carry[1]; }
}

// index of the edge->vertex-id slot of edge e of cell (zr, y, x): 4 slots per sample (x-, y-, z-edge
// starting at that sample, centre vertex of the cell whose origin it is), samples slab-relative
constexpr int SLOTS = 4;
__device__ __forceinline__ void cell_coords(const Dims& D, int cell, int& zr, int& y, int& x) {
  x = cell % D.c2;
  const int q = cell / D.c2;
  y = q % D.c1;
  zr = q / D.c1;
}
__device__ __forceinline__ int64_t edge_slot(const Dims& D, int zr, int y, int x, int e) {
  if (e == 12) return SLOTS * (((int64_t)zr * D.n1 + y) * D.n2 + x) + 3;
  const int a = DJ_MC_EDGE[e][0], b = DJ_MC_EDGE[e][1];
  const int ax = DJ_MC_CORNER[a][0], ay = DJ_MC_CORNER[a][1], az = DJ_MC_CORNER[a][2];
  const int bx = DJ_MC_CORNER[b][0], by = DJ_MC_CORNER[b][1], bz = DJ_MC_CORNER[b][2];
  const int dir = (ax != bx) ? 0 : ((ay != by) ? 1 : 2);
  const int64_t p = ((int64_t)(zr + min(az, bz)) * D.n1 + (y + min(ay, by))) * D.n2 + (x + min(ax, bx));
  return SLOTS * p + dir;
}

// offset of the vertex on edge e from the cell origin, (x,y,z) order: inverse-distance weights
// 1/(FLT_EPSILON + |v - level|) in float64, explicit roundings (no fused multiply-add)
__device__ __forceinline__ void edge_frac(const double (&f)[8], int e, double (&fr)[3]) {
  const int a = DJ_MC_EDGE[e][0], b = DJ_MC_EDGE[e][1];
  const double w1 = __ddiv_rn(1.0, __dadd_rn(FLT_EPS_D, fabs(f[a])));
  const double w2 = __ddiv_rn(1.0, __dadd_rn(FLT_EPS_D, fabs(f[b])));
  const double ws = __dadd_rn(w1, w2);
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const int ca = DJ_MC_CORNER[a][d], cb = DJ_MC_CORNER[b][d];
    fr[d] = (ca == cb) ? (double)ca : __ddiv_rn(cb ? w2 : w1, ws);
  }
}

// ---- pass 3: vertices (ids in sweep first-use order) + edge -> id table.  One thread per cut cell, dense.
constexpr int EB = 128;  // records per block of the emit kernels
// thread i <-> entry i % span of arena i / span (span = the fullest arena, rounded up: live if below the arena's cursor)
__device__ __forceinline__ int64_t live_record(int64_t i, int span, int arena_cap, const int* __restrict__ cursors) {
  const int a = (int)(i / span), k = (int)(i - (int64_t)a * span);
  return (a < NARENA && k < __ldg(cursors + a)) ? (int64_t)a * arena_cap + k : -1;
}
__global__ void __launch_bounds__(EB) emit_vertices_kernel(Dims D, const Record* __restrict__ records, int span, int arena_cap,
                                                           const int* __restrict__ cursors, const int2* __restrict__ block_off,
                                                           const longlong2* __restrict__ group_off, int* __restrict__ edge_id,
                                                           float* __restrict__ verts) {
  const int64_t i = live_record((int64_t)blockIdx.x * EB + threadIdx.x, span, arena_cap, cursors);
  if (i < 0) return;
  const Record r = records[i];
  int zr, y, x;
  cell_coords(D, r.cell, zr, y, x);
  const unsigned own = owned_mask(D, zr, y, x);
  int64_t id = (int64_t)block_off[r.unit].x + group_off[r.unit >> 10].x + (r.pre & 0xFFFF);
  double f[8];
#pragma unroll
  for (int k = 0; k < 8; k++) f[k] = __dsub_rn((double)r.f[k], D.level);
  const int ne = DJ_MC_NEDGE[r.var];
  for (int k = 0; k < ne; k++) {
    const int e = DJ_MC_ORDER[r.var][k];
    if (!((own >> e) & 1u)) continue;
    double fr[3];
    if (e == 12) {
      // cell-centre vertex: mean of the cell's edge vertices, accumulated in edge-id order
      double sum[3] = {0.0, 0.0, 0.0};
      int cnt = 0;
      for (int ee = 0; ee < 12; ee++) {
        const int a = DJ_MC_EDGE[ee][0], b = DJ_MC_EDGE[ee][1];
        if ((f[a] > 0.0) == (f[b] > 0.0)) continue;
        double t[3];
        edge_frac(f, ee, t);
#pragma unroll
        for (int d = 0; d < 3; d++) sum[d] = __dadd_rn(sum[d], t[d]);
        cnt++;
      }
#pragma unroll
      for (int d = 0; d < 3; d++) fr[d] = __ddiv_rn(sum[d], (double)cnt);
    } else {
      edge_frac(f, e, fr);
    }
    const int base[3] = {x, y, D.z_off + zr};
    float p[3];
#pragma unroll
    for (int d = 0; d < 3; d++) p[d] = (float)__dadd_rn((double)base[d], fr[d]);
    verts[3 * id + 0] = p[2];  // (axis0, axis1, axis2)
    verts[3 * id + 1] = p[1];
    verts[3 * id + 2] = p[0];
    edge_id[edge_slot(D, zr, y, x, e)] = (int)id;
    id++;
  }
}

// seam mode: references to plane-0 x/y edges resolve to -2 - (3*sample + dir) (patched by the gather step)
__global__ void seam_fill_kernel(int* __restrict__ edge_id, int64_t nsamples_plane) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nsamples_plane * 2) {
    const int64_t p = i >> 1;
    const int dir = (int)(i & 1);
    edge_id[SLOTS * p + dir] = (int)(-2 - (3 * p + dir));
  }
}

// top plane of a slab: ids of its x/y edges, 3 ints per sample (slot 2 unused)
__global__ void top_plane_kernel(const int* __restrict__ edge_id_plane, int* __restrict__ out, int64_t nsamples_plane) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nsamples_plane * 3) {
    const int64_t p = i / 3;
    const int dir = (int)(i % 3);
    out[i] = dir < 2 ? edge_id_plane[SLOTS * p + dir] : -1;
  }
}

// ---- pass 4: faces, appended cell by cell in table order.  One thread per cut cell, dense.
__global__ void __launch_bounds__(EB) emit_faces_kernel(Dims D, const Record* __restrict__ records, int span, int arena_cap,
                                                        const int* __restrict__ cursors, const int2* __restrict__ block_off,
                                                        const longlong2* __restrict__ group_off,
                                                        const int* __restrict__ edge_id, int descent, int* __restrict__ faces) {
  const int64_t i = live_record((int64_t)blockIdx.x * EB + threadIdx.x, span, arena_cap, cursors);
  if (i < 0) return;
  const int cell = records[i].cell, block = records[i].unit, var = records[i].var, pre = records[i].pre;
  int zr, y, x;
  cell_coords(D, cell, zr, y, x);
  const int64_t f0 = (int64_t)block_off[block].y + group_off[block >> 10].y + (pre >> 16);
  const int nt = DJ_MC_NTRI[var];
  for (int t = 0; t < nt; t++) {
    int id[3];
#pragma unroll
    for (int k = 0; k < 3; k++) id[k] = edge_id[edge_slot(D, zr, y, x, DJ_MC_TRI[var][3 * t + k])];
    int* o = faces + 3 * (f0 + t);
    if (descent) { o[0] = id[2]; o[1] = id[1]; o[2] = id[0]; }
    else { o[0] = id[0]; o[1] = id[1]; o[2] = id[2]; }
  }
}

}  // namespace mc
}  // namespace dj
