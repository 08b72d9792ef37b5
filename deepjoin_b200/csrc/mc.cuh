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
// Kernels: classify (cube index + asymptotic-decider variant, per-block counts, warp-shuffle
// reduction) -> two-level scan of the block counts -> emit vertices (+ edge->id table) -> emit faces.
// A block owns 256 consecutive cells of one cell layer (blockIdx.y = layer, so the sweep-order block
// index needs no 64-bit division).  Only ~3 % of the cells are cut by the surface: the emit kernels
// compact the block's owned vertices / triangles into shared memory with the block scan and then give
// every thread ONE vertex (one inverse-distance interpolation in float64) or ONE triangle, instead of
// one cell with up to 13 vertices.  HBM-bound: 4 B/sample read by classify; the emit passes stream the
// 2 B/cell variant array.
#pragma once
#include "common.cuh"

#define DJ_MC_QUAL __device__ const
#include "mc_tables.h"

namespace dj {
namespace mc {

constexpr int CB = 256;  // cells per block
// variants of the two configurations without a surface (DJ_MC_VARBASE[0], DJ_MC_VARBASE[255]; no ambiguous faces)
constexpr int DJ_MC_VAR_EMPTY0 = 0, DJ_MC_VAR_EMPTY255 = 655;
// edges that are owned only on the x==0 / y==0 / z==0 boundary of the slab
constexpr unsigned MX = 0x988, MY = 0x311, MZ = 0x00F;
constexpr double FLT_EPS_D = 1.1920928955078125e-07;

struct Dims {
  int n0, n1, n2;   // samples of this slab (n0 planes along axis 0)
  int z_off;        // global index of the slab's first plane (added to axis-0 coordinates)
  int seam;         // 1: upper slab of a sharded sweep -- plane-0 edges belong to the slab below
  int c1, c2;       // cells along axis 1 / 2
  int64_t ncells;
  double level;
};

// block -> cells: blockIdx.y = cell layer zr, blockIdx.x*CB + threadIdx.x = cell within the layer (row-major y, x)
__device__ __forceinline__ bool block_cell(const Dims& D, int t, int& zr, int& y, int& x, int64_t& ci) {
  zr = (int)blockIdx.y;
  const unsigned idx = blockIdx.x * (unsigned)CB + (unsigned)t;
  const unsigned plane = (unsigned)D.c1 * (unsigned)D.c2;
  y = (int)(idx / (unsigned)D.c2);
  x = (int)(idx - (unsigned)y * (unsigned)D.c2);
  ci = (int64_t)zr * plane + idx;
  return idx < plane;
}
__device__ __forceinline__ int64_t block_linear() { return (int64_t)blockIdx.y * gridDim.x + blockIdx.x; }

// corner i of the cell at (z, y, x): Lewiner's numbering (0,0,0) (1,0,0) (1,1,0) (0,1,0) and the same at z+1 --
// the order of DJ_MC_CORNER, written out so that the eight loads are one base address plus constant offsets
__device__ __forceinline__ void load_cube(const Dims& D, const float* __restrict__ vol, int zr, int y, int x, double (&f)[8]) {
  const int64_t sy = D.n2, sz = (int64_t)D.n1 * D.n2;
  const float* p = vol + ((int64_t)zr * D.n1 + y) * D.n2 + x;
  const float v[8] = {__ldg(p), __ldg(p + 1), __ldg(p + sy + 1), __ldg(p + sy),
                      __ldg(p + sz), __ldg(p + sz + 1), __ldg(p + sz + sy + 1), __ldg(p + sz + sy)};
#pragma unroll
  for (int i = 0; i < 8; i++) f[i] = __dsub_rn((double)v[i], D.level);
}

__device__ __forceinline__ int cube_variant(const double (&f)[8]) {
  int cfg = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) cfg |= (f[i] > 0.0) ? (1 << i) : 0;
  int var = DJ_MC_VARBASE[cfg];
  const int na = DJ_MC_NAMB[cfg];
  for (int a = 0; a < na; a++) {
    const unsigned char* c = DJ_MC_FACE[DJ_MC_AMBFACE[cfg][a]];
    // Lewiner's test_face: with A, C one diagonal and B, D the other, the corners above the level are joined through
    // the face when |AC - BD| < FLT_EPSILON (his dead band) or the face-centre saddle (AC - BD) / (A + C - B - D)
    // is positive; products and difference rounded separately (no fused multiply-add), as the C library computes them
    const double p02 = __dmul_rn(f[c[0]], f[c[2]]);
    const double p13 = __dmul_rn(f[c[1]], f[c[3]]);
    const double det = __dsub_rn(p02, p13);
    const bool joined = (fabs(det) < FLT_EPS_D) || (((cfg >> c[0]) & 1) ? (det > 0.0) : (det < 0.0));
    var += joined ? (1 << a) : 0;
  }
  return var;
}

// bit 12 = the cell-centre vertex, always created by its own cell
__device__ __forceinline__ unsigned owned_mask(const Dims& D, int zr, int y, int x) {
  return 0x1FFFu & ~(((zr | D.seam) ? MZ : 0u) | (y ? MY : 0u) | (x ? MX : 0u));
}

__device__ __forceinline__ int owned_count(int var, unsigned own) {
  int n = 0;
  const int ne = DJ_MC_NEDGE[var];
  for (int k = 0; k < ne; k++) n += (own >> DJ_MC_ORDER[var][k]) & 1u;
  return n;
}

// block-wide exclusive scan of one int per thread (CB threads); returns exclusive prefix, total in *tot
__device__ __forceinline__ int block_excl_scan(int v, int* smem /*[CB/32]*/, int* tot) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = (lane < CB / 32) ? smem[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane < CB / 32) smem[lane] = s;
  }
  __syncthreads();
  const int base = w ? smem[w - 1] : 0;
  *tot = smem[CB / 32 - 1];
  __syncthreads();
  return base + inc - v;
}

// ---- pass 0 (optional): min / max of the volume for skimage's level-range ValueError
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

// ---- pass 1: classify every cell; per-block (vertex, triangle) counts (block_counts zeroed by the caller).
// ~97 % of the cells are not cut by the surface: their path is 8 loads, 8 compares and one 2-byte store -- the
// float64 differences, the table look-ups and the count atomics are only touched by cut cells.
__global__ void __launch_bounds__(CB) classify_kernel(Dims D, const float* __restrict__ vol, uint16_t* __restrict__ var_out,
                                                      int2* __restrict__ block_counts) {
  int zr, y, x;
  int64_t ci;
  int nv = 0, nt = 0;
  if (block_cell(D, threadIdx.x, zr, y, x, ci)) {
    const int64_t sy = D.n2, sz = (int64_t)D.n1 * D.n2;
    const float* p = vol + ((int64_t)zr * D.n1 + y) * D.n2 + x;
    const float v[8] = {__ldg(p), __ldg(p + 1), __ldg(p + sy + 1), __ldg(p + sy),
                        __ldg(p + sz), __ldg(p + sz + 1), __ldg(p + sz + sy + 1), __ldg(p + sz + sy)};
    int cfg = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) cfg |= ((double)v[i] > D.level) ? (1 << i) : 0;  // == ((double)v - level > 0), exactly
    if (cfg == 0 || cfg == 255) {
      var_out[ci] = (uint16_t)(cfg == 0 ? DJ_MC_VAR_EMPTY0 : DJ_MC_VAR_EMPTY255);
    } else {
      double f[8];
#pragma unroll
      for (int i = 0; i < 8; i++) f[i] = __dsub_rn((double)v[i], D.level);
      const int var = cube_variant(f);
      var_out[ci] = (uint16_t)var;
      nt = DJ_MC_NTRI[var];
      if (nt) nv = owned_count(var, owned_mask(D, zr, y, x));
    }
  }
  if (__any_sync(0xffffffffu, nt != 0)) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      nv += __shfl_xor_sync(0xffffffffu, nv, o);
      nt += __shfl_xor_sync(0xffffffffu, nt, o);
    }
    if ((threadIdx.x & 31) == 0) {
      int2* c = block_counts + block_linear();
      if (nv) atomicAdd(&c->x, nv);
      atomicAdd(&c->y, nt);
    }
  }
}

// ---- pass 2: exclusive scan of the per-block counts in two levels.
// level 1: every CTA scans 1024 counts in place (-> offsets within its group) and writes the group total;
// level 2: one CTA scans the group totals and writes the grand totals for the host.  Consumers add
// offsets[b] + group_off[b >> 10].
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
__global__ void __launch_bounds__(SG) scan_groups_kernel(const int2* __restrict__ counts, int64_t nblocks, int2* __restrict__ offsets,
                                                         longlong2* __restrict__ group_tot) {
  __shared__ long long wsum[2][32];
  const int64_t i = (int64_t)blockIdx.x * SG + threadIdx.x;
  const int2 c = (i < nblocks) ? counts[i] : make_int2(0, 0);
  long long a = c.x, b = c.y;
  scan1024(a, b, wsum);
  if (i < nblocks) offsets[i] = make_int2((int)a, (int)b);
  if (threadIdx.x == SG - 1) group_tot[blockIdx.x] = make_longlong2(a + c.x, b + c.y);
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

// ---- pass 3: vertices (sweep first-use order) + edge -> id table.  One thread per owned vertex.
constexpr int MAXV = 13;  // owned vertices of one cell: 12 edges + the centre
__global__ void __launch_bounds__(CB, 4) emit_vertices_kernel(Dims D, const float* __restrict__ vol,
                                                              const uint16_t* __restrict__ var_in,
                                                              const int2* __restrict__ block_counts,
                                                              const int2* __restrict__ block_off,
                                                              const longlong2* __restrict__ group_off, int* __restrict__ edge_id,
                                                              float* __restrict__ verts) {
  __shared__ int sscan[CB / 32];
  __shared__ uint16_t slist[CB * MAXV];  // (cell's thread << 4) | edge, in id order
  if (block_counts[block_linear()].x == 0) return;  // most blocks: the surface creates no vertex here
  int zr, y, x;
  int64_t ci;
  int var = 0, nv = 0;
  unsigned own = 0;
  if (block_cell(D, threadIdx.x, zr, y, x, ci)) {
    var = var_in[ci];
    if (DJ_MC_NTRI[var]) {
      own = owned_mask(D, zr, y, x);
      nv = owned_count(var, own);
    }
  }
  int tot;
  const int pre = block_excl_scan(nv, sscan, &tot);
  if (tot == 0) return;
  if (nv) {
    const int ne = DJ_MC_NEDGE[var];
    int r = pre;
    for (int k = 0; k < ne; k++) {
      const int e = DJ_MC_ORDER[var][k];
      if ((own >> e) & 1u) slist[r++] = (uint16_t)((threadIdx.x << 4) | e);
    }
  }
  __syncthreads();
  const int64_t bl = block_linear();
  const int64_t id0 = (int64_t)block_off[bl].x + group_off[bl >> 10].x;
  for (int v = threadIdx.x; v < tot; v += CB) {
    const int ent = slist[v];
    const int e = ent & 15;
    int cz, cy, cx;
    int64_t cci;
    block_cell(D, ent >> 4, cz, cy, cx, cci);
    double f[8];
    load_cube(D, vol, cz, cy, cx, f);
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
    const int base[3] = {cx, cy, D.z_off + cz};
    float p[3];
#pragma unroll
    for (int d = 0; d < 3; d++) p[d] = (float)__dadd_rn((double)base[d], fr[d]);
    const int64_t id = id0 + v;
    verts[3 * id + 0] = p[2];  // (axis0, axis1, axis2)
    verts[3 * id + 1] = p[1];
    verts[3 * id + 2] = p[0];
    edge_id[edge_slot(D, cz, cy, cx, e)] = (int)id;
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

// ---- pass 4: faces, appended cell by cell in table order.  One thread per triangle.
constexpr int MAXT = 12;
static_assert(MAXT >= DJ_MC_MAXTRI, "triangle list capacity");
__global__ void __launch_bounds__(CB, 4) emit_faces_kernel(Dims D, const uint16_t* __restrict__ var_in,
                                                           const int2* __restrict__ block_counts,
                                                           const int2* __restrict__ block_off,
                                                           const longlong2* __restrict__ group_off,
                                                           const int* __restrict__ edge_id, int descent, int* __restrict__ faces) {
  __shared__ int sscan[CB / 32];
  __shared__ uint16_t slist[CB * MAXT];  // (cell's thread << 4) | triangle index, in face order
  if (block_counts[block_linear()].y == 0) return;  // most blocks: no triangle
  int zr, y, x;
  int64_t ci;
  int var = 0, nt = 0;
  if (block_cell(D, threadIdx.x, zr, y, x, ci)) {
    var = var_in[ci];
    nt = DJ_MC_NTRI[var];
  }
  int tot;
  const int pre = block_excl_scan(nt, sscan, &tot);
  if (tot == 0) return;
  for (int t = 0; t < nt; t++) slist[pre + t] = (uint16_t)((threadIdx.x << 4) | t);
  __syncthreads();
  const int64_t bl = block_linear();
  const int64_t f0 = (int64_t)block_off[bl].y + group_off[bl >> 10].y;
  for (int v = threadIdx.x; v < tot; v += CB) {
    const int ent = slist[v];
    const int t = ent & 15;
    int cz, cy, cx;
    int64_t cci;
    block_cell(D, ent >> 4, cz, cy, cx, cci);
    const int cvar = var_in[cci];
    int id[3];
#pragma unroll
    for (int k = 0; k < 3; k++) id[k] = edge_id[edge_slot(D, cz, cy, cx, DJ_MC_TRI[cvar][3 * t + k])];
    int* o = faces + 3 * (f0 + v);
    if (descent) { o[0] = id[2]; o[1] = id[1]; o[2] = id[0]; }
    else { o[0] = id[0]; o[1] = id[1]; o[2] = id[2]; }
  }
}

}  // namespace mc
}  // namespace dj
