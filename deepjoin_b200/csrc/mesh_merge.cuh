// Duplicate-vertex merge of the mesh create_mesh returns, on the device.
//
// replaces: trimesh.Trimesh(vertices, faces) with the default process=True at
// core/reconstruct.py:189 -- trimesh merges vertices that coincide after rounding to 8 decimals
// (tol.merge = 1e-8).  Marching cubes already shares one vertex per intersected edge, so duplicates
// only appear where the surface passes (numerically) through a grid sample; the common case is
// "nothing to merge" and costs two passes over the vertices.
//
// Semantics = deepjoin_b200/mesh.py:Mesh.merge_vertices (the host restatement the CPU tests pin):
// key(v) = rint(v * 1e8) per coordinate (np.round(v, 8)); vertices with equal keys collapse onto the
// one with the smallest index; survivors keep their relative order; faces are re-indexed.
#pragma once
#include "common.cuh"

namespace dj {
namespace meshmerge {

constexpr int32_t EMPTY = 0x7fffffff;

struct Key {
  long long a, b, c;
};
__device__ __forceinline__ Key key_of(const double* __restrict__ v, int64_t i) {
  Key k;
  k.a = __double2ll_rn(__dmul_rn(v[3 * i + 0], 1e8));
  k.b = __double2ll_rn(__dmul_rn(v[3 * i + 1], 1e8));
  k.c = __double2ll_rn(__dmul_rn(v[3 * i + 2], 1e8));
  return k;
}
__device__ __forceinline__ bool same(const Key& x, const Key& y) { return x.a == y.a && x.b == y.b && x.c == y.c; }
__device__ __forceinline__ uint32_t hash_of(const Key& k) {
  unsigned long long h = (unsigned long long)k.a * 0x9E3779B97F4A7C15ull;
  h ^= (unsigned long long)k.b * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
  h ^= (unsigned long long)k.c * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
  h ^= h >> 29;
  return (uint32_t)(h ^ (h >> 32));
}

__global__ void fill_kernel(int32_t* __restrict__ t, int64_t n, int32_t v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) t[i] = v;
}

// open addressing, linear probing; a slot is owned by one key and holds the smallest vertex index seen with it
__global__ void insert_kernel(const double* __restrict__ v, int64_t nv, int32_t* __restrict__ table, uint32_t mask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv) return;
  const Key k = key_of(v, i);
  uint32_t h = hash_of(k) & mask;
  for (;;) {
    int32_t cur = table[h];
    if (cur == EMPTY) {
      cur = atomicCAS(&table[h], EMPTY, (int32_t)i);
      if (cur == EMPTY) return;
    }
    if (same(key_of(v, cur), k)) {
      atomicMin(&table[h], (int32_t)i);
      return;
    }
    h = (h + 1) & mask;
  }
}

// canon[i] = smallest index with i's key; keep[i] = (canon[i] == i); *n_unique += kept
__global__ void lookup_kernel(const double* __restrict__ v, int64_t nv, const int32_t* __restrict__ table, uint32_t mask,
                              int32_t* __restrict__ canon, unsigned long long* __restrict__ n_unique) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int kept = 0;
  if (i < nv) {
    const Key k = key_of(v, i);
    uint32_t h = hash_of(k) & mask;
    for (;;) {
      const int32_t cur = table[h];
      if (same(key_of(v, cur), k)) {
        canon[i] = cur;
        kept = (cur == (int32_t)i);
        break;
      }
      h = (h + 1) & mask;
    }
  }
  const unsigned b = __ballot_sync(0xffffffffu, kept);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_unique, (unsigned long long)__popc(b));
}

// ---- compaction (only when something merged): newid = exclusive scan of keep
constexpr int SCAN_BLOCK = 1024;
__global__ void block_count_kernel(const int32_t* __restrict__ canon, int64_t nv, int32_t* __restrict__ block_sums) {
  __shared__ int s[32];
  const int64_t i = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  const int kept = (i < nv) && (canon[i] == (int32_t)i);
  const unsigned b = __ballot_sync(0xffffffffu, kept);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = __popc(b);
  __syncthreads();
  if (threadIdx.x < 32) {
    int x = s[threadIdx.x];
    for (int o = 16; o; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = x;
  }
}
// single block: exclusive scan of the block sums in place
__global__ void scan_sums_kernel(int32_t* __restrict__ block_sums, int nblocks) {
  __shared__ int s[SCAN_BLOCK];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += SCAN_BLOCK) {
    const int i = base + threadIdx.x;
    const int x = i < nblocks ? block_sums[i] : 0;
    s[threadIdx.x] = x;
    __syncthreads();
    for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
      const int y = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
      __syncthreads();
      s[threadIdx.x] += y;
      __syncthreads();
    }
    if (i < nblocks) block_sums[i] = carry + s[threadIdx.x] - x;
    __syncthreads();
    if (threadIdx.x == 0) carry += s[SCAN_BLOCK - 1];
    __syncthreads();
  }
}
__global__ void compact_kernel(const double* __restrict__ v, const int32_t* __restrict__ canon, int64_t nv,
                               const int32_t* __restrict__ block_offs, int32_t* __restrict__ newid, double* __restrict__ vout) {
  __shared__ int s[32];
  const int64_t i = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  const int kept = (i < nv) && (canon[i] == (int32_t)i);
  const unsigned b = __ballot_sync(0xffffffffu, kept);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) s[wid] = __popc(b);
  __syncthreads();
  if (wid == 0) {
    int x = s[lane], inc = x;
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += y;
    }
    s[lane] = inc - x;
  }
  __syncthreads();
  if (kept) {
    const int id = block_offs[blockIdx.x] + s[wid] + __popc(b & ((1u << lane) - 1u));
    newid[i] = id;
    vout[3 * (int64_t)id + 0] = v[3 * i + 0];
    vout[3 * (int64_t)id + 1] = v[3 * i + 1];
    vout[3 * (int64_t)id + 2] = v[3 * i + 2];
  }
}
__global__ void remap_faces_kernel(int32_t* __restrict__ faces, int64_t n, const int32_t* __restrict__ canon,
                                   const int32_t* __restrict__ newid) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) faces[i] = newid[canon[faces[i]]];
}

}  // namespace meshmerge
}  // namespace dj
