// Brute-force nearest neighbour on the device.
// replaces: scipy.spatial.cKDTree(targets).query(queries) as used by the evaluation metrics that follow
// meshing in the reference's pipeline (core/metrics.py:46-50 chamfer, :79 normal consistency) -- SURVEY.md 8(f4).
// 2,000 .. 30,000 points per side: N x M distance evaluations from shared-memory tiles beat building a tree.
#pragma once
#include "common.cuh"

namespace dj {
namespace nn {

constexpr int QB = 256;    // queries per block (one per thread)
constexpr int TT = 1024;   // targets per shared-memory tile

// dist2[i] = min_j |q_i - t_j|^2 (fp32, evaluated as the reference does: difference first), idx[i] = the smallest j
// attaining it
__global__ void __launch_bounds__(QB) nn_kernel(const float* __restrict__ q, const float* __restrict__ t, int64_t n, int64_t m,
                                                float* __restrict__ dist2, int32_t* __restrict__ idx) {
  __shared__ float sx[TT], sy[TT], sz[TT];
  const int64_t i = (int64_t)blockIdx.x * QB + threadIdx.x;
  float x = 0.f, y = 0.f, z = 0.f;
  if (i < n) { x = q[3 * i]; y = q[3 * i + 1]; z = q[3 * i + 2]; }
  float best = INFINITY;
  int bj = -1;
  for (int64_t base = 0; base < m; base += TT) {
    const int cnt = (int)min((int64_t)TT, m - base);
    __syncthreads();
    for (int k = threadIdx.x; k < cnt; k += QB) {
      sx[k] = t[3 * (base + k)];
      sy[k] = t[3 * (base + k) + 1];
      sz[k] = t[3 * (base + k) + 2];
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < cnt; k++) {
      const float dx = x - sx[k], dy = y - sy[k], dz = z - sz[k];
      const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      if (d < best) { best = d; bj = (int)(base + k); }
    }
  }
  if (i < n) { dist2[i] = best; idx[i] = bj; }
}

}  // namespace nn
}  // namespace dj
