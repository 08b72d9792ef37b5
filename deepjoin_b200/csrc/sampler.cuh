// On-device mini-batch schedule with the semantics of core.data.select_samples
// (core/data.py:15-73) + select_uniform_ratio_samples (:122-176), for all iterations at once.
//
// The reference, per iteration: keeps the surface half of the pool, a random k-subset of the
// uniform half (k = int(half*ratio/(1-ratio))), permutes the m = half+k rows, then takes a
// contiguous window of num/2 positive rows and one of num-num/2 non-positive rows at random
// starts of the (permuted-order) index lists.  Here the two random permutations are keyed
// bijections (cycle-walking Feistel networks over the next power of two) instead of numpy's
// Mersenne-Twister shuffles: same distribution family, different stream -- parity tests inject
// the reference's own schedule instead (dj_latent_optimize takes any schedule).
#pragma once
#include "common.cuh"

namespace dj {
namespace sampler {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// bijection on [0, n): 4-round Feistel on `bits` bits, walk until the value lands inside
__device__ __forceinline__ uint32_t feistel_perm(uint32_t i, uint32_t n, int bits, uint32_t k0, uint32_t k1) {
  const int hb = bits >> 1, lb = bits - hb;
  const uint32_t hmask = (1u << hb) - 1u, lmask = (1u << lb) - 1u;
  uint32_t v = i;
  do {
    uint32_t l = v & lmask, h = (v >> lb) & hmask;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      // unbalanced Feistel: alternate which half is updated
      if (r & 1) l = (l ^ mix32(h * 0x9E3779B9u + k0 + r * 0x85EBCA6Bu + k1)) & lmask;
      else h = (h ^ mix32(l * 0xC2B2AE35u + k1 + r * 0x27D4EB2Fu + k0)) & hmask;
    }
    v = (h << lb) | l;
  } while (v >= n);
  return v;
}

struct Sched {
  const float* sdf;
  int64_t pool_n;
  int half, k, m, bits_m, bits_h;
  int num, pos_num, neg_num;
  uint64_t seed;
  int32_t* idx;
  int* status;  // per iteration: 0 ok, 1 empty class, 2 class smaller than the request
};

__device__ __forceinline__ int row_of(const Sched& S, uint32_t j, uint32_t ka, uint32_t kb) {
  const uint32_t q = feistel_perm(j, (uint32_t)S.m, S.bits_m, ka, kb);
  if ((int)q < S.half) return (int)q;
  return S.half + (int)feistel_perm(q - S.half, (uint32_t)S.half, S.bits_h, kb ^ 0x5bd1e995u, ka + 0x1234567u);
}

// one CTA per iteration
__global__ void __launch_bounds__(1024) schedule_kernel(Sched S) {
  __shared__ int wsum[32];
  __shared__ int s_tot[2];
  __shared__ int s_run;
  const int e = blockIdx.x;
  const uint32_t ka = mix32((uint32_t)S.seed ^ (uint32_t)(e * 2654435761u)), kb = mix32((uint32_t)(S.seed >> 32) + 0x9E3779B9u * (uint32_t)(e + 1));
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // pass 1: number of positive rows in the selection
  int cnt = 0;
  for (int j = threadIdx.x; j < S.m; j += blockDim.x) cnt += (S.sdf[row_of(S, j, ka, kb)] > 0.f) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) wsum[w] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int P = 0;
    for (int i = 0; i < 32; i++) P += wsum[i];
    s_tot[0] = P;
    s_tot[1] = S.m - P;
    int st = 0;
    if (P == 0 || S.m - P == 0) st = 1;
    else if (P <= S.pos_num || S.m - P <= S.neg_num) st = 2;
    S.status[e] = st;
  }
  __syncthreads();
  const int P = s_tot[0], Q = s_tot[1];
  if (P <= S.pos_num || Q <= S.neg_num) return;
  // np.random.randint(len - num): start in [0, len - num)
  const int start_p = (int)(mix32(ka ^ 0xA5A5A5A5u) % (uint32_t)(P - S.pos_num));
  const int start_q = (int)(mix32(kb ^ 0x3C3C3C3Cu) % (uint32_t)(Q - S.neg_num));
  int32_t* out = S.idx + (size_t)e * S.num;
  // pass 2: ordinal of every row within its class (block scan over the permuted order), emit windows
  int run_p = 0;  // positives seen before this chunk (identical in all threads)
  for (int base = 0; base < S.m; base += blockDim.x) {
    const int j = base + threadIdx.x;
    int row = -1, pos = 0;
    if (j < S.m) { row = row_of(S, j, ka, kb); pos = (S.sdf[row] > 0.f) ? 1 : 0; }
    int inc = pos;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
      int s = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
      wsum[lane] = s;
      if (lane == 31) s_run = s;
    }
    __syncthreads();
    if (j < S.m) {
      const int excl = (w ? wsum[w - 1] : 0) + inc - pos;
      const int op = run_p + excl;            // ordinal among positives
      const int oq = j - op;                  // ordinal among non-positives
      if (pos) { if (op >= start_p && op < start_p + S.pos_num) out[op - start_p] = row; }
      else { if (oq >= start_q && oq < start_q + S.neg_num) out[S.pos_num + oq - start_q] = row; }
    }
    run_p += s_run;
    __syncthreads();
  }
}

inline int sample_schedule(const float* pool_sdf_dev, int64_t pool_n, int num_iterations, int num_samples, float ratio,
                           uint64_t seed, int32_t* idx_dev, cudaStream_t st) {
  if (!pool_sdf_dev || !idx_dev || pool_n < 2 || num_iterations < 1 || num_samples < 2)
    return fail(DJ_ERR_INVALID, "dj_sample_schedule: bad argument");
  if (!(ratio > 0.f) || ratio > 0.5f) return fail(DJ_ERR_UNSUPPORTED, "uniform_ratio %g: only (0, 0.5] is implemented", ratio);
  if (pool_n >= (1LL << 30)) return fail(DJ_ERR_UNSUPPORTED, "pool too large");
  Sched S;
  S.sdf = pool_sdf_dev;
  S.pool_n = pool_n;
  S.half = (int)(pool_n / 2);
  S.k = (ratio == 0.5f) ? S.half : (int)(((double)S.half * (double)ratio) / (1.0 - (double)ratio));
  S.m = S.half + S.k;
  auto bits = [](int n) { int b = 2; while ((1LL << b) < n) b++; return b; };
  S.bits_m = bits(S.m);
  S.bits_h = bits(S.half);
  S.num = num_samples;
  S.pos_num = num_samples / 2;
  S.neg_num = num_samples - S.pos_num;
  S.seed = seed;
  S.idx = idx_dev;
  int* status = nullptr;
  DJ_CUDA(cudaMalloc((void**)&status, num_iterations * sizeof(int)));
  S.status = status;
  schedule_kernel<<<num_iterations, 1024, 0, st>>>(S);
  launch_counter().fetch_add(1);
  std::vector<int> h(num_iterations);
  cudaError_t e = cudaMemcpyAsync(h.data(), status, num_iterations * sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(status);
  if (e != cudaSuccess) return fail(DJ_ERR_CUDA, "dj_sample_schedule: %s", cudaGetErrorString(e));
  for (int v : h) {
    if (v == 1) return fail(DJ_ERR_NO_SAMPLES, "a class (sdf > 0 / sdf <= 0) has no rows");
    if (v == 2) return fail(DJ_ERR_UNSUPPORTED, "a class has no more rows than requested (sampling with replacement is host-only)");
  }
  return DJ_OK;
}

}  // namespace sampler
}  // namespace dj
