// Diagnostics only (not part of the product library): what does tcgen05.mma.cta_group::2 with M = 128 (64 rows per
// CTA) do on sm_100a?  (1) where the accumulator rows / columns land in tensor memory, (2) where an A operand read
// from tensor memory is expected, (3) how many cycles one such MMA takes next to the M = 256 one the decoder
// kernels use.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_m128 probe_m128.cu ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include "../../deepjoin_b200/csrc/ptx.cuh"

using namespace dj;

struct Args {
  const __half* A;   // [2][64][64]   rows of CTA r (only the first m_rows are used when M = 256: [2][128][64])
  const __half* B;   // [2][N/2][64]
  const uint32_t* Atm;  // [2][128 lanes][32 cols] words stored at TMEM columns [256, 288) (A operand from TMEM)
  float* D;          // [2][128 lanes][256 cols] dump of TMEM columns [0, 256)
  long long* cycles; // [1]
  int M, N, ksteps, reps, a_tmem, a_col_step, a_lane_base, d_alt;
  unsigned int* err;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ Args P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t sA = base, sB = base + 16384, bar_d = base + 49152;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem + 49152 + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int rows = P.M / 2, brows = P.N / 2;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_d, 1);
    ptx::mbar_fence_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc2(ptx::smem_u32((const void*)slot), 512);
    ptx::tmem_relinquish2();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
  {  // sentinel in columns [0, 256), the A words in [256, 288)
    uint32_t s[32];
#pragma unroll
    for (int i = 0; i < 32; i++) s[i] = __float_as_uint(-7.0f);
    for (int c = 0; c < 256; c += 32) ptx::tmem_st32(tl + c, s);
#pragma unroll
    for (int i = 0; i < 32; i++) s[i] = P.Atm[((size_t)rank * 128 + warp * 32 + lane) * 32 + i];
    ptx::tmem_st32(tl + 256, s);
    ptx::tmem_wait_st();
  }
  // operands into shared memory, K-major rows of 128 B, 128-byte swizzle
  for (int i = threadIdx.x; i < rows * 8; i += 128) {
    const int r = i >> 3, g = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(P.A + ((size_t)rank * rows + r) * 64 + g * 8);
    *reinterpret_cast<uint4*>(smem + r * 128 + ((g ^ (r & 7)) << 4)) = v;
  }
  for (int i = threadIdx.x; i < brows * 8; i += 128) {
    const int r = i >> 3, g = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(P.B + ((size_t)rank * brows + r) * 64 + g * 8);
    *reinterpret_cast<uint4*>(smem + 16384 + r * 128 + ((g ^ (r & 7)) << 4)) = v;
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = ptx::umma_idesc_f16((uint32_t)P.M, (uint32_t)P.N);
    const uint64_t ad = ptx::umma_desc_sw128_kmajor(sA), bd = ptx::umma_desc_sw128_kmajor(sB);
    const long long t0 = clock64();
    if (P.ksteps == 4 && P.reps > 1) {  // timing: 8 MMAs per trip, constant offsets, as the decoder kernel issues them
      const int a_tmem = P.a_tmem, reps = P.reps / 2;
      const uint32_t at = tmem + 256u + (uint32_t)P.a_lane_base;
      for (int r = 0; r < reps; r++) {
        if (a_tmem) {
#pragma unroll
          for (int k = 0; k < 8; k++) ptx::mma2_f16_ts(tmem + (uint32_t)((k >> 2) * P.d_alt), at + (uint32_t)((k & 3) * 8), bd + (uint64_t)(2 * (k & 3)), idesc, 1u);
        } else {
#pragma unroll
          for (int k = 0; k < 8; k++) ptx::mma2_f16_ss(tmem + (uint32_t)((k >> 2) * P.d_alt), ad + (uint64_t)(2 * (k & 3)), bd + (uint64_t)(2 * (k & 3)), idesc, 1u);
        }
      }
    } else
    for (int r = 0; r < P.reps; r++)
      for (int k = 0; k < P.ksteps; k++) {
        if (P.a_tmem) ptx::mma2_f16_ts(tmem, tmem + 256u + (uint32_t)P.a_lane_base + (uint32_t)(k * P.a_col_step), bd + (uint64_t)(2 * k), idesc, (uint32_t)((r | k) != 0));
        else ptx::mma2_f16_ss(tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (uint32_t)((r | k) != 0));
      }
    ptx::mma2_commit(bar_d, 3);
    ptx::mbar_wait(bar_d, 0, P.err, 0x801);
    P.cycles[0] = clock64() - t0;
  }
  ptx::mbar_wait(bar_d, 0, P.err, 0x802);
  ptx::tc_fence_after();
  for (int c0 = 0; c0 < 256; c0 += 32) {
    uint32_t r[32];
    ptx::tmem_ld32(tl + c0, r);
    ptx::tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; i++) P.D[((size_t)rank * 128 + warp * 32 + lane) * 256 + c0 + i] = __uint_as_float(r[i]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 0) ptx::tmem_dealloc2(tmem, 512);
}


// (5) can a 1-D bulk copy into THIS CTA's shared memory signal the mbarrier of the OTHER CTA of the pair?  Both CTAs copy
// `bytes` into their own shared memory; mode 0: each signals its own barrier and CTA 1 relays with a remote arrive (what
// the decoder kernels do), mode 1: both signal CTA 0's barrier directly.  CTA 0 measures issue -> completion.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 1) relay_kernel(const uint8_t* src, uint32_t bytes, int mode, int reps,
                                                                                 long long* cycles, uint32_t* sums, unsigned int* err) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t bar = base + 32768, go = bar + 8;
  const uint32_t rank = ptx::cluster_ctarank();
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, mode == 0 ? (rank == 0 ? 2 : 1) : 1);
    ptx::mbar_init(go, 1);
    ptx::mbar_fence_init();
  }
  __syncthreads();
  ptx::cluster_sync();
  const uint32_t lead_bar = ptx::mapa(bar, 0), peer_go = ptx::mapa(go, 1);
  long long total = 0;
  for (int r = 0; r < reps; r++) {
    const uint32_t ph = (uint32_t)(r & 1);
    if (rank == 0 && threadIdx.x == 0) {
      const long long t0 = clock64();
      ptx::mbar_arrive_cluster(peer_go);  // tell the peer to start its copy now
      if (mode == 0) ptx::mbar_expect_tx(bar, bytes);
      else ptx::mbar_expect_tx(bar, 2 * bytes);
      ptx::bulk_g2s(base, src, bytes, mode == 0 ? bar : lead_bar);
      ptx::mbar_wait(bar, ph, err, 0x901);
      total += clock64() - t0;
    }
    if (rank == 1 && threadIdx.x == 0) {
      ptx::mbar_wait(go, ph, err, 0x902);
      if (mode == 0) {
        ptx::mbar_expect_tx(bar, bytes);
        ptx::bulk_g2s(base, src + bytes, bytes, bar);
      } else {
        ptx::bulk_g2s(base, src + bytes, bytes, lead_bar);
      }
    }
    if (rank == 1 && threadIdx.x == 32 && mode == 0) {  // relay warp
      ptx::mbar_spin(bar, ph);
      ptx::mbar_arrive_cluster(lead_bar);
    }
    __syncthreads();
    ptx::cluster_sync();
  }
  // checksum of what landed
  uint32_t s = 0;
  for (uint32_t i = threadIdx.x; i < bytes / 4; i += 64) s += reinterpret_cast<const uint32_t*>(smem)[i];
  atomicAdd(&sums[rank], s);
  if (rank == 0 && threadIdx.x == 0) cycles[0] = total;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Run {
  std::vector<__half> A, B;
  std::vector<uint32_t> Atm;
  std::vector<float> D;
  long long cycles = 0;
};

static bool launch(Run& R, int M, int N, int ksteps, int reps, int a_tmem, int a_col_step, int a_lane_base = 0, int d_alt = 0) {
  __half *dA, *dB; uint32_t* dAt; float* dD; long long* dC; unsigned int* dE;
  R.A.resize(2 * 128 * 64); R.B.resize(2 * 128 * 64); R.Atm.resize(2 * 128 * 32); R.D.assign(2 * 128 * 256, 0.f);
  CK(cudaMalloc(&dA, R.A.size() * 2)); CK(cudaMalloc(&dB, R.B.size() * 2)); CK(cudaMalloc(&dAt, R.Atm.size() * 4));
  CK(cudaMalloc(&dD, R.D.size() * 4)); CK(cudaMalloc(&dC, 8)); CK(cudaMalloc(&dE, 4));
  CK(cudaMemset(dE, 0, 4)); CK(cudaMemset(dC, 0, 8));
  CK(cudaMemcpy(dA, R.A.data(), R.A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, R.B.data(), R.B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dAt, R.Atm.data(), R.Atm.size() * 4, cudaMemcpyHostToDevice));
  Args P{dA, dB, dAt, dD, dC, M, N, ksteps, reps, a_tmem, a_col_step, a_lane_base, d_alt, dE};
  const int smem = 49152 + 64 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<<<2, 128, smem>>>(P);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned int h = 0;
  bool ok = e == cudaSuccess;
  if (ok) { CK(cudaMemcpy(&h, dE, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(R.D.data(), dD, R.D.size() * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(&R.cycles, dC, 8, cudaMemcpyDeviceToHost)); }
  else printf("  launch failed: %s\n", cudaGetErrorString(e));
  if (h) { printf("  watchdog 0x%08x\n", h); ok = false; }
  cudaFree(dA); cudaFree(dB); cudaFree(dAt); cudaFree(dD); cudaFree(dC); cudaFree(dE);
  return ok;
}

// prints, for CTA c, the map (lane range, column range) -> (row range, column range of D) in compressed form
static void describe(const std::vector<int>& row_of, const std::vector<int>& col_of, int ncols, const char* what) {
  printf("%s\n", what);
  for (int c = 0; c < 2; c++)
    for (int l0 = 0; l0 < 128; l0 += 16) {
      printf("  cta %d lanes %3d-%3d:", c, l0, l0 + 15);
      // runs of columns with the same affine pattern
      int j = 0;
      while (j < ncols) {
        const int idx = (c * 128 + l0) * 256 + j;
        if (row_of[idx] < 0) { int j1 = j; while (j1 < ncols && row_of[(c * 128 + l0) * 256 + j1] < 0) j1++; printf(" cols %d-%d untouched;", j, j1 - 1); j = j1; continue; }
        int j1 = j + 1;
        while (j1 < ncols && row_of[(c * 128 + l0) * 256 + j1] == row_of[idx] && col_of[(c * 128 + l0) * 256 + j1] == col_of[idx] + (j1 - j)) j1++;
        // check the 16 lanes are consecutive rows
        bool consec = true;
        for (int l = 0; l < 16; l++) consec &= row_of[(c * 128 + l0 + l) * 256 + j] == row_of[idx] + l && col_of[(c * 128 + l0 + l) * 256 + j] == col_of[idx];
        printf(" cols %d-%d = D rows %d%s, D cols %d-%d;", j, j1 - 1, row_of[idx], consec ? "..+15" : " (lanes NOT consecutive rows)", col_of[idx], col_of[idx] + (j1 - j) - 1);
        j = j1;
      }
      printf("\n");
    }
}

int main(int argc, char** argv) {
  const int section = argc > 1 ? atoi(argv[1]) : 0;  // 1 layout, 2 A from TMEM, 3 timing SS, 4 timing TS
  // ---- (1) accumulator layout, A and B from shared memory: D[m][n] = (m + 1) + 256 (n + 1)
  for (int cfg = 0; cfg < 3 && section == 1; cfg++) {
    const int M = cfg == 0 ? 256 : 128, N = cfg == 2 ? 256 : 128;
    Run R; R.A.assign(2 * 128 * 64, __float2half(0.f)); R.B.assign(2 * 128 * 64, __float2half(0.f)); R.Atm.assign(2 * 128 * 32, 0);
    for (int m = 0; m < M; m++) { R.A[(size_t)m * 64 + 0] = __float2half((float)(m + 1)); R.A[(size_t)m * 64 + 1] = __float2half(1.f); }
    for (int n = 0; n < N; n++) { R.B[(size_t)n * 64 + 0] = __float2half(1.f); R.B[(size_t)n * 64 + 1] = __float2half((float)(128 * (n + 1))); }
    printf("=== accumulator layout, cta_group::2 M=%d N=%d (SS) ===\n", M, N);
    if (!launch(R, M, N, 1, 1, 0, 0)) continue;
    std::vector<int> row_of(R.D.size(), -1), col_of(R.D.size(), -1);
    for (size_t i = 0; i < R.D.size(); i++) {
      const float v = R.D[i];
      if (v == -7.0f) continue;
      const int iv = (int)v;
      row_of[i] = iv % 128 == 0 ? 127 : iv % 128 - 1;  // (m+1) + 128 (n+1), m < 128 only decodable mod 128
      col_of[i] = (iv - (row_of[i] + 1)) / 128 - 1;
    }
    describe(row_of, col_of, 256, "  (rows are reported modulo 128: CTA 1 holds the upper half)");
  }
  // ---- (2) A operand from tensor memory, M = 128: identity B picks A[m][k] into D[m][k]
  for (int pass = 0; pass < 2 && section == 2; pass++) {
    Run R; R.A.assign(2 * 128 * 64, __float2half(0.f)); R.B.assign(2 * 128 * 64, __float2half(0.f)); R.Atm.assign(2 * 128 * 32, 0);
    for (int n = 0; n < 16; n++) R.B[(size_t)n * 64 + n] = __float2half(1.f);
    for (int c = 0; c < 2; c++)
      for (int l = 0; l < 128; l++)
        for (int w = 0; w < 32; w++) {
          const float v0 = pass == 0 ? (float)(l + 1) : (float)(2 * w + 1), v1 = pass == 0 ? (float)(l + 1) : (float)(2 * w + 2);
          const __half h0 = __float2half(v0), h1 = __float2half(v1);
          uint16_t b0, b1; memcpy(&b0, &h0, 2); memcpy(&b1, &h1, 2);
          R.Atm[((size_t)c * 128 + l) * 32 + w] = (uint32_t)b0 | ((uint32_t)b1 << 16);
        }
    printf("=== A from tensor memory, cta_group::2 M=128 N=128 K=16: D[m][k] = %s of the TMEM word A[m][k] was read from ===\n",
           pass == 0 ? "lane + 1" : "2*column + half + 1");
    for (int lb = 0; lb < 2; lb++) {
    printf(" -- A address lane base %d\n", lb * 64);
    if (!launch(R, 128, 128, 1, 1, 1, 8, (lb * 64) << 16)) continue;
    for (int c = 0; c < 2; c++)
      for (int l = 0; l < 128; l += (pass == 0 ? 8 : 32)) {
        printf("  cta %d D-lane %3d, cols 0-15 and 64-79:", c, l);
        for (int j = 0; j < 16; j++) printf(" %g", R.D[((size_t)c * 128 + l) * 256 + j]);
        printf(" |");
        for (int j = 64; j < 80; j++) printf(" %g", R.D[((size_t)c * 128 + l) * 256 + j]);
        printf("\n");
      }
    }
  }
  // ---- (3) cycles per MMA
  for (int cfg = 0; cfg < 6 && section >= 3; cfg++) {
    const int M = (cfg % 3) == 0 ? 256 : 128, N = (cfg % 3) == 2 ? 256 : 128, a_tmem = cfg / 3;
    if (a_tmem != section - 3) continue;
    Run R; R.A.assign(2 * 128 * 64, __float2half(0.f)); R.B.assign(2 * 128 * 64, __float2half(0.f)); R.Atm.assign(2 * 128 * 32, 0);
    const int reps = 2048;
    for (int alt = 0; alt < 2; alt++) {
      if (!launch(R, M, N, 4, reps, a_tmem, 8, 0, alt * 128)) continue;
      printf("timing cta_group::2 M=%d N=%d K=16 %s%s: %lld cycles for %d MMAs = %.1f cycles/MMA, %.0f MAC/cycle/SM\n", M, N, a_tmem ? "TS" : "SS",
             alt ? ", two accumulators alternating every 4 MMAs" : "", R.cycles, reps * 4, (double)R.cycles / (reps * 4),
             (double)M * N * 16 / 2 / ((double)R.cycles / (reps * 4)));
    }
  }

  if (section == 5) {
    const uint32_t bytes = 16384;
    std::vector<uint32_t> h(2 * bytes / 4);
    uint32_t want[2] = {0, 0};
    for (size_t i = 0; i < h.size(); i++) { h[i] = (uint32_t)(i * 2654435761u + 12345u); want[i / (bytes / 4)] += h[i]; }
    uint8_t* d; long long* dC; uint32_t* dS; unsigned int* dE;
    CK(cudaMalloc(&d, 2 * bytes)); CK(cudaMalloc(&dC, 8)); CK(cudaMalloc(&dS, 8)); CK(cudaMalloc(&dE, 4));
    CK(cudaMemcpy(d, h.data(), 2 * bytes, cudaMemcpyHostToDevice));
    const int smem = 32768 + 64 + 1024;
    CK(cudaFuncSetAttribute(relay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int mode = 0; mode < 2; mode++) {
      const int reps = 200;
      CK(cudaMemset(dS, 0, 8)); CK(cudaMemset(dE, 0, 4)); CK(cudaMemset(dC, 0, 8));
      relay_kernel<<<2, 64, smem>>>(d, bytes, mode, reps, dC, dS, dE);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: launch failed: %s\n", mode, cudaGetErrorString(e)); break; }
      long long c; uint32_t sgot[2]; unsigned int eh;
      CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(sgot, dS, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&eh, dE, 4, cudaMemcpyDeviceToHost));
      printf("bulk copy completion %s: %.0f cycles from issue to the leader seeing both halves (16 KB each, L2 hits); data %s; watchdog 0x%x\n",
             mode == 0 ? "on own barriers + remote arrive relay" : "signalled on the leader's barrier by both CTAs", (double)c / reps,
             (sgot[0] == want[0] && sgot[1] == want[1]) ? "OK" : "WRONG", eh);
    }
  }
  return 0;
}
