// Debug-only UMMA probe: one 128 x 128 x K bf16 GEMM built from exactly the pieces the fused
// decoder kernel uses (swizzled A written by threads, pre-swizzled B tiles, smem descriptors,
// tcgen05.mma / commit / ld).  Lets tests/tools validate descriptor encodings in isolation and
// sweep alternatives (descriptor template, K advance) from Python.  Not part of the product path.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace dj {
namespace probe {

struct ProbeParams {
  const float* A;        // [128, K] fp32 row-major (converted to bf16 in-kernel)
  const uint8_t* Bblob;  // K/64 pre-swizzled [128 x 64] bf16 tiles (pack_tile layout)
  float* D;              // [128, 128] fp32 out
  int K;                 // multiple of 64, <= 256
  int use_bulk;          // bit 0: cp.async.bulk for B (else generic loads + proxy fence); bit 1: A operand in TMEM
  int k_adv_bytes;       // start-address advance per K=16 step (32)
  uint32_t idesc;
  uint64_t desc_tmpl;    // descriptor without the address field
  unsigned int* err;
};

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const __grid_constant__ ProbeParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - ptx::smem_u32(smem_raw));
  const int kc = P.K / 64;
  const uint32_t sA = base, sB = base + 4 * 16384;
  const uint32_t bar_w = base + 8 * 16384, bar_d = bar_w + 8;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem + 8 * 16384 + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_w, 1);
    ptx::mbar_init(bar_d, 1);
    ptx::mbar_fence_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(ptx::smem_u32((const void*)slot), 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  // A: thread = row, swizzled 16-byte chunks (same formula as the epilogue of mlp_tc.cuh)
  const int row = threadIdx.x;
  for (int c = 0; c < kc; c++)
    for (int gq = 0; gq < 8; gq++) {
      const float* a = P.A + (size_t)row * P.K + c * 64 + gq * 8;
      const uint32_t dst = sA + c * 16384 + row * 128 + ((gq ^ (row & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(ptx::pack_bf16x2(a[0], a[1])),
                   "r"(ptx::pack_bf16x2(a[2], a[3])), "r"(ptx::pack_bf16x2(a[4], a[5])), "r"(ptx::pack_bf16x2(a[6], a[7]))
                   : "memory");
    }
  const bool a_tmem = (P.use_bulk & 2) != 0;
  if (a_tmem) {
    // the epilogue's tensor-memory store of mlp_tc.cuh: bf16 pairs, TMEM lane = row, columns 128 + k/2
    for (int c = 0; c < kc; c++)
      for (int hh = 0; hh < 2; hh++) {
        uint32_t pk[16];
        const float* a = P.A + (size_t)row * P.K + c * 64 + hh * 32;
#pragma unroll
        for (int i = 0; i < 16; i++) pk[i] = ptx::pack_bf16x2(a[2 * i], a[2 * i + 1]);
        ptx::tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + 128u + (uint32_t)(c * 32 + hh * 16), pk);
      }
    ptx::tmem_wait_st();
    ptx::tc_fence_before();
  }
  if (P.use_bulk & 1) {
    if (threadIdx.x == 0) {
      ptx::mbar_expect_tx(bar_w, (uint32_t)(kc * 16384));
      for (int c = 0; c < kc; c++) ptx::bulk_g2s(sB + c * 16384, P.Bblob + (size_t)c * 16384, 16384, bar_w);
    }
  } else {
    const uint4* src = reinterpret_cast<const uint4*>(P.Bblob);
    uint4* dst = reinterpret_cast<uint4*>(smem + 4 * 16384);
    for (int i = threadIdx.x; i < kc * 1024; i += 128) dst[i] = src[i];
  }
  ptx::fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (P.use_bulk & 1) ptx::mbar_wait(bar_w, 0, P.err, 0x701);
    ptx::tc_fence_after();
    for (int c = 0; c < kc; c++)
      for (int k = 0; k < 4; k++) {
        const uint64_t ad = P.desc_tmpl | (uint64_t)(((sA + c * 16384 + k * P.k_adv_bytes) & 0x3FFFF) >> 4);
        const uint64_t bd = P.desc_tmpl | (uint64_t)(((sB + c * 16384 + k * P.k_adv_bytes) & 0x3FFFF) >> 4);
        if (a_tmem) ptx::mma_f16_ts(tmem, tmem + 128u + (uint32_t)(c * 32 + k * 8), bd, P.idesc, (uint32_t)((c | k) != 0));
        else ptx::mma_f16_ss(tmem, ad, bd, P.idesc, (uint32_t)((c | k) != 0));
      }
    ptx::mma_commit(bar_d);
  }
  ptx::mbar_wait(bar_d, 0, P.err, 0x702);
  ptx::tc_fence_after();
  const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t r[32];
    ptx::tmem_ld32(tl + c0, r);
    ptx::tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; i++) P.D[(size_t)(warp * 32 + lane) * 128 + c0 + i] = __uint_as_float(r[i]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 256);
}

}  // namespace probe
}  // namespace dj
