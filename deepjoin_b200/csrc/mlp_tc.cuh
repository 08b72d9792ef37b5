// Fused persistent decoder forward on tcgen05 tensor cores (sm_100a).
//
// replaces: Decoder.forward  (networks/decoder_z_lb_both_leaky_n_sdf.py:174-457) for one latent
// code shared by all rows, i.e. what decode_samples (core/reconstruct.py:115-129) feeds it.
//
// One CTA owns a tile of 128 query points and walks BOTH MLPs (fnet then hnet) without the
// activations ever leaving the SM:
//   * the latent columns of lin0 / lin4 / hnet_lin0 are folded into per-shape constants by
//     fold_kernel (mlp_simt.cuh); the remaining K=3 xyz products run on CUDA cores;
//   * every 512-wide layer is a 128 x N x K bf16 GEMM issued as 128 x 128 x 16 tcgen05.mma:
//     B = weight tiles streamed from L2 by the TMA engine (cp.async.bulk, 16 KB pre-swizzled stages,
//     mbarrier ring), D = one of two 128-column fp32 accumulator buffers in TMEM;
//   * the activations ping-pong between shared memory (K-major, 128-byte swizzle, A operand through
//     a shared-memory descriptor) and tensor memory (bf16 pairs in TMEM columns [0,256), A operand
//     read by the tensor core straight from TMEM): layer l reads one and its epilogue writes the
//     other, so the drain of accumulator chunk j (+ bias, LeakyReLU, -> bf16) runs while the tensor
//     core computes chunk j+1 of the same layer and the next layer starts on the K chunks that are
//     already published - the tensor pipe never waits for a whole-layer drain;
//   * heads (tanh / sigmoid / B,R combine) in the last epilogue; 4 B per point leave the SM.
//
// Warp roles (320 threads): warp 0 = weight producer, warp 1 = MMA issuer (one thread) + TMEM
// allocation, warps 2..9 = epilogue (TMEM lane quarter = warp % 4; two warps share a quarter and
// split every 64-column group 32/32).  CTAs of a cluster share one weight stream by multicast.
#pragma once
#include "common.cuh"
#include "ptx.cuh"
#include "heads.cuh"

namespace dj {
namespace tc {

constexpr int TILE_M = 128;
constexpr int KCH = 64;                       // k-chunk: 64 bf16 = one 128 B swizzle row
constexpr int NCH = 128;                      // accumulator chunk (TMEM columns)
constexpr int A_CHUNK_BYTES = TILE_M * KCH * 2;  // 16 KB
constexpr int STAGE_BYTES = NCH * KCH * 2;       // 16 KB
constexpr int NSTAGES = 5;
constexpr int MAX_LAYERS = 10;
constexpr int NEPI = 256;                     // epilogue threads (8 warps)
constexpr int NTHREADS = 64 + NEPI;

constexpr int SM_A = 0;
constexpr int SM_W = SM_A + 8 * A_CHUNK_BYTES;           // 131072
constexpr int SM_BIAS = SM_W + NSTAGES * STAGE_BYTES;    // + 81920
constexpr int SM_TAB = SM_BIAS + 2 * 512 * 4;            // + 4096
constexpr int SM_BAR = SM_TAB + 512 * 16;                // + 8192
constexpr int SM_END = SM_BAR + 256;
constexpr int SMEM_BYTES = SM_END + 1024;                // slack for 1024-B alignment

enum { KIND_HIDDEN = 0, KIND_HIDDEN_XYZ = 1, KIND_HEAD = 2 };

struct TcLayer {
  int16_t n_chunks;   // accumulator chunks of this layer
  int16_t k_chunks;   // 64-wide chunks of the input
  int16_t mma_n;      // N of each MMA (128 hidden, 16 head)
  int16_t kind;
  int32_t bias_off;   // float offset into TcParams::bias (KIND_HIDDEN / KIND_HEAD)
  int32_t table;      // xyzw table id (KIND_HIDDEN_XYZ)
  int32_t stage_bytes;
  int32_t pad;
};
struct TcNet {
  int32_t n_layers;
  int32_t l0_table;
  int32_t head_leaky;
  int32_t pad;
  const uint8_t* stream;
  TcLayer layers[MAX_LAYERS];
};
struct TcParams {
  TcNet nets[2];
  int32_t net_mask;
  int32_t grid_mode;
  int32_t debug;         // bench diagnostics only (DEEPJOIN_B200_DEBUG): 1 = do not re-stream weights, 2 = skip the MMAs
  const float* bias;
  const float4* tables;  // [3][512] (wx, wy, wz, c) per-shape constants
  const float* xyz;
  int64_t n;
  // grid mode: r = (i*d0 + j)*d2 + k -> (lin0[j], lin1[i], lin2[k])
  int32_t d0, d1, d2, pad0;
  double step0, step1, step2, stop, offs;
  int64_t r_lo;
  HeadCfg head;
  float* out;
  unsigned int* err;
};

__device__ __forceinline__ float lin_axis(int t, int d, double step, double stop, double offs) {
  // np.linspace(0, stop, d)[t] - offs in float64, no fused multiply-add, then f32
  double v = (t == d - 1) ? stop : __dmul_rn((double)t, step);
  return (float)__dsub_rn(v, offs);
}

__device__ __forceinline__ void grid_point(const TcParams& P, int64_t r, float& x, float& y, float& z) {
  int k = (int)(r % P.d2);
  int64_t q = r / P.d2;
  int j = (int)(q % P.d0);
  int i = (int)(q / P.d0);
  x = lin_axis(j, P.d0, P.step0, P.stop, P.offs);
  y = lin_axis(i, P.d1, P.step1, P.stop, P.offs);
  z = lin_axis(k, P.d2, P.step2, P.stop, P.offs);
}

__device__ __forceinline__ float leaky(float v, float slope) { return fmaxf(v, v * slope); }

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void st_shared_v4(uint32_t dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// TMEM columns: [0,256) bf16 activations (A operand of odd layers), [256,512) two fp32 accumulator buffers
constexpr uint32_t TM_A = 0;
constexpr uint32_t TM_ACC = 256;

// CLUSTER = number of CTAs that share one weight stream: every CTA fetches 1/CLUSTER of each stage and
// multicasts it to all (cp.async.bulk ... .multicast::cluster), cutting L2 reads per point by CLUSTER.
template <int CLUSTER>
__global__ void __launch_bounds__(NTHREADS, 1) tc_forward_kernel(const __grid_constant__ TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t sA = base + SM_A, sW = base + SM_W;
  float* sBias = reinterpret_cast<float*>(smem + SM_BIAS);
  float4* sTab = reinterpret_cast<float4*>(smem + SM_TAB);
  const uint32_t bar0 = base + SM_BAR;
  auto W_FULL = [&](int s) { return bar0 + 8u * s; };
  auto W_EMPTY = [&](int s) { return bar0 + 8u * (NSTAGES + s); };
  // A_READY(b, c): K chunk c of the activations in buffer b (0 = shared memory, 1 = tensor memory) is published
  auto A_READY = [&](int b, int c) { return bar0 + 8u * (2 * NSTAGES + 8 * b + c); };
  auto ACC_FULL = [&](int i) { return bar0 + 8u * (2 * NSTAGES + 16 + i); };
  auto ACC_EMPTY = [&](int i) { return bar0 + 8u * (2 * NSTAGES + 18 + i); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SM_BAR + 8 * (2 * NSTAGES + 20));
  static_assert(8 * (2 * NSTAGES + 20) + 4 <= 256, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = (P.n + TILE_M - 1) / TILE_M;
  // every CTA of a cluster must consume the same number of weight passes: uniform iteration count,
  // iterations past the last tile compute on zeros and write nothing
  const int64_t iters = (ntiles + gridDim.x - 1) / gridDim.x;
  constexpr uint16_t CMASK = (uint16_t)((1u << CLUSTER) - 1u);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NSTAGES; s++) {
      ptx::mbar_init(W_FULL(s), 1);
      ptx::mbar_init(W_EMPTY(s), CLUSTER);
    }
    for (int c = 0; c < 16; c++) ptx::mbar_init(A_READY(0, c), NEPI);
    for (int i = 0; i < 2; i++) {
      ptx::mbar_init(ACC_FULL(i), 1);
      ptx::mbar_init(ACC_EMPTY(i), NEPI);
    }
    ptx::mbar_fence_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32((const void*)tmem_slot), 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) ptx::cluster_sync();  // peers' barriers are initialised before anything is multicast to them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== weight producer (TMA engine, 1-D bulk copies) =====================
    // warp-uniform control flow; one elected lane arms the barrier and issues the copy
    const uint32_t rank = CLUSTER > 1 ? ptx::cluster_ctarank() : 0u;
    uint32_t stage = 0, phase = 0;
    for (int64_t it = 0; it < iters; it++) {
      for (int net = 0; net < 2; net++) {
        if (!((P.net_mask >> net) & 1)) continue;
        const TcNet& N = P.nets[net];
        const uint8_t* src = N.stream;
        for (int l = 0; l < N.n_layers; l++) {
          const int ns = N.layers[l].n_chunks * N.layers[l].k_chunks;
          const uint32_t bytes = (uint32_t)N.layers[l].stage_bytes;
          const uint32_t part = bytes / CLUSTER;
          for (int s = 0; s < ns; s++) {
            ptx::mbar_wait(W_EMPTY(stage), phase ^ 1, P.err, 0x100 | stage);  // every CTA of the cluster freed the slot
            if ((P.debug & 1) && (it > 0 || net > 0 || l > 0)) {
              if (ptx::elect_one()) ptx::mbar_arrive(W_FULL(stage));
            } else if (ptx::elect_one()) {
              ptx::mbar_expect_tx(W_FULL(stage), bytes);
              if (CLUSTER > 1)
                ptx::bulk_g2s_multicast(sW + stage * STAGE_BYTES + rank * part, src + rank * part, part, W_FULL(stage), CMASK);
              else
                ptx::bulk_g2s(sW + stage * STAGE_BYTES, src, bytes, W_FULL(stage));
            }
            __syncwarp();
            src += bytes;
            if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // GEMM layer l of a net reads its A operand from shared memory when l is even (layer 0 wrote it
    // there) and from tensor memory when l is odd; accumulator buffers alternate chunk by chunk.
    // The whole warp walks the (warp-uniform) issue program so that addresses and descriptors can live
    // in uniform registers; one elected lane issues tcgen05.mma / tcgen05.commit.  The per-stage path - a barrier
    // poll, four MMAs, one commit - must stay well under the 256 tensor-pipe cycles a stage is worth:
    // no out-of-line calls here.
    {
      const uint64_t desc_hi = ptx::umma_desc_sw128_kmajor(0);
      const uint64_t a_desc0 = desc_hi + (uint64_t)((sA & 0x3FFFF) >> 4);
      const uint64_t b_desc0 = desc_hi + (uint64_t)((sW & 0x3FFFF) >> 4);
      uint32_t stage = 0, phase = 0, a_ph = 0, acc_it = 0;
      for (int64_t it = 0; it < iters; it++) {
        for (int net = 0; net < 2; net++) {
          if (!((P.net_mask >> net) & 1)) continue;
          const TcNet& N = P.nets[net];
          const int nl = N.n_layers;
          for (int l = 0; l < nl; l++) {
            const int n_chunks = N.layers[l].n_chunks, k_chunks = N.layers[l].k_chunks;
            const uint32_t idesc = ptx::umma_idesc_bf16(128, (uint32_t)N.layers[l].mma_n);
            const int src = l & 1;
            for (int j = 0; j < n_chunks; j++) {
              const uint32_t buf = acc_it & 1u;
              ptx::mbar_spin(ACC_EMPTY(buf), ((acc_it >> 1) & 1u) ^ 1u);
              acc_it++;
              const uint32_t d_tmem = tmem_base + TM_ACC + buf * NCH;
              uint64_t ad = a_desc0;
              uint32_t at = tmem_base + TM_A;
              uint32_t ab = A_READY(0, 8 * src);
#pragma unroll 1
              for (int c = 0; c < k_chunks; c++) {
                if (j == 0) {
                  ptx::mbar_spin(ab, (a_ph >> (8 * src + c)) & 1u);
                  ab += 8;
                }
                ptx::mbar_spin(W_FULL(stage), phase);
                ptx::tc_fence_after();
                const uint64_t bd = b_desc0 + (uint64_t)(stage * (STAGE_BYTES >> 4));
                if (ptx::elect_one()) {
                  if (P.debug & 2) {
                  } else if (src == 0) {
#pragma unroll
                    for (int k = 0; k < KCH / 16; k++)
                      ptx::mma_f16_ss(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (uint32_t)((c | k) != 0));
                  } else {
#pragma unroll
                    for (int k = 0; k < KCH / 16; k++)
                      ptx::mma_f16_ts(d_tmem, at + 8 * k, bd + 2 * k, idesc, (uint32_t)((c | k) != 0));
                  }
                  if (CLUSTER > 1) ptx::mma_commit_multicast(W_EMPTY(stage), CMASK);
                  else ptx::mma_commit(W_EMPTY(stage));
                  if (c == k_chunks - 1) ptx::mma_commit(ACC_FULL(buf));
                }
                __syncwarp();
                ad += A_CHUNK_BYTES >> 4;
                at += KCH / 2;
                if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
              }
              if (j == 0) a_ph ^= ((1u << k_chunks) - 1u) << (8 * src);
            }
          }
        }
      }
    }
  } else {
    // ===================== epilogue: 8 warps, a point per thread pair (32-column halves) ==========
    const int q = warp & 3;              // TMEM lane quarter this warp may read
    const int h = (warp - 2) >> 2;       // which 32 columns of every 64-column group
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;     // 0..255 among epilogue threads
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t a_row = sA + row * 128;
    const int sw = row & 7;
    const float slope = P.head.slope;
    const uint64_t slope2 = ptx::pack_f32x2(slope, slope);
    uint32_t acc_it = 0;
    int bias_buf = 0;

    for (int64_t it = 0; it < iters; it++) {
      const int64_t tile = it * gridDim.x + blockIdx.x;
      const int64_t g = tile * TILE_M + row;
      float x = 0.f, y = 0.f, z = 0.f;
      if (g < P.n) {
        if (P.grid_mode) grid_point(P, P.r_lo + g, x, y, z);
        else { x = __ldg(P.xyz + 3 * g); y = __ldg(P.xyz + 3 * g + 1); z = __ldg(P.xyz + 3 * g + 2); }
      }
      float yc[5] = {0, 0, 0, 0, 0}, yt[5] = {0, 0, 0, 0, 0};

      for (int net = 0; net < 2; net++) {
        if (!((P.net_mask >> net) & 1)) continue;
        const TcNet& N = P.nets[net];
        // ---- layer 0 on CUDA cores: K = 3 after folding the latent into table.w; writes the shared-memory
        // activations (free: their last readers, the MMAs of the previous even layer, completed before the
        // accumulator barrier this thread passed last)
        epi_bar();  // everybody is past the previous use of sTab / sBias
        {
          const float4* t = P.tables + N.l0_table * 512;
          sTab[et] = __ldg(t + et);
          sTab[et + 256] = __ldg(t + et + 256);
          if (N.layers[0].kind == KIND_HIDDEN)
            reinterpret_cast<float2*>(sBias + bias_buf * 512)[et] =
                __ldg(reinterpret_cast<const float2*>(P.bias + N.layers[0].bias_off) + et);
        }
        epi_bar();
#pragma unroll 1
        for (int c = 0; c < 8; c++) {
#pragma unroll
          for (int gq = 0; gq < 4; gq++) {
            float v[8];
#pragma unroll
            for (int f = 0; f < 8; f++) {
              const float4 t = sTab[c * 64 + h * 32 + gq * 8 + f];
              v[f] = leaky(fmaf(t.x, x, fmaf(t.y, y, fmaf(t.z, z, t.w))), slope);
            }
            st_shared_v4(a_row + c * A_CHUNK_BYTES + (((4 * h + gq) ^ sw) << 4), ptx::pack_bf16x2(v[0], v[1]),
                         ptx::pack_bf16x2(v[2], v[3]), ptx::pack_bf16x2(v[4], v[5]), ptx::pack_bf16x2(v[6], v[7]));
          }
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive(A_READY(0, c));
        }

        for (int l = 0; l < N.n_layers; l++) {
          const TcLayer L = N.layers[l];
          if (L.kind == KIND_HEAD) {
            const uint32_t buf = acc_it & 1u;
            ptx::mbar_wait(ACC_FULL(buf), (acc_it >> 1) & 1u, P.err, 0x500 | l);
            acc_it++;
            ptx::tc_fence_after();
            uint32_t r[8];
            if (h == 0) {
              ptx::tmem_ld8(tmem_lane + TM_ACC + buf * NCH, r);
              ptx::tmem_wait_ld();
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(ACC_EMPTY(buf));
            if (h == 0) {
              float* dst = net == 0 ? yc : yt;
#pragma unroll
              for (int i = 0; i < 5; i++) {
                const float v = __uint_as_float(r[i]) + __ldg(P.bias + L.bias_off + i);
                dst[i] = N.head_leaky ? leaky(v, slope) : v;
              }
            }
            continue;
          }
          const int dsti = (l + 1) & 1;  // where the next layer reads its A operand: 0 shared memory, 1 tensor memory
          epi_bar();  // all epilogue warps finished the previous drain: smem tables may be rewritten
          // prefetch what the NEXT layer's drain needs
          const int next_kind = (l + 1 < N.n_layers) ? N.layers[l + 1].kind : -1;
          float2 pre_b = make_float2(0.f, 0.f);
          float4 pre_t0 = make_float4(0, 0, 0, 0), pre_t1 = pre_t0;
          if (next_kind == KIND_HIDDEN) {
            pre_b = __ldg(reinterpret_cast<const float2*>(P.bias + N.layers[l + 1].bias_off) + et);
          } else if (next_kind == KIND_HIDDEN_XYZ) {
            const float4* t = P.tables + N.layers[l + 1].table * 512;
            pre_t0 = __ldg(t + et);
            pre_t1 = __ldg(t + et + 256);
          }
          const float* bias = sBias + bias_buf * 512;
          for (int j = 0; j < L.n_chunks; j++) {
            const uint32_t buf = acc_it & 1u;
            ptx::mbar_wait(ACC_FULL(buf), (acc_it >> 1) & 1u, P.err, 0x600 | j);
            acc_it++;
            ptx::tc_fence_after();
#pragma unroll 1
            for (int half = 0; half < 2; half++) {
              const int col0 = j * NCH + half * 64 + h * 32;  // output column of this layer
              uint32_t r[32];
              ptx::tmem_ld32(tmem_lane + TM_ACC + buf * NCH + half * 64 + h * 32, r);
              ptx::tmem_wait_ld();
              if (half == 1) {
                ptx::tc_fence_before();
                ptx::mbar_arrive(ACC_EMPTY(buf));
              }
              const int c = 2 * j + half;
              uint32_t pk[16];
#pragma unroll
              for (int gq = 0; gq < 4; gq++) {
                if (L.kind == KIND_HIDDEN_XYZ) {
#pragma unroll
                  for (int f = 0; f < 8; f += 2) {
                    const float4 t0 = sTab[col0 + gq * 8 + f], t1 = sTab[col0 + gq * 8 + f + 1];
                    const float v0 = leaky(__uint_as_float(r[gq * 8 + f]) + fmaf(t0.x, x, fmaf(t0.y, y, fmaf(t0.z, z, t0.w))), slope);
                    const float v1 = leaky(__uint_as_float(r[gq * 8 + f + 1]) + fmaf(t1.x, x, fmaf(t1.y, y, fmaf(t1.z, z, t1.w))), slope);
                    pk[gq * 4 + (f >> 1)] = ptx::pack_bf16x2(v0, v1);
                  }
                } else {
                  const float4 b0 = *reinterpret_cast<const float4*>(bias + col0 + gq * 8);
                  const float4 b1 = *reinterpret_cast<const float4*>(bias + col0 + gq * 8 + 4);
                  const uint64_t bb[4] = {ptx::pack_f32x2(b0.x, b0.y), ptx::pack_f32x2(b0.z, b0.w),
                                          ptx::pack_f32x2(b1.x, b1.y), ptx::pack_f32x2(b1.z, b1.w)};
#pragma unroll
                  for (int f = 0; f < 4; f++) {
                    const uint64_t acc = (uint64_t)r[gq * 8 + 2 * f] | ((uint64_t)r[gq * 8 + 2 * f + 1] << 32);
                    const uint64_t t = ptx::add_f32x2(acc, bb[f]);       // FADD2
                    const uint64_t u = ptx::mul_f32x2(t, slope2);        // FMUL2
                    pk[gq * 4 + f] = ptx::pack_bf16x2(fmaxf(ptx::f32x2_lo(t), ptx::f32x2_lo(u)), fmaxf(ptx::f32x2_hi(t), ptx::f32x2_hi(u)));
                  }
                }
                if (dsti == 0)
                  st_shared_v4(a_row + c * A_CHUNK_BYTES + (((4 * h + gq) ^ sw) << 4), pk[gq * 4], pk[gq * 4 + 1],
                               pk[gq * 4 + 2], pk[gq * 4 + 3]);
              }
              if (dsti == 0) {
                ptx::fence_proxy_async_smem();
              } else {
                // bf16 pair (k, k+1) of row m -> TMEM lane m, column k/2: the layout tcgen05.mma reads A in
                ptx::tmem_st16(tmem_lane + TM_A + (uint32_t)(c * (KCH / 2) + h * 16), pk);
                ptx::tmem_wait_st();
                ptx::tc_fence_before();
              }
              ptx::mbar_arrive(A_READY(0, 8 * dsti + c));
            }
          }
          // publish the prefetched tables for the next drain (made visible by its epi_bar)
          if (next_kind == KIND_HIDDEN) {
            bias_buf ^= 1;
            reinterpret_cast<float2*>(sBias + bias_buf * 512)[et] = pre_b;
          } else if (next_kind == KIND_HIDDEN_XYZ) {
            sTab[et] = pre_t0;
            sTab[et + 256] = pre_t1;
          }
        }
      }
      if (h == 0 && g < P.n) write_heads(P.head, yc, yt, P.out, g);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) ptx::cluster_sync();  // nobody exits while a peer may still signal its barriers
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace dj
