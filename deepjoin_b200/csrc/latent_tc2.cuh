// Fused forward + backward of the joint decoder for the latent-code optimisation, on tcgen05 tensor
// cores (sm_100a, CTA pairs / cta_group::2).  One launch = one Adam iteration's worth of decoder work.
//
// replaces, per iteration of reconstruct() (reconstruct.py:107-224): the decoder forward on the
// mini-batch (:147), the clamped-L1 + BCE + normal loss (:166-192) and loss.backward() (:223) --
// restricted to what the optimiser needs: the reference back-propagates into all decoder weights and
// throws that away; only d loss / d [latent, break_latent] is consumed (reconstruct.py:58).
//
// Per tile of 128 samples per CTA (256 per pair) the kernel walks a 28-entry layer program without the
// activations ever leaving the SM (same machinery as mlp_tc2.cuh: weights streamed by the TMA engine
// through an mbarrier ring, A operand ping-ponging between shared and tensor memory, two accumulator
// buffers in TMEM, MMAs issued by the leader CTA for the pair):
//   forward   fnet layer 0 (CUDA cores) -> lin1..lin7 -> head;  hnet layer 0 -> lin1..lin4 -> head.
//             Every hidden drain also records the SIGN of its 64 outputs (LeakyReLU'(x) is all the
//             backward needs of a hidden layer): one 64-bit word per thread and chunk, 832 B per sample
//             for the 13 hidden layers, written to and read back from global memory by the SAME thread.
//   loss      in registers of the thread that owns the sample: heads, B = select(C, T), clamped L1 / BCE /
//             normal MSE, d loss / d head pre-activations (simt::loss_row, shared with the fp32 path);
//             loss terms reduced by warp shuffles -> 3 atomics per warp.
//   backward  d head (10 numbers per sample, split hi/lo into a K = 16 bf16 operand) x W_head^T, then
//             dPre_l = (dPre_{l+1} . W_{l+1}) * LeakyReLU'(pre_l) through the transposed weight stream.
//             The three activations the code enters through (fnet lin0, the re-injection layer, hnet
//             lin0) are column-summed over the tile (warp transpose-reduce -> atomics into colsum[3][512]);
//             grad_finish_kernel turns those into d loss / d code exactly as the fp32 path does.
// Arithmetic: bf16 operands, fp32 accumulation, fp32 loss; the gradient carries bf16 rounding noise
// (~1e-2 relative), the loss terms are those of the bf16 forward.  DJ_PREC_FP32 keeps the exact path.
#pragma once
#include <type_traits>
#include "mlp_simt.cuh"
#include "mlp_tc2.cuh"

namespace dj {
namespace lat2 {

using tc2::A_CHUNK_BYTES;
using tc2::EPI_WARPS;
using tc2::KCH;
using tc2::NCH;
using tc2::NEPI;
using tc2::NSTAGES;
using tc2::NSTAGES_X2;
using tc2::NTHREADS;
using tc2::SM_X;
using tc2::TILE_PTS_X2;
using tc2::SM_A;
using tc2::SM_BAR;
using tc2::SM_BIAS;
using tc2::SM_TAB;
using tc2::SM_W;
using tc2::SMEM_BYTES;
using tc2::STAGE_BYTES;
using tc2::TILE_M;
using tc2::TM_A;
using tc2::TM_ACC;

constexpr int MAX_PROG = 32;
constexpr int MASK_LAYERS = 13;  // hidden activations whose sign the backward needs (8 fnet + 5 hnet)
constexpr int ACC_STRIDE = 3 * 512 + 8;  // per shape: colsum[3][512] | loss terms [8]

enum {
  K_L0 = 0,        // epilogue only: layer 0 on CUDA cores -> shared-memory activations
  K_FWD = 1,       // hidden layer: + bias, LeakyReLU, -> bf16, record signs
  K_FWD_XYZ = 2,   // the re-injection layer: + (w_xyz . xyz + c) instead of a bias
  K_HEAD_F = 3,    // fnet head: 5 pre-activations into registers
  K_HEAD_H = 4,    // hnet head (+ its LeakyReLU), then the loss and the hnet head gradient operand
  K_BWD = 5,       // acc * LeakyReLU'(sign) -> bf16
  K_BWD_SUM = 6,   // same + column sums into colsum[slot]
  K_SUM = 7,       // column sums only (nothing consumes this gradient as a GEMM operand)
};

struct Layer {
  int8_t kind, src, dst, k_steps;   // src / dst: 0 shared memory, 1 tensor memory, 2 none; k_steps: MMAs per ring stage
  int8_t n_chunks, k_stages, mask_layer, slot;  // slot: colsum slot (K_BWD_SUM / K_SUM), table id (K_L0 / K_FWD_XYZ)
  int16_t mma_n, after;             // after: 1 = write the fnet head gradient operand after this layer's last chunk
  int32_t bias_off;
  int32_t stage_bytes, tile_bytes;
  // split-operand program (X2): ring stage s of a chunk reads K range korder[chunk 0 ? 0 : 1] nibble s of the
  // activations (k_stages = 2 k_ranges: W_lo stages, then W_hi; see mlp_tc2.cuh), accumulator * inv_scale
  int32_t k_ranges;
  uint32_t korder[2];
  float inv_scale;
};
struct Params {
  Layer prog[MAX_PROG];
  int32_t n_prog, debug;
  int32_t skip_lo, skip_hi;  // program entries [skip_lo, skip_hi) are not executed (single-net packs: the second net)
  const uint8_t* stream[2];  // per cluster rank, stages in program order
  const float* bias;
  const float4* tables;      // [3][512] per-shape constants (fold_kernel)
  // Several shapes (latent codes) share one launch: tile t belongs to shape t / tiles_per_shape.  Per shape s:
  // tables + s*3*512, sample arrays xyz[s] / sdf[s] / nrm[s], index row idx + s*idx_stride, accumulators
  // acc + s*ACC_STRIDE ([3][512] column sums then [8] loss terms).
  const float* const* xyz;   // device array [n_shapes] of sample pools / batches [.,3]
  const float* const* sdf;   // [.]
  const float* const* nrm;   // [.,3]
  const int32_t* idx;        // nullptr: rows 0..n-1; else the mini-batch's row indices into the pool
  int64_t idx_stride;
  int64_t n;                 // samples per shape
  int32_t n_shapes, tiles_per_shape;
  simt::LossCfg loss;
  unsigned long long* masks;  // [gridDim.x][MASK_LAYERS][4][2][128]: one slot per CTA, reused by every pass
  float* acc;                 // [n_shapes][ACC_STRIDE], zeroed by the caller
  unsigned int* err;
};

// warp transpose-reduce: on return lane L holds sum over the 32 lanes of v[L] (in v[0]) -- 31 shuffles
// Activation-sign words (LeakyReLU' of the backward): the 64 outputs a thread produces for one chunk are recorded in
// two 32-bit words, element e of a word at bit 31 - e.  neg = (+0) - pre has its sign bit set exactly when pre > 0
// (pre = +-0 gives +0), so one funnel shift appends the flag: 1 instruction per element.
__device__ __forceinline__ uint32_t push_sign(uint32_t w, float neg) { return __funnelshift_l(__float_as_uint(neg), w, 1); }
__device__ __forceinline__ bool sign_bit(uint32_t w, int e) { return (w >> (31 - e)) & 1u; }

__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; i++) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// F16: the FORWARD layers (activations + their weight tiles) are IEEE half instead of bf16; gradients stay bf16
// X2 (with F16): split-operand mode (DJ_PREC_FP16X2), forward AND backward.  As in mlp_tc2.cuh a CTA owns 64 samples,
// operand row m < 64 carries the high half and row m + 64 the low half of sample m's activation (forward: IEEE half,
// weights scaled per layer) or gradient (backward: bf16 pairs -- fp32's exponent range, gradients are ~1/N), the weight
// stream carries every stage as W_lo and W_hi, and the drain adds TMEM lanes m and m + 64.  The forward then agrees
// with the fp32 decoder to ~1e-5 (so the loss gates -- clamp window, sign(pred - gt), the B = select(C, T) mask --
// are the fp32 ones) and the gradient is fp32-class; 4x the MMA work of the single-pass step.
template <bool F16, bool X2 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1) latent_tc2_kernel(const __grid_constant__ Params P) {
  constexpr int NST = X2 ? NSTAGES_X2 : NSTAGES;
  constexpr int TILE_PTS = X2 ? TILE_PTS_X2 : TILE_M;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t sA = base + SM_A, sW = base + SM_W, sX = base + SM_X;
  float* sBias = reinterpret_cast<float*>(smem + SM_BIAS);
  float4* sTab = reinterpret_cast<float4*>(smem + SM_TAB);
  const uint32_t bar0 = base + SM_BAR;
  auto W_FULL = [&](int s) { return bar0 + 8u * s; };
  auto W_EMPTY = [&](int s) { return bar0 + 8u * (NSTAGES + s); };
  auto A_READY = [&](int b, int j) { return bar0 + 8u * (2 * NSTAGES + 4 * b + j); };
  auto ACC_FULL = [&](int i) { return bar0 + 8u * (2 * NSTAGES + 8 + i); };
  auto ACC_EMPTY = [&](int i) { return bar0 + 8u * (2 * NSTAGES + 10 + i); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SM_BAR + 8 * (2 * NSTAGES + 12));
  (void)sX;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int64_t ntiles = (int64_t)P.n_shapes * P.tiles_per_shape;
  const int64_t iters = (ntiles + gridDim.x - 1) / gridDim.x;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NST; s++) {
      ptx::mbar_init(W_FULL(s), rank == 0 ? 2 : 1);
      ptx::mbar_init(W_EMPTY(s), 1);
    }
    for (int c = 0; c < 8; c++) ptx::mbar_init(A_READY(0, c), 2 * EPI_WARPS);
    for (int i = 0; i < 2; i++) {
      ptx::mbar_init(ACC_FULL(i), 1);
      ptx::mbar_init(ACC_EMPTY(i), 2 * EPI_WARPS);
    }
    ptx::mbar_fence_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc2(ptx::smem_u32((const void*)tmem_slot), 512);
    ptx::tmem_relinquish2();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== weight producer: this CTA's half of every stage, program order =====================
    uint32_t stage = 0, phase = 0;
    for (int64_t it = 0; it < iters; it++) {
      const uint8_t* src = P.stream[rank];
      for (int l = 0; l < P.n_prog; l++) {
        const int ns = P.prog[l].n_chunks * P.prog[l].k_stages;
        const uint32_t bytes = (uint32_t)P.prog[l].stage_bytes;
        if (l >= P.skip_lo && l < P.skip_hi) { src += (size_t)ns * bytes; continue; }
        for (int s = 0; s < ns; s++) {
          ptx::mbar_wait(W_EMPTY(stage), phase ^ 1, P.err, 0x100 | stage);
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(W_FULL(stage), bytes);
            ptx::bulk_g2s(sW + stage * STAGE_BYTES, src, bytes, W_FULL(stage));
          }
          __syncwarp();
          src += bytes;
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && rank != 0) {
    // ===================== peer: relay "my half of the stage has landed" to the leader =====================
    int per_pass = 0;
    for (int l = 0; l < P.n_prog; l++)
      if (l < P.skip_lo || l >= P.skip_hi) per_pass += P.prog[l].n_chunks * P.prog[l].k_stages;
    const uint32_t lead_full0 = ptx::mapa(W_FULL(0), 0);
    uint32_t stage = 0, phase = 0;
    for (int64_t i = 0; i < iters * per_pass; i++) {
      ptx::mbar_spin(W_FULL(stage), phase);
      if (ptx::elect_one()) ptx::mbar_arrive_cluster(lead_full0 + 8u * stage);
      __syncwarp();
      if (++stage == NST) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== leader: MMA issuer for the pair =====================
    const uint64_t desc_hi = ptx::umma_desc_sw128_kmajor(0);
    const uint64_t a_desc0 = desc_hi + (uint64_t)((sA & 0x3FFFF) >> 4);
    const uint64_t b_desc0 = desc_hi + (uint64_t)((sW & 0x3FFFF) >> 4);
    uint32_t stage = 0, phase = 0, a_ph = 0, acc_it = 0;
    uint32_t ok0 = 0, ok1 = 0;
    for (int64_t it = 0; it < iters; it++) {
      for (int l = 0; l < P.n_prog; l++) {
        const int n_chunks = P.prog[l].n_chunks, k_stages = P.prog[l].k_stages, k_steps = P.prog[l].k_steps;
        if (n_chunks == 0 || (l >= P.skip_lo && l < P.skip_hi)) continue;  // epilogue-only / skipped entry
        const bool fwd_l = P.prog[l].kind <= K_HEAD_H;
        const uint32_t idesc = (F16 && fwd_l) ? ptx::umma_idesc_f16(256, (uint32_t)P.prog[l].mma_n) : ptx::umma_idesc_bf16(256, (uint32_t)P.prog[l].mma_n);
        const uint32_t tile16 = (uint32_t)P.prog[l].tile_bytes >> 4;
        const int src = P.prog[l].src;
        const int k_ranges = X2 ? P.prog[l].k_ranges : k_stages;
        const uint32_t korder0 = X2 ? P.prog[l].korder[0] : 0x76543210u, korder1 = X2 ? P.prog[l].korder[1] : 0x76543210u;
        for (int j = 0; j < n_chunks; j++) {
          const uint32_t buf = acc_it & 1u;
          const uint32_t korder = j == 0 ? korder0 : korder1;
          ptx::mbar_spin_cluster(ACC_EMPTY(buf), ((acc_it >> 1) & 1u) ^ 1u);
          acc_it++;
          const uint32_t d_tmem = tmem_base + TM_ACC + buf * NCH;
          auto issue_stage = [&](uint32_t st, int sidx) {
            const uint64_t bd = b_desc0 + (uint64_t)(st * (STAGE_BYTES >> 4));
            if (k_steps == 1) {  // K = 16 operand (head gradient): one MMA, A in the first 32 B of the shared-memory rows
              ptx::mma2_f16_ss(d_tmem, a_desc0, bd, idesc, 0u);
              return;
            }
            const int ak = (int)((korder >> (4 * sidx)) & 15u);  // K range of the activations this stage reads
            if (src == 0) {
              const uint64_t ad = a_desc0 + (uint64_t)(ak * 2 * (A_CHUNK_BYTES >> 4));
#pragma unroll
              for (int kc = 0; kc < 2; kc++)
#pragma unroll
                for (int k = 0; k < KCH / 16; k++)
                  ptx::mma2_f16_ss(d_tmem, ad + (uint64_t)(kc * (A_CHUNK_BYTES >> 4) + 2 * k), bd + (uint64_t)(kc * tile16 + 2 * k), idesc,
                                   (uint32_t)((sidx | kc | k) != 0));
            } else {
              const uint32_t at = tmem_base + TM_A + (uint32_t)(ak * KCH);
#pragma unroll
              for (int kc = 0; kc < 2; kc++)
#pragma unroll
                for (int k = 0; k < KCH / 16; k++)
                  ptx::mma2_f16_ts(d_tmem, at + (uint32_t)(kc * (KCH / 2) + 8 * k), bd + (uint64_t)(kc * tile16 + 2 * k), idesc,
                                   (uint32_t)((sidx | kc | k) != 0));
            }
          };
          int sidx = 0;
#pragma unroll 1
          while (sidx < k_stages) {
            const bool two = sidx + 1 < k_stages;
            const uint32_t st0 = stage, ph0 = phase;
            const uint32_t st1 = (st0 + 1 == NST) ? 0u : st0 + 1, ph1 = ph0 ^ (uint32_t)(st1 == 0);
            if (j == 0) {
              const int ak0 = (int)((korder >> (4 * sidx)) & 15u), ak1 = (int)((korder >> (4 * sidx + 4)) & 15u);
              ptx::mbar_spin_cluster(A_READY(src, ak0), (a_ph >> (4 * src + ak0)) & 1u);
              if (two && ak1 != ak0) ptx::mbar_spin_cluster(A_READY(src, ak1), (a_ph >> (4 * src + ak1)) & 1u);
            }
            if (!ok0) ptx::mbar_spin_cluster(W_FULL(st0), ph0);
            if (two && !ok1) ptx::mbar_spin_cluster(W_FULL(st1), ph1);
            ptx::tc_fence_after();
            const uint32_t nst0 = two ? ((st1 + 1 == NST) ? 0u : st1 + 1) : st1;
            const uint32_t nph0 = two ? (ph1 ^ (uint32_t)(nst0 == 0)) : ph1;
            const uint32_t nst1 = (nst0 + 1 == NST) ? 0u : nst0 + 1, nph1 = nph0 ^ (uint32_t)(nst1 == 0);
            ok0 = ptx::mbar_test(W_FULL(nst0), nph0);
            ok1 = ptx::mbar_test(W_FULL(nst1), nph1);
            if (ptx::elect_one()) {
              issue_stage(st0, sidx);
              ptx::mma2_commit(W_EMPTY(st0), 3);
              if (two) {
                issue_stage(st1, sidx + 1);
                ptx::mma2_commit(W_EMPTY(st1), 3);
              }
              if (sidx + (two ? 2 : 1) == k_stages) ptx::mma2_commit(ACC_FULL(buf), 3);
            }
            __syncwarp();
            stage = nst0;
            phase = nph0;
            sidx += two ? 2 : 1;
          }
          if (j == 0) a_ph ^= ((1u << (k_steps == 1 ? 1 : k_ranges)) - 1u) << (4 * src);
        }
      }
    }
  } else if constexpr (!X2) {
    // ===================== epilogue: 8 warps; thread = (sample row, 64-column half of a chunk) ==========
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t a_row = sA + row * 128;
    const int sw = row & 7;
    const float slope = P.loss.slope;
    const uint64_t slope2 = ptx::pack_f32x2(slope, slope);
    const uint32_t lead_aready0 = ptx::mapa(A_READY(0, 0), 0);
    const uint32_t lead_accempty0 = ptx::mapa(ACC_EMPTY(0), 0);
    uint32_t acc_it = 0;

    // the K = 16 bf16 operand of a head-gradient GEMM in the first 32 B of this sample's shared-memory row:
    // [g_hi(5) | g_lo(5) | g_hi(5) | 0], matching the rows [W, W, W_lo] of the packed head tile
    auto write_head_operand = [&](const float (&gv)[5]) {
      if (h == 0) {
        uint32_t w[8];
        float hi[5], lo[5];
#pragma unroll
        for (int i = 0; i < 5; i++) {
          hi[i] = __bfloat162float(__float2bfloat16_rn(gv[i]));
          lo[i] = gv[i] - hi[i];
        }
        const float k[16] = {hi[0], hi[1], hi[2], hi[3], hi[4], lo[0], lo[1], lo[2], lo[3], lo[4],
                             hi[0], hi[1], hi[2], hi[3], hi[4], 0.f};
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = ptx::pack_bf16x2(k[2 * i], k[2 * i + 1]);
        tc::st_shared_v4(a_row + ((0 ^ sw) << 4), w[0], w[1], w[2], w[3]);
        tc::st_shared_v4(a_row + ((1 ^ sw) << 4), w[4], w[5], w[6], w[7]);
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(lead_aready0);  // A_READY(shared, K range 0)
    };

    for (int64_t it = 0; it < iters; it++) {
      const int64_t tile = it * gridDim.x + blockIdx.x;
      const int shape_raw = (int)(tile / P.tiles_per_shape);
      const int shape = shape_raw < P.n_shapes ? shape_raw : 0;  // passes beyond the last tile: dead rows of shape 0
      const int64_t g = (tile - (int64_t)shape_raw * P.tiles_per_shape) * TILE_M + row;  // sample within the shape
      const bool live = shape_raw < P.n_shapes && g < P.n;
      const float4* tables = P.tables + (size_t)shape * 3 * 512;
      float* colsum = P.acc + (size_t)shape * ACC_STRIDE;
      float* loss_acc = colsum + 3 * 512;
      float x = 0.f, y = 0.f, z = 0.f, s_gt = 0.f, n_gt[3] = {0.f, 0.f, 0.f};
      if (live) {
        const int64_t r = P.idx ? (int64_t)__ldg(P.idx + (size_t)shape * P.idx_stride + g) : g;
        const float* px = P.xyz[shape];
        x = __ldg(px + 3 * r); y = __ldg(px + 3 * r + 1); z = __ldg(px + 3 * r + 2);
        if (h == 0) {
          const float* pn = P.nrm[shape];
          s_gt = __ldg(P.sdf[shape] + r);
          n_gt[0] = __ldg(pn + 3 * r); n_gt[1] = __ldg(pn + 3 * r + 1); n_gt[2] = __ldg(pn + 3 * r + 2);
        }
      }
      // sign words of THIS CTA's current tile: written by the forward drains and read back by the backward drains of the
      // same pass, so every pass reuses the CTA's 104 KB slot -- 15 MB for the whole grid, resident in L2 and
      // overwritten there (indexed by tile, the 200 MB of a 240,000-sample launch went out to HBM: r01 ncu, 150 MB written)
      unsigned long long* mrow = P.masks + ((size_t)blockIdx.x * MASK_LAYERS * 8 + (size_t)h) * 128 + row;  // + (layer*4 + chunk)*256
      float yc[5] = {0, 0, 0, 0, 0}, yt[5] = {0, 0, 0, 0, 0};
      float gc[5] = {0, 0, 0, 0, 0};

      for (int l = 0; l < P.n_prog; l++) {
        const Layer L = P.prog[l];
        if (l >= P.skip_lo && l < P.skip_hi) continue;
        if (L.kind == K_L0) {
          // ---- layer 0 on CUDA cores (K = 3 after folding the latent into table.w) -> shared-memory activations
          tc::epi_bar();
          {
            const float4* t = tables + L.slot * 512;
            sTab[et] = __ldg(t + et);
            sTab[et + 256] = __ldg(t + et + 256);
          }
          tc::epi_bar();
#pragma unroll 1
          for (int j = 0; j < 4; j++) {
            const int c = 2 * j + h;
            uint32_t mw[2] = {0u, 0u};  // sign words, see push_sign
#pragma unroll
            for (int gq = 0; gq < 8; gq++) {
              float v[8];
#pragma unroll
              for (int f = 0; f < 8; f++) {
                const float4 t = sTab[c * 64 + gq * 8 + f];
                const float pre = fmaf(t.x, x, fmaf(t.y, y, fmaf(t.z, z, t.w)));
                v[f] = tc::leaky(pre, slope);
                mw[gq >> 2] = push_sign(mw[gq >> 2], 0.f - pre);
              }
              tc::st_shared_v4(a_row + c * A_CHUNK_BYTES + ((gq ^ sw) << 4), ptx::pack_op16x2<F16>(v[0], v[1]),
                               ptx::pack_op16x2<F16>(v[2], v[3]), ptx::pack_op16x2<F16>(v[4], v[5]), ptx::pack_op16x2<F16>(v[6], v[7]));
            }
            const unsigned long long m = (unsigned long long)mw[0] | ((unsigned long long)mw[1] << 32);
            mrow[(size_t)(L.mask_layer * 4 + j) * 256] = m;
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(lead_aready0 + 8u * j);
          }
          continue;
        }
        if (L.kind == K_HEAD_F || L.kind == K_HEAD_H) {
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
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(lead_accempty0 + 8u * buf);
          if (h == 0) {
#pragma unroll
            for (int i = 0; i < 5; i++) {
              const float v = __uint_as_float(r[i]) + __ldg(P.bias + L.bias_off + i);
              if (L.kind == K_HEAD_F) yc[i] = v;
              else yt[i] = tc::leaky(v, slope);  // decoder :259-270
            }
          }
          // the loss follows the last head: the hnet's, or the fnet's own when the second net is skipped
          if (L.kind == K_HEAD_H || P.skip_hi > P.skip_lo) {
            // ---- loss and d loss / d head pre-activations of this sample (reconstruct.py:166-192)
            float gt[5] = {0, 0, 0, 0, 0};
            float l3[3] = {0.f, 0.f, 0.f};
            if (h == 0 && live) simt::loss_row(P.loss, yc, yt, s_gt, n_gt, gc, gt, l3);
            if (h == 0) {
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                l3[0] += __shfl_xor_sync(0xffffffffu, l3[0], o);
                l3[1] += __shfl_xor_sync(0xffffffffu, l3[1], o);
                l3[2] += __shfl_xor_sync(0xffffffffu, l3[2], o);
              }
              if (lane < 3) atomicAdd(loss_acc + lane, lane == 0 ? l3[0] : (lane == 1 ? l3[1] : l3[2]));
            }
            // the last readers of the shared-memory activations (the hnet head, or the fnet's last hidden layer)
            // are complete: every accumulator barrier up to this head has been passed
            if (L.kind == K_HEAD_H) write_head_operand(gt);
            else write_head_operand(gc);
          }
          continue;
        }
        // ---------------- GEMM layers with a 128-column chunk drain
        const bool fwd = (L.kind == K_FWD || L.kind == K_FWD_XYZ);
        if (fwd) {
          tc::epi_bar();  // previous users of sBias / sTab are done
          if (L.kind == K_FWD) {
            reinterpret_cast<float2*>(sBias)[et] = __ldg(reinterpret_cast<const float2*>(P.bias + L.bias_off) + et);
          } else {
            const float4* t = tables + L.slot * 512;
            sTab[et] = __ldg(t + et);
            sTab[et + 256] = __ldg(t + et + 256);
          }
          tc::epi_bar();
        }
        auto drain = [&](auto kind_c, auto dst_c, int j) {
          constexpr int KIND = decltype(kind_c)::value;
          constexpr int DST = decltype(dst_c)::value;
          const uint32_t buf = acc_it & 1u;
          ptx::mbar_wait(ACC_FULL(buf), (acc_it >> 1) & 1u, P.err, 0x600 | j);
          acc_it++;
          ptx::tc_fence_after();
          uint32_t r[64];
          ptx::tmem_ld64(tmem_lane + TM_ACC + buf * NCH + h * 64, r);
          ptx::tmem_wait_ld();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(lead_accempty0 + 8u * buf);
          const int col0 = j * NCH + h * 64;
          const int c = 2 * j + h;
          unsigned long long* mptr = mrow + (size_t)(L.mask_layer * 4 + j) * 256;
          constexpr bool BWD = (KIND == K_BWD || KIND == K_BWD_SUM || KIND == K_SUM);
          uint32_t mw[2] = {0u, 0u};  // sign words of this thread's 64 columns (push_sign / sign_bit)
          if (BWD) {
            const unsigned long long m = *mptr;
            mw[0] = (uint32_t)m;
            mw[1] = (uint32_t)(m >> 32);
          }
          const uint64_t neg1 = ptx::pack_f32x2(-1.f, -1.f), zero2 = 0ull;
          uint32_t pk[32];
          float* vf = reinterpret_cast<float*>(r);  // fp32 results in place (column sums)
#pragma unroll
          for (int gq = 0; gq < 8; gq++) {
            float v[8];
            if (KIND == K_FWD_XYZ) {
#pragma unroll
              for (int f = 0; f < 8; f++) {
                const float4 t = sTab[col0 + gq * 8 + f];
                const float pre = __uint_as_float(r[gq * 8 + f]) + fmaf(t.x, x, fmaf(t.y, y, fmaf(t.z, z, t.w)));
                v[f] = tc::leaky(pre, slope);
                mw[gq >> 2] = push_sign(mw[gq >> 2], 0.f - pre);
              }
            } else if (KIND == K_FWD) {
              const float4 b0 = *reinterpret_cast<const float4*>(sBias + col0 + gq * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(sBias + col0 + gq * 8 + 4);
              const uint64_t bb[4] = {ptx::pack_f32x2(b0.x, b0.y), ptx::pack_f32x2(b0.z, b0.w),
                                      ptx::pack_f32x2(b1.x, b1.y), ptx::pack_f32x2(b1.z, b1.w)};
#pragma unroll
              for (int f = 0; f < 4; f++) {
                const uint64_t acc = (uint64_t)r[gq * 8 + 2 * f] | ((uint64_t)r[gq * 8 + 2 * f + 1] << 32);
                const uint64_t t = ptx::add_f32x2(acc, bb[f]);
                const uint64_t u = ptx::mul_f32x2(t, slope2);
                const uint64_t ng = ptx::fma_f32x2(t, neg1, zero2);  // (+0) - pre, both lanes
                v[2 * f] = fmaxf(ptx::f32x2_lo(t), ptx::f32x2_lo(u));
                v[2 * f + 1] = fmaxf(ptx::f32x2_hi(t), ptx::f32x2_hi(u));
                mw[gq >> 2] = push_sign(push_sign(mw[gq >> 2], ptx::f32x2_lo(ng)), ptx::f32x2_hi(ng));
              }
            } else {
#pragma unroll
              for (int f = 0; f < 4; f++) {
                const uint64_t a2 = (uint64_t)r[gq * 8 + 2 * f] | ((uint64_t)r[gq * 8 + 2 * f + 1] << 32);
                const uint64_t u2 = ptx::mul_f32x2(a2, slope2);
                // LeakyReLU'(pre) from the recorded sign
                v[2 * f] = sign_bit(mw[gq >> 2], (gq & 3) * 8 + 2 * f) ? ptx::f32x2_lo(a2) : ptx::f32x2_lo(u2);
                v[2 * f + 1] = sign_bit(mw[gq >> 2], (gq & 3) * 8 + 2 * f + 1) ? ptx::f32x2_hi(a2) : ptx::f32x2_hi(u2);
              }
            }
            if (KIND == K_BWD_SUM || KIND == K_SUM) {
#pragma unroll
              for (int f = 0; f < 8; f++) vf[gq * 8 + f] = v[f];
            }
            if (KIND != K_SUM) {
#pragma unroll
              for (int f = 0; f < 4; f++)
                pk[gq * 4 + f] = (KIND == K_FWD || KIND == K_FWD_XYZ) ? ptx::pack_op16x2<F16>(v[2 * f], v[2 * f + 1]) : ptx::pack_bf16x2(v[2 * f], v[2 * f + 1]);
              if (DST == 0)
                tc::st_shared_v4(a_row + c * A_CHUNK_BYTES + ((gq ^ sw) << 4), pk[gq * 4], pk[gq * 4 + 1], pk[gq * 4 + 2],
                                 pk[gq * 4 + 3]);
            }
          }
          if (!BWD) *mptr = (unsigned long long)mw[0] | ((unsigned long long)mw[1] << 32);
          if (KIND != K_SUM) {
            if (DST == 0) {
              ptx::fence_proxy_async_smem();
            } else {
              ptx::tmem_st32(tmem_lane + TM_A + (uint32_t)(c * (KCH / 2)), pk);
              ptx::tmem_wait_st();
              ptx::tc_fence_before();
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(lead_aready0 + 8u * (4 * DST + j));
          }
          if (KIND == K_BWD_SUM || KIND == K_SUM) {
            // column sums over the tile's rows: warp transpose-reduce, then one atomic per column and warp
            float lo[32], hi[32];
#pragma unroll
            for (int i = 0; i < 32; i++) { lo[i] = vf[i]; hi[i] = vf[32 + i]; }
            const float s0 = warp_colsum32(lo, lane), s1 = warp_colsum32(hi, lane);
            float* cs = colsum + L.slot * 512 + col0;
            atomicAdd(cs + lane, s0);
            atomicAdd(cs + 32 + lane, s1);
          }
        };
        using I0 = std::integral_constant<int, 0>;
        using I1 = std::integral_constant<int, 1>;
        using KF = std::integral_constant<int, K_FWD>;
        using KX = std::integral_constant<int, K_FWD_XYZ>;
        using KB = std::integral_constant<int, K_BWD>;
        using KBS = std::integral_constant<int, K_BWD_SUM>;
        using KS = std::integral_constant<int, K_SUM>;
        for (int j = 0; j < L.n_chunks; j++) {
          switch (L.kind) {
            case K_FWD: if (L.dst) drain(KF{}, I1{}, j); else drain(KF{}, I0{}, j); break;
            case K_FWD_XYZ: if (L.dst) drain(KX{}, I1{}, j); else drain(KX{}, I0{}, j); break;
            case K_BWD: if (L.dst) drain(KB{}, I1{}, j); else drain(KB{}, I0{}, j); break;
            case K_BWD_SUM: if (L.dst) drain(KBS{}, I1{}, j); else drain(KBS{}, I0{}, j); break;
            default: drain(KS{}, I0{}, j); break;
          }
        }
        // the fnet head gradient goes into the shared-memory rows once their last readers (this layer's MMAs,
        // complete with the accumulator barrier of its last chunk) are done
        if (L.after) write_head_operand(gc);
      }
    }
  } else {
    // ===================== epilogue, split-operand mode (X2): see mlp_tc2.cuh for the warp pairing =====================
    // quarters 0, 1 ("high" warps) own sample m = (q & 1) * 32 + lane: its operand row m (high halves), its losses,
    // its sign words; quarters 2, 3 ("low" warps) drain TMEM lanes m + 64 (the low-half rows of the same samples),
    // hand their partial sums to the high warp of the pair through the exchange slot and put the low halves of
    // tensor-memory operands into their lanes.
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    const bool lo_warp = q >= 2;
    const int prow = (q & 1) * 32 + lane;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t a_row = sA + prow * 128;  // high-half row; the low-half row is 64 rows = 8192 B further
    const int sw = prow & 7;
    const uint32_t pair = (uint32_t)((q & 1) * 2 + h);
    const uint32_t xs = sX + pair * 4096u + (uint32_t)lane * 16u;
    const uint32_t B_FULL = 2u + pair, B_FREE = 6u + pair, B_BACK = 10u + pair;
    const float slope = P.loss.slope;
    const uint32_t lead_aready0 = ptx::mapa(A_READY(0, 0), 0);
    const uint32_t lead_accempty0 = ptx::mapa(ACC_EMPTY(0), 0);
    uint32_t acc_it = 0;
    if (!lo_warp) ptx::bar_arrive_n(B_FREE, 64);

    // K = 16 bf16 operand of a head-gradient GEMM, [g_hi(5) | g_lo(5) | g_hi(5) | 0] in the sample's high row, zeros in
    // its low row (the tiles hold [W | W | W_lo]: three of the four partial products, the fourth is below 2^-16)
    auto write_head_operand = [&](const float (&gv)[5]) {
      if (h == 0) {
        uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (!lo_warp) {
          float hi[5], lo[5];
#pragma unroll
          for (int i = 0; i < 5; i++) {
            hi[i] = __bfloat162float(__float2bfloat16_rn(gv[i]));
            lo[i] = gv[i] - hi[i];
          }
          const float k[16] = {hi[0], hi[1], hi[2], hi[3], hi[4], lo[0], lo[1], lo[2], lo[3], lo[4],
                               hi[0], hi[1], hi[2], hi[3], hi[4], 0.f};
#pragma unroll
          for (int i = 0; i < 8; i++) w[i] = ptx::pack_bf16x2(k[2 * i], k[2 * i + 1]);
        }
        const uint32_t my_row = sA + row * 128;
        tc::st_shared_v4(my_row + ((0 ^ sw) << 4), w[0], w[1], w[2], w[3]);
        tc::st_shared_v4(my_row + ((1 ^ sw) << 4), w[4], w[5], w[6], w[7]);
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(lead_aready0);
    };

    for (int64_t it = 0; it < iters; it++) {
      const int64_t tile = it * gridDim.x + blockIdx.x;
      const int shape_raw = (int)(tile / P.tiles_per_shape);
      const int shape = shape_raw < P.n_shapes ? shape_raw : 0;
      const int64_t g = (tile - (int64_t)shape_raw * P.tiles_per_shape) * TILE_PTS + prow;
      const bool live = !lo_warp && shape_raw < P.n_shapes && g < P.n;
      const float4* tables = P.tables + (size_t)shape * 3 * 512;
      float* colsum = P.acc + (size_t)shape * ACC_STRIDE;
      float* loss_acc = colsum + 3 * 512;
      float x = 0.f, y = 0.f, z = 0.f, s_gt = 0.f, n_gt[3] = {0.f, 0.f, 0.f};
      if (live) {
        const int64_t r = P.idx ? (int64_t)__ldg(P.idx + (size_t)shape * P.idx_stride + g) : g;
        const float* px = P.xyz[shape];
        x = __ldg(px + 3 * r); y = __ldg(px + 3 * r + 1); z = __ldg(px + 3 * r + 2);
        if (h == 0) {
          const float* pn = P.nrm[shape];
          s_gt = __ldg(P.sdf[shape] + r);
          n_gt[0] = __ldg(pn + 3 * r); n_gt[1] = __ldg(pn + 3 * r + 1); n_gt[2] = __ldg(pn + 3 * r + 2);
        }
      }
      // sign words of this CTA's current tile (high warps only; slot layout as in the single-pass kernel)
      unsigned long long* mrow = P.masks + ((size_t)blockIdx.x * MASK_LAYERS * 8 + (size_t)h) * 128 + prow;
      float yc[5] = {0, 0, 0, 0, 0}, yt[5] = {0, 0, 0, 0, 0};
      float gc[5] = {0, 0, 0, 0, 0};

      for (int l = 0; l < P.n_prog; l++) {
        const Layer L = P.prog[l];
        if (l >= P.skip_lo && l < P.skip_hi) continue;
        if (L.kind == K_L0) {
          // ---- layer 0 in fp32 on CUDA cores (K = 3 after the fold), split into the high / low rows by the high warps
          tc::epi_bar();
          {
            const float4* t = tables + L.slot * 512;
            sTab[et] = __ldg(t + et);
            sTab[et + 256] = __ldg(t + et + 256);
          }
          tc::epi_bar();
#pragma unroll 1
          for (int j = 0; j < 4; j++) {
            if (!lo_warp) {
              const int c = 2 * j + h;
              uint32_t mw[2] = {0u, 0u};
#pragma unroll
              for (int gq = 0; gq < 8; gq++) {
                uint32_t ph[4], pl[4];
#pragma unroll
                for (int f = 0; f < 4; f++) {
                  float v[2];
#pragma unroll
                  for (int e = 0; e < 2; e++) {
                    const float4 t = sTab[c * 64 + gq * 8 + 2 * f + e];
                    const float pre = fmaf(t.x, x, fmaf(t.y, y, fmaf(t.z, z, t.w)));
                    v[e] = tc::leaky(pre, slope);
                    mw[gq >> 2] = push_sign(mw[gq >> 2], 0.f - pre);
                  }
                  ph[f] = ptx::pack_f16x2(v[0], v[1]);
                  const float2 back = ptx::f16x2_to_float2(ph[f]);
                  pl[f] = ptx::pack_f16x2(v[0] - back.x, v[1] - back.y);
                }
                const uint32_t dsta = a_row + c * A_CHUNK_BYTES + ((gq ^ sw) << 4);
                tc::st_shared_v4(dsta, ph[0], ph[1], ph[2], ph[3]);
                tc::st_shared_v4(dsta + 64 * 128, pl[0], pl[1], pl[2], pl[3]);
              }
              mrow[(size_t)(L.mask_layer * 4 + j) * 256] = (unsigned long long)mw[0] | ((unsigned long long)mw[1] << 32);
              ptx::fence_proxy_async_smem();
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(lead_aready0 + 8u * j);
          }
          continue;
        }
        const float inv = L.inv_scale;
        if (L.kind == K_HEAD_F || L.kind == K_HEAD_H) {
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
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(lead_accempty0 + 8u * buf);
          if (lo_warp) {
            ptx::bar_sync_n(B_FREE, 64);
            if (h == 0) {
              ptx::st_shared_v4u(xs, r[0], r[1], r[2], r[3]);
              ptx::st_shared_v4u(xs + 512, r[4], r[5], r[6], r[7]);
            }
            __threadfence_block();
            ptx::bar_arrive_n(B_FULL, 64);
          } else {
            ptx::bar_sync_n(B_FULL, 64);
            if (h == 0) {
              uint32_t s2[8];
              ptx::ld_shared_v4u(xs, s2[0], s2[1], s2[2], s2[3]);
              ptx::ld_shared_v4u(xs + 512, s2[4], s2[5], s2[6], s2[7]);
#pragma unroll
              for (int i = 0; i < 5; i++) {
                const float v = fmaf(__uint_as_float(r[i]) + __uint_as_float(s2[i]), inv, __ldg(P.bias + L.bias_off + i));
                if (L.kind == K_HEAD_F) yc[i] = v;
                else yt[i] = tc::leaky(v, slope);
              }
            }
            ptx::bar_arrive_n(B_FREE, 64);
          }
          if (L.kind == K_HEAD_H || P.skip_hi > P.skip_lo) {
            float gt[5] = {0, 0, 0, 0, 0};
            float l3[3] = {0.f, 0.f, 0.f};
            if (h == 0 && live) simt::loss_row(P.loss, yc, yt, s_gt, n_gt, gc, gt, l3);
            if (h == 0 && !lo_warp) {
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                l3[0] += __shfl_xor_sync(0xffffffffu, l3[0], o);
                l3[1] += __shfl_xor_sync(0xffffffffu, l3[1], o);
                l3[2] += __shfl_xor_sync(0xffffffffu, l3[2], o);
              }
              if (lane < 3) atomicAdd(loss_acc + lane, lane == 0 ? l3[0] : (lane == 1 ? l3[1] : l3[2]));
            }
            if (L.kind == K_HEAD_H) write_head_operand(gt);
            else write_head_operand(gc);
          }
          continue;
        }
        // ---------------- GEMM layers with a 128-column chunk drain
        const bool fwd = (L.kind == K_FWD || L.kind == K_FWD_XYZ);
        if (fwd) {
          tc::epi_bar();
          if (L.kind == K_FWD) {
            reinterpret_cast<float2*>(sBias)[et] = __ldg(reinterpret_cast<const float2*>(P.bias + L.bias_off) + et);
          } else {
            const float4* t = tables + L.slot * 512;
            sTab[et] = __ldg(t + et);
            sTab[et + 256] = __ldg(t + et + 256);
          }
          tc::epi_bar();
        }
        auto drain = [&](auto kind_c, auto dst_c, int j) {
          constexpr int KIND = decltype(kind_c)::value;
          constexpr int DST = decltype(dst_c)::value;
          constexpr bool BWD = (KIND == K_BWD || KIND == K_BWD_SUM || KIND == K_SUM);
          constexpr bool HAS_DST = KIND != K_SUM;
          const uint32_t buf = acc_it & 1u;
          ptx::mbar_wait(ACC_FULL(buf), (acc_it >> 1) & 1u, P.err, 0x600 | j);
          acc_it++;
          ptx::tc_fence_after();
          uint32_t r[64];
          ptx::tmem_ld64(tmem_lane + TM_ACC + buf * NCH + h * 64, r);
          ptx::tmem_wait_ld();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(lead_accempty0 + 8u * buf);
          const int col0 = j * NCH + h * 64;
          const int c = 2 * j + h;
          if (lo_warp) {
            ptx::bar_sync_n(B_FREE, 64);
#pragma unroll
            for (int gq = 0; gq < 8; gq++) {
              uint32_t w[4];
#pragma unroll
              for (int f = 0; f < 4; f++)
                w[f] = ptx::pack_bf16x2(__uint_as_float(r[gq * 8 + 2 * f]), __uint_as_float(r[gq * 8 + 2 * f + 1]));
              ptx::st_shared_v4u(xs + 512 * gq, w[0], w[1], w[2], w[3]);
            }
            __threadfence_block();
            ptx::bar_arrive_n(B_FULL, 64);
            if (HAS_DST && DST == 1) {
              ptx::bar_sync_n(B_BACK, 64);
              uint32_t pk[32];
#pragma unroll
              for (int gq = 0; gq < 8; gq++)
                ptx::ld_shared_v4u(xs + 512 * gq, pk[gq * 4], pk[gq * 4 + 1], pk[gq * 4 + 2], pk[gq * 4 + 3]);
              ptx::tmem_st32(tmem_lane + TM_A + (uint32_t)(c * (KCH / 2)), pk);
              ptx::tmem_wait_st();
              ptx::tc_fence_before();
            }
          } else {
            unsigned long long* mptr = mrow + (size_t)(L.mask_layer * 4 + j) * 256;
            uint32_t mw[2] = {0u, 0u};
            if (BWD) {
              const unsigned long long m = *mptr;
              mw[0] = (uint32_t)m;
              mw[1] = (uint32_t)(m >> 32);
            }
            ptx::bar_sync_n(B_FULL, 64);
            uint32_t pkh[32];
            float* vf = reinterpret_cast<float*>(r);  // fp32 results in place (column sums)
#pragma unroll
            for (int gq = 0; gq < 8; gq++) {
              uint32_t w[4], pl[4];
              ptx::ld_shared_v4u(xs + 512 * gq, w[0], w[1], w[2], w[3]);
              float v[8];
#pragma unroll
              for (int f = 0; f < 4; f++) {
                const uint64_t acc = (uint64_t)r[gq * 8 + 2 * f] | ((uint64_t)r[gq * 8 + 2 * f + 1] << 32);
                const uint64_t t = ptx::add_f32x2(acc, ptx::bf16x2_to_f32x2(w[f]));
                const float t0 = ptx::f32x2_lo(t), t1 = ptx::f32x2_hi(t);
                if (KIND == K_FWD_XYZ || KIND == K_FWD) {
                  float b0, b1;
                  if (KIND == K_FWD_XYZ) {
                    const float4 ta = sTab[col0 + gq * 8 + 2 * f], tb = sTab[col0 + gq * 8 + 2 * f + 1];
                    b0 = fmaf(ta.x, x, fmaf(ta.y, y, fmaf(ta.z, z, ta.w)));
                    b1 = fmaf(tb.x, x, fmaf(tb.y, y, fmaf(tb.z, z, tb.w)));
                  } else {
                    const float2 bb = *reinterpret_cast<const float2*>(sBias + col0 + gq * 8 + 2 * f);
                    b0 = bb.x;
                    b1 = bb.y;
                  }
                  const float p0 = fmaf(t0, inv, b0), p1 = fmaf(t1, inv, b1);
                  v[2 * f] = tc::leaky(p0, slope);
                  v[2 * f + 1] = tc::leaky(p1, slope);
                  mw[gq >> 2] = push_sign(push_sign(mw[gq >> 2], 0.f - p0), 0.f - p1);
                } else {
                  // LeakyReLU'(pre) from the recorded sign
                  v[2 * f] = (t0 * inv) * (sign_bit(mw[gq >> 2], (gq & 3) * 8 + 2 * f) ? 1.f : slope);
                  v[2 * f + 1] = (t1 * inv) * (sign_bit(mw[gq >> 2], (gq & 3) * 8 + 2 * f + 1) ? 1.f : slope);
                }
              }
              if (KIND == K_BWD_SUM || KIND == K_SUM) {
#pragma unroll
                for (int f = 0; f < 8; f++) vf[gq * 8 + f] = v[f];
              }
              if (HAS_DST) {
#pragma unroll
                for (int f = 0; f < 4; f++) {
                  if (BWD) {  // gradients: bf16 pairs (exponent range)
                    const uint32_t ph = ptx::pack_bf16x2(v[2 * f], v[2 * f + 1]);
                    pkh[gq * 4 + f] = ph;
                    pl[f] = ptx::pack_bf16x2(v[2 * f] - __uint_as_float(ph << 16), v[2 * f + 1] - __uint_as_float(ph & 0xFFFF0000u));
                  } else {    // activations: IEEE-half pairs
                    const uint32_t ph = ptx::pack_f16x2(v[2 * f], v[2 * f + 1]);
                    const float2 back = ptx::f16x2_to_float2(ph);
                    pkh[gq * 4 + f] = ph;
                    pl[f] = ptx::pack_f16x2(v[2 * f] - back.x, v[2 * f + 1] - back.y);
                  }
                }
                if (DST == 0) {
                  const uint32_t dsta = a_row + c * A_CHUNK_BYTES + ((gq ^ sw) << 4);
                  tc::st_shared_v4(dsta, pkh[gq * 4], pkh[gq * 4 + 1], pkh[gq * 4 + 2], pkh[gq * 4 + 3]);
                  tc::st_shared_v4(dsta + 64 * 128, pl[0], pl[1], pl[2], pl[3]);
                } else {
                  ptx::st_shared_v4u(xs + 512 * gq, pl[0], pl[1], pl[2], pl[3]);
                }
              }
            }
            if (!BWD) *mptr = (unsigned long long)mw[0] | ((unsigned long long)mw[1] << 32);
            if (HAS_DST) {
              if (DST == 0) {
                ptx::fence_proxy_async_smem();
              } else {
                __threadfence_block();
                ptx::bar_arrive_n(B_BACK, 64);
                ptx::tmem_st32(tmem_lane + TM_A + (uint32_t)(c * (KCH / 2)), pkh);
                ptx::tmem_wait_st();
                ptx::tc_fence_before();
              }
            }
            ptx::bar_arrive_n(B_FREE, 64);
            if (KIND == K_BWD_SUM || KIND == K_SUM) {
              float lo[32], hi[32];
#pragma unroll
              for (int i = 0; i < 32; i++) { lo[i] = live ? vf[i] : 0.f; hi[i] = live ? vf[32 + i] : 0.f; }
              const float s0 = warp_colsum32(lo, lane), s1 = warp_colsum32(hi, lane);
              float* cs = colsum + L.slot * 512 + col0;
              atomicAdd(cs + lane, s0);
              atomicAdd(cs + 32 + lane, s1);
            }
          }
          if (HAS_DST) {
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(lead_aready0 + 8u * (4 * DST + j));
          }
        };
        using I0 = std::integral_constant<int, 0>;
        using I1 = std::integral_constant<int, 1>;
        using KF = std::integral_constant<int, K_FWD>;
        using KX = std::integral_constant<int, K_FWD_XYZ>;
        using KB = std::integral_constant<int, K_BWD>;
        using KBS = std::integral_constant<int, K_BWD_SUM>;
        using KS = std::integral_constant<int, K_SUM>;
        for (int j = 0; j < L.n_chunks; j++) {
          switch (L.kind) {
            case K_FWD: if (L.dst) drain(KF{}, I1{}, j); else drain(KF{}, I0{}, j); break;
            case K_FWD_XYZ: if (L.dst) drain(KX{}, I1{}, j); else drain(KX{}, I0{}, j); break;
            case K_BWD: if (L.dst) drain(KB{}, I1{}, j); else drain(KB{}, I0{}, j); break;
            case K_BWD_SUM: if (L.dst) drain(KBS{}, I1{}, j); else drain(KBS{}, I0{}, j); break;
            default: drain(KS{}, I0{}, j); break;
          }
        }
        if (L.after) write_head_operand(gc);
      }
    }
    if (lo_warp) ptx::bar_sync_n(B_FREE, 64);
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 1) ptx::tmem_dealloc2(tmem_base, 512);
}

}  // namespace lat2
}  // namespace dj
