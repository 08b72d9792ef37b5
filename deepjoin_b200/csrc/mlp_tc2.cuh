// Fused persistent decoder forward on tcgen05 tensor cores, CTA-pair edition (sm_100a, cta_group::2).
//
// replaces: Decoder.forward  (networks/decoder_z_lb_both_leaky_n_sdf.py:174-457) for one latent
// code shared by all rows, i.e. what decode_samples (core/reconstruct.py:115-129) feeds it.
//
// Same dataflow as mlp_tc.cuh (activations ping-pong between shared memory and tensor memory, two
// accumulator buffers, latent folded into per-shape constants) but the two CTAs of a cluster (one TPC)
// run ONE 256 x 128 x 16 tcgen05.mma.cta_group::2 stream: each CTA keeps its own 128 query points
// (A operand and accumulators in its own shared / tensor memory) and only HALF of every weight tile
// (64 of the 128 output rows), so
//   * each weight byte is fetched from L2 once per 256 points instead of once per 128,
//   * the tensor core reads half as many B bytes from shared memory per MMA,
//   * one 16 KB ring stage per CTA now carries K = 128, i.e. 8 MMAs / 512 tensor-pipe cycles per trip of
//     the issue loop, which hides the loop's own latency (barrier polls, descriptor set-up).
// Only the leader CTA (cluster rank 0) issues MMAs; completion is multicast to both CTAs' barriers
// (tcgen05.commit ... multicast::cluster); the peer's epilogue warps arrive on the leader's barriers
// through mapa / mbarrier.arrive.release.cluster, and the peer's idle MMA warp relays "my half of the
// weight stage has landed" to the leader.
//
// Warp roles (320 threads per CTA): warp 0 = weight producer (TMA engine, 1-D bulk copies),
// warp 1 = MMA issuer (leader) / weight-barrier relay (peer) + TMEM allocation, warps 2..9 = epilogue
// (TMEM lane quarter = warp % 4; warps 2-5 drain columns [0,64) of every 128-column accumulator chunk,
// warps 6-9 columns [64,128)).
#pragma once
#include <type_traits>
#include "mlp_tc.cuh"

namespace dj {
namespace tc2 {

using tc::KIND_HEAD;
using tc::KIND_HIDDEN;
using tc::KIND_HIDDEN_XYZ;

constexpr int TILE_M = 128;                      // query points per CTA (256 per CTA pair)
constexpr int KCH = 64;                          // k-chunk: 64 bf16 = one 128 B swizzle row
constexpr int NCH = 128;                         // accumulator chunk (TMEM columns)
constexpr int A_CHUNK_BYTES = TILE_M * KCH * 2;  // 16 KB
constexpr int STAGE_BYTES = 16384;               // per CTA: two [64 x 64] bf16 tiles = K 128 of its 64 weight rows
constexpr int NSTAGES = 5;
constexpr int MAX_LAYERS = 10;
constexpr int NEPI = 256;
constexpr int NTHREADS = 64 + NEPI;
constexpr int EPI_WARPS = NEPI / 32;

constexpr int SM_A = 0;
constexpr int SM_W = SM_A + 8 * A_CHUNK_BYTES;
constexpr int SM_BIAS = SM_W + NSTAGES * STAGE_BYTES;
constexpr int SM_TAB = SM_BIAS + 2 * 512 * 4;
constexpr int SM_BAR = SM_TAB + 512 * 16;
constexpr int SM_END = SM_BAR + 256;
constexpr int SMEM_BYTES = SM_END + 1024;  // slack for 1024-B alignment

// split-operand mode (X2): the ring has 4 stages and the fifth slot is the exchange buffer between the warp that
// drains TMEM lanes [0,64) (high halves) and the one that drains lanes [64,128) (low halves) of the same points
constexpr int NSTAGES_X2 = 4;
constexpr int SM_X = SM_W + NSTAGES_X2 * STAGE_BYTES;  // 16 KB: 4 warp pairs x 32 lanes x 32 words
constexpr int TILE_PTS_X2 = 64;

// three-pass split-operand mode (X2 = 2): the 128 KB activation area holds the four K ranges of a layer's input as
// separate 64-row a_hi / a_lo tiles (16 KB each: two [64 x 64] k-chunks), K range R at R * 32 KB, single-buffered
constexpr int X3_TILE = 64 * KCH * 2;          // 8 KB: one [64 rows x 64] k-chunk
constexpr int X3_PART = 2 * X3_TILE;           // 16 KB: a_hi (or a_lo) of one K range
constexpr int X3_RANGE = 2 * X3_PART;          // 32 KB: a_hi + a_lo of one K range
constexpr int X3_NCHUNK = 256;                 // output columns per accumulator chunk (N of the MMA)

constexpr uint32_t TM_A = 0;      // TMEM columns [0,256): bf16 activations (A operand of odd layers)
constexpr uint32_t TM_ACC = 256;  // TMEM columns [256,512): two fp32 accumulator buffers

struct Layer {
  int16_t n_chunks;   // 128-column accumulator chunks of this layer
  int16_t k_stages;   // ring stages per chunk (K / 128)
  int16_t mma_n;      // N of each MMA over the pair (128 hidden, 32 head)
  int16_t kind;
  int32_t bias_off;   // float offset into Params::bias (KIND_HIDDEN / KIND_HEAD)
  int32_t table;      // xyzw table id (KIND_HIDDEN_XYZ)
  int32_t stage_bytes;  // per CTA
  int32_t tile_bytes;   // one [mma_n/2 x 64] tile
  int32_t k_steps;      // MMAs per ring stage: 8 (K = 128), or 1 for the K = 16 operand of layer 0
  int16_t k_ranges;     // K / 128 of the activations (k_stages = k_ranges, or 2 k_ranges in split-operand mode)
  int16_t pad;
  uint32_t korder[2];   // nibble s = K range ring stage s of a chunk reads; [0]: chunk 0, [1]: the other chunks
  float inv_scale;      // split-operand mode: the weights of this layer are stored times a power of two (so that the
                        // low halves stay normal IEEE-half numbers); the drain multiplies the accumulator by this
};
constexpr int KIND_L0TC = 3;  // layer 0 on the tensor core: xyz split hi/lo into a K = 16 operand, drained like KIND_HIDDEN
struct Net {
  int32_t n_layers;
  int32_t l0_table;
  int32_t head_leaky;
  int32_t pad;
  const uint8_t* stream[2];  // per cluster rank
  Layer layers[MAX_LAYERS];
};
struct Params {
  Net nets[2];
  int32_t net_mask;
  int32_t grid_mode;
  int32_t debug;  // bench diagnostics only (DEEPJOIN_B200_DEBUG): 1 = do not re-stream weights, 2 = skip the MMAs
  int32_t pad1;
  const float* bias;
  const float4* tables;  // [3][512] (wx, wy, wz, c) per-shape constants
  const float* xyz;
  int64_t n;
  int32_t d0, d1, d2, pad0;
  double step0, step1, step2, stop, offs;
  int64_t r_lo;
  HeadCfg head;
  float* out;
  unsigned int* err;
  unsigned long long* prof;  // [gridDim.x][16] cycle counters, filled when debug & 8 (tools/gpu_diag.py prof)
};

__device__ __forceinline__ void grid_point(const Params& P, int64_t r, float& x, float& y, float& z) {
  int k = (int)(r % P.d2);
  int64_t q = r / P.d2;
  int j = (int)(q % P.d0);
  int i = (int)(q / P.d0);
  x = tc::lin_axis(j, P.d0, P.step0, P.stop, P.offs);
  y = tc::lin_axis(i, P.d1, P.step1, P.stop, P.offs);
  z = tc::lin_axis(k, P.d2, P.step2, P.stop, P.offs);
}

// DBG = 1 compiles the bench diagnostics in (Params::debug knobs and cycle counters); the product path
// launches DBG = 0.
// F16 = 1: IEEE-half operands (DJ_PREC_FP16) instead of bf16 -- same instruction, same speed, 8x finer rounding.
// X2 = 2 (with F16 = 1): the three-pass split-operand program, what DJ_PREC_FP16X2 launches -- described where its
// issuer and its drain are (128 x 256 x 16 MMAs over separate 64-row a_hi / a_lo tiles, two accumulators per chunk).
// X2 = 1 (with F16 = 1): the four-pass split-operand program (DEEPJOIN_B200_X2_FOURPASS=1; the scheme the latent
// step uses).  Every activation a and weight W is carried as a high + low IEEE-half pair (22 significant bits) and
// ALL FOUR partial products are accumulated in fp32:
//   * a CTA owns 64 points; row m < 64 of its A operand holds a_hi of point m, row m + 64 holds a_lo of the same
//     point, so one 256 x 128 x 16 MMA of the pair multiplies both halves with a weight tile;
//   * the weight stream carries the K = 128 stages of a chunk twice, all of W_lo and then all of W_hi (scaled by a
//     power of two per layer so that W_lo stays a normal half), accumulating into the same TMEM columns: lanes
//     [0,64) end up with a_hi.W, lanes [64,128) with a_lo.W.  W_lo goes first because the tensor core truncates
//     (rounds toward zero) whatever it adds to the accumulator at the accumulator's exponent, ~0.46 ulp per MMA
//     (tools/probe_rounding.py): the small a.W_lo terms are summed while the accumulator is still small.  Chunk 0 of
//     a layer takes its LAST K range after everything else (W_lo, W_hi of the other ranges, then W_lo, W_hi of the
//     last): that range is published by the drain of the previous layer's last chunk, which the MMAs of the
//     other ranges hide;
//   * the drain adds the two (the low-half warp hands its 64 columns to the high-half warp through shared memory,
//     as bf16 pairs: it is a 2^-11 correction), applies scale / bias / LeakyReLU in fp32 and splits the result
//     into the next layer's a_hi / a_lo rows.
// 4x the MMA work of the single-pass kernel per point (the a_lo.W_lo quarter is not needed but cannot be
// skipped: both halves of a point ride in the same MMA), error ~1e-5 of fp64 on trained-norm weights, where single
// pass bf16 has 5e-2 and IEEE half 8e-3.  A 3-pass K-stacked variant would need a_hi and a_lo of 128 points per CTA
// on chip (2 x 128 KB per buffer, x2 for the ping-pong): it does not fit next to the accumulators.
// EW: epilogue warps (8; the three-pass program also runs with 16: two warps per scheduler cannot hide the latencies
// of its drain -- TMEM loads, bias loads and stores that queue behind the tensor core's operand reads)
template <int DBG, int F16, int X2 = 0, int EW = 8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EW, 1) tc2_forward_kernel(const __grid_constant__ Params P) {
  static_assert(EW == 8 || (EW == 16 && X2 == 2), "16 epilogue warps: three-pass program only");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t sA = base + SM_A, sW = base + SM_W;
  float* sBias = reinterpret_cast<float*>(smem + SM_BIAS);
  float4* sTab = reinterpret_cast<float4*>(smem + SM_TAB);
  const uint32_t bar0 = base + SM_BAR;
  auto W_FULL = [&](int s) { return bar0 + 8u * s; };
  auto W_EMPTY = [&](int s) { return bar0 + 8u * (NSTAGES + s); };
  // A_READY(b, j): K range [128 j, 128 j + 128) of the activations in buffer b (0 shared memory, 1 tensor
  // memory) is published by BOTH CTAs (leader's barrier; 2 x 8 warp arrivals)
  auto A_READY = [&](int b, int j) { return bar0 + 8u * (2 * NSTAGES + 4 * b + j); };
  auto ACC_FULL = [&](int i) { return bar0 + 8u * (2 * NSTAGES + 8 + i); };
  auto ACC_EMPTY = [&](int i) { return bar0 + 8u * (2 * NSTAGES + 10 + i); };  // leader's barrier, 2 x 8 arrivals
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SM_BAR + 8 * (2 * NSTAGES + 12));
  static_assert(8 * (2 * NSTAGES + 12) + 4 <= 256, "barrier block");

  constexpr int NST = X2 == 1 ? NSTAGES_X2 : NSTAGES;     // ring depth
  constexpr int TILE_PTS = X2 ? TILE_PTS_X2 : TILE_M;    // points per CTA and pass
  const uint32_t sX = base + SM_X;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int64_t ntiles = (P.n + TILE_PTS - 1) / TILE_PTS;
  // both CTAs of a pair (and, for simplicity, all pairs) run the same number of passes; passes beyond
  // the last tile compute on zeros and write nothing
  const int64_t iters = (ntiles + gridDim.x - 1) / gridDim.x;
  const int debug = DBG ? P.debug : 0;
  const bool prof = (debug & 8) != 0;
  auto tick = [&]() -> long long { return prof ? clock64() : 0ll; };
  unsigned long long* pc = P.prof + (size_t)blockIdx.x * 16;
  long long w0 = 0, w1 = 0, w2 = 0, w3 = 0;
  const long long t_begin = tick();

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NST; s++) {
      ptx::mbar_init(W_FULL(s), rank == 0 ? 2 : 1);  // leader: own copy + the peer's relay
      ptx::mbar_init(W_EMPTY(s), 1);
    }
    if constexpr (X2 == 2) {
      // A_READY(0, r): K range r of the next layer's operand is written (4 warps per CTA own a range);
      // A_READY(1, r), r = 0, 1 = "R_FREE(r)": the last MMAs that read K range r of the current layer's operand are
      // done (tcgen05.commit), the drain of chunk 0 may overwrite it
      for (int c = 0; c < 4; c++) ptx::mbar_init(A_READY(0, c), EW);
      for (int c = 0; c < 2; c++) ptx::mbar_init(A_READY(1, c), 1);
    } else {
      for (int c = 0; c < 8; c++) ptx::mbar_init(A_READY(0, c), 2 * EPI_WARPS);
    }
    for (int i = 0; i < 2; i++) {
      ptx::mbar_init(ACC_FULL(i), 1);
      ptx::mbar_init(ACC_EMPTY(i), 2 * EW);
    }
    ptx::mbar_fence_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc2(ptx::smem_u32((const void*)tmem_slot), 512);
    ptx::tmem_relinquish2();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();  // the peer's barriers are initialised before anything is signalled to them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== weight producer: this CTA's half of every stage =====================
    uint32_t stage = 0, phase = 0;
    for (int64_t it = 0; it < iters; it++) {
      for (int net = 0; net < 2; net++) {
        if (!((P.net_mask >> net) & 1)) continue;
        const Net& N = P.nets[net];
        const uint8_t* src = N.stream[rank];
        for (int l = 0; l < N.n_layers; l++) {
          const int ns = N.layers[l].n_chunks * N.layers[l].k_stages;
          const uint32_t bytes = (uint32_t)N.layers[l].stage_bytes;
          for (int s = 0; s < ns; s++) {
            { const long long t = tick(); ptx::mbar_wait(W_EMPTY(stage), phase ^ 1, P.err, 0x100 | stage); w0 += tick() - t; }
            if ((debug & 1) && (it > 0 || net > 0 || l > 0)) {
              if (ptx::elect_one()) ptx::mbar_arrive(W_FULL(stage));
            } else if (ptx::elect_one()) {
              ptx::mbar_expect_tx(W_FULL(stage), bytes);
              ptx::bulk_g2s(sW + stage * STAGE_BYTES, src, bytes, W_FULL(stage));
            }
            __syncwarp();
            src += bytes;
            if (++stage == NST) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    if (prof && lane == 0) { pc[9] = tick() - t_begin; pc[10] = w0; }
  } else if (warp == 1 && rank != 0) {
    // ===================== peer: relay "my half of the stage has landed" to the leader =====================
    int per_pass = 0;
    for (int net = 0; net < 2; net++) {
      if (!((P.net_mask >> net) & 1)) continue;
      for (int l = 0; l < P.nets[net].n_layers; l++) per_pass += P.nets[net].layers[l].n_chunks * P.nets[net].layers[l].k_stages;
    }
    const uint32_t lead_full0 = ptx::mapa(W_FULL(0), 0);
    uint32_t stage = 0, phase = 0;
    for (int64_t i = 0; i < iters * per_pass; i++) {
      { const long long t = tick(); ptx::mbar_spin(W_FULL(stage), phase); w0 += tick() - t; }
      if (ptx::elect_one()) ptx::mbar_arrive_cluster(lead_full0 + 8u * stage);
      __syncwarp();
      if (++stage == NST) { stage = 0; phase ^= 1; }
    }
    if (prof && lane == 0) { pc[11] = tick() - t_begin; pc[12] = w0; }
  } else if (warp == 1) {
    // ===================== leader: MMA issuer for the pair =====================
    if constexpr (X2 == 2) {
      // Three-pass split-operand program: every MMA is 128 x 256 x 16 over the pair (64 rows per CTA -- the a_hi OR
      // the a_lo tile of the CTA's 64 points -- against 128 weight rows per CTA).  Per K range and 64-wide k-chunk:
      // one ring stage of W_lo (4 MMAs with a_hi) and one of W_hi (4 MMAs with a_hi, 4 with a_lo), all into the same
      // accumulator: a_hi.W_lo + a_hi.W_hi + a_lo.W_hi.  All operands come from shared memory (an A operand in tensor
      // memory would have to be present in both lane halves for this shape: the lanes [64,128) that accumulate output
      // columns [128,256) read their own copy -- tools/x3_diag.py blocks).  Chunk 1 of a layer reads K range 1, then
      // range 0, first and commits R_FREE(1), R_FREE(0): the drain of chunk 0, which runs meanwhile, overwrites
      // exactly these two ranges with the next layer's.
      const uint64_t desc_hi = ptx::umma_desc_sw128_kmajor(0);
      const uint64_t a_desc0 = desc_hi + (uint64_t)((sA & 0x3FFFF) >> 4);
      const uint64_t b_desc0 = desc_hi + (uint64_t)((sW & 0x3FFFF) >> 4);
      uint32_t stage = 0, phase = 0, a_ph = 0, acc_it = 0;
      auto next_stage = [&]() { if (++stage == NST) { stage = 0; phase ^= 1; } };
      for (int64_t it = 0; it < iters; it++) {
        for (int net = 0; net < 2; net++) {
          if (!((P.net_mask >> net) & 1)) continue;
          const Net& N = P.nets[net];
          const int nl = N.n_layers;
          for (int l = 0; l < nl; l++) {
            const int n_chunks = N.layers[l].n_chunks, k_steps = N.layers[l].k_steps, k_ranges = N.layers[l].k_ranges;
            const uint32_t korder0 = N.layers[l].korder[0], korder1 = N.layers[l].korder[1];
            const uint32_t idesc = ptx::umma_idesc_f16(128, (uint32_t)N.layers[l].mma_n);
            for (int j = 0; j < n_chunks; j++) {
              const uint32_t buf = acc_it & 1u;
              { const long long t = tick(); ptx::mbar_spin_cluster(ACC_EMPTY(buf), ((acc_it >> 1) & 1u) ^ 1u); w1 += tick() - t; }
              acc_it++;
              // two accumulators per chunk: the a_hi.W_hi products go to `d_tmem`, the two small ones (2^-11 of it)
              // to `d_small`, and the drain adds them in fp32.  tcgen05.mma truncates every accumulation at the
              // accumulator's exponent (~0.46 ulp per MMA, toward zero): kept apart, the 64 small MMAs cost nothing
              // and only the 32 large ones contribute, as in the four-pass program.
              const uint32_t d_tmem = tmem_base + buf * 256u, d_small = d_tmem + 128u;
              if (k_steps == 1) {  // layer 0: one K = 16 MMA per chunk, operand in the first 32 B of the a_hi rows
                if (j == 0) ptx::mbar_spin_cluster(A_READY(0, 0), a_ph & 1u);
                ptx::mbar_spin_cluster(W_FULL(stage), phase);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                  ptx::mma2_f16_ss(d_tmem, a_desc0, b_desc0 + (uint64_t)(stage * (STAGE_BYTES >> 4)), idesc, 0u);
                  ptx::mma2_commit(W_EMPTY(stage), 3);
                  ptx::mma2_commit(ACC_FULL(buf), 3);
                  if (j == 1) { ptx::mma2_commit(A_READY(1, 0), 3); ptx::mma2_commit(A_READY(1, 1), 3); }
                }
                __syncwarp();
                next_stage();
                continue;
              }
              const uint32_t korder = j == 0 ? korder0 : korder1;
#pragma unroll 1
              for (int ri = 0; ri < k_ranges; ri++) {
                const int r = (int)((korder >> (4 * ri)) & 15u);
                if (j == 0) { const long long t = tick(); ptx::mbar_spin_cluster(A_READY(0, r), (a_ph >> r) & 1u); w2 += tick() - t; }
                const uint32_t a_off = (uint32_t)(r * X3_RANGE);
#pragma unroll 1
                for (int kc = 0; kc < 2; kc++) {
                  const uint32_t st0 = stage, ph0 = phase;
                  next_stage();
                  const uint32_t st1 = stage, ph1 = phase;
                  next_stage();
                  { const long long t = tick(); ptx::mbar_spin_cluster(W_FULL(st0), ph0); ptx::mbar_spin_cluster(W_FULL(st1), ph1); w3 += tick() - t; }
                  ptx::tc_fence_after();
                  if (ptx::elect_one()) {
                    const uint64_t b_lo = b_desc0 + (uint64_t)(st0 * (STAGE_BYTES >> 4)), b_hi = b_desc0 + (uint64_t)(st1 * (STAGE_BYTES >> 4));
                    const uint32_t first = (uint32_t)((ri | kc) != 0);
                    const uint64_t ad_hi = a_desc0 + (uint64_t)((a_off + (uint32_t)(kc * X3_TILE)) >> 4), ad_lo = ad_hi + (uint64_t)(X3_PART >> 4);
#pragma unroll
                    for (int k = 0; k < 4; k++) ptx::mma2_f16_ss(d_small, ad_hi + (uint64_t)(2 * k), b_lo + (uint64_t)(2 * k), idesc, k ? 1u : first);
                    ptx::mma2_commit(W_EMPTY(st0), 3);
#pragma unroll
                    for (int k = 0; k < 4; k++) ptx::mma2_f16_ss(d_small, ad_lo + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), idesc, 1u);
#pragma unroll
                    for (int k = 0; k < 4; k++) ptx::mma2_f16_ss(d_tmem, ad_hi + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), idesc, k ? 1u : first);
                    ptx::mma2_commit(W_EMPTY(st1), 3);
                    if (ri + 1 == k_ranges && kc == 1) ptx::mma2_commit(ACC_FULL(buf), 3);
                    // chunk 1 reads K ranges 1, 0 first: once they are done the drain of chunk 0 may overwrite them
                    if (j == 1 && ri < 2 && kc == 1) ptx::mma2_commit(A_READY(1, r), 3);
                  }
                  __syncwarp();
                }
              }
              if (j == 0) a_ph ^= (1u << k_ranges) - 1u;
            }
            if (k_steps == 1) a_ph ^= 1u;
          }
        }
      }
      if (prof && lane == 0) { pc[0] = tick() - t_begin; pc[1] = w1; pc[2] = w2; pc[3] = w3; }
    } else {
    // The whole warp walks the (warp-uniform) issue program; one elected lane issues tcgen05.mma /
    // tcgen05.commit.  GEMM layer l of a net reads its A operand from shared memory when l is even
    // (layer 0 wrote it there) and from tensor memory when l is odd; accumulator buffers alternate.
    const uint64_t desc_hi = ptx::umma_desc_sw128_kmajor(0);
    const uint64_t a_desc0 = desc_hi + (uint64_t)((sA & 0x3FFFF) >> 4);
    const uint64_t b_desc0 = desc_hi + (uint64_t)((sW & 0x3FFFF) >> 4);
    uint32_t stage = 0, phase = 0, a_ph = 0, acc_it = 0;
    // One trip of the issue loop covers up to two ring stages (K = 256, 16 MMAs, 1024 tensor-pipe cycles) so
    // that the loop's own latency (polls, election, descriptor set-up; ~100 dependent instructions of one
    // warp) stays hidden behind the MMAs already queued.  ok0 / ok1: the two stages of the coming trip were
    // already seen complete by the non-blocking probes issued one trip earlier.
    uint32_t ok0 = 0, ok1 = 0;
    for (int64_t it = 0; it < iters; it++) {
      for (int net = 0; net < 2; net++) {
        if (!((P.net_mask >> net) & 1)) continue;
        const Net& N = P.nets[net];
        const int nl = N.n_layers;
        for (int l = 0; l < nl; l++) {
          const int n_chunks = N.layers[l].n_chunks, k_stages = N.layers[l].k_stages, k_steps = N.layers[l].k_steps;
          const int k_ranges = N.layers[l].k_ranges;
          const uint32_t korder0 = N.layers[l].korder[0], korder1 = N.layers[l].korder[1];
          const uint32_t idesc = F16 ? ptx::umma_idesc_f16(256, (uint32_t)N.layers[l].mma_n) : ptx::umma_idesc_bf16(256, (uint32_t)N.layers[l].mma_n);
          const uint32_t tile16 = (uint32_t)N.layers[l].tile_bytes >> 4;
          const int src = l & 1;
          for (int j = 0; j < n_chunks; j++) {
            const uint32_t buf = acc_it & 1u;
            const uint32_t korder = j == 0 ? korder0 : korder1;
            { const long long t = tick(); ptx::mbar_spin_cluster(ACC_EMPTY(buf), ((acc_it >> 1) & 1u) ^ 1u); w1 += tick() - t; }
            acc_it++;
            const uint32_t d_tmem = tmem_base + TM_ACC + buf * NCH;
            // the 8 MMAs of ring stage st = K range [128 sidx, 128 sidx + 128) of this chunk
            auto issue_stage = [&](uint32_t st, int sidx) {
              const uint64_t bd = b_desc0 + (uint64_t)(st * (STAGE_BYTES >> 4));
              if (debug & 2) return;
              if (k_steps == 1) {  // layer 0: K = 16 operand in the first 32 B of the shared-memory rows
                ptx::mma2_f16_ss(d_tmem, a_desc0, bd, idesc, 0u);
                return;
              }
              const int ak = (int)((korder >> (4 * sidx)) & 15u);  // K range [128 ak, 128 ak + 128) of the activations
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
                const long long t = tick();
                const int ak0 = (int)((korder >> (4 * sidx)) & 15u), ak1 = (int)((korder >> (4 * sidx + 4)) & 15u);
                ptx::mbar_spin_cluster(A_READY(src, ak0), (a_ph >> (4 * src + ak0)) & 1u);
                if (two && ak1 != ak0) ptx::mbar_spin_cluster(A_READY(src, ak1), (a_ph >> (4 * src + ak1)) & 1u);
                w2 += tick() - t;
              }
              if (!ok0) { const long long t = tick(); ptx::mbar_spin_cluster(W_FULL(st0), ph0); w3 += tick() - t; }
              if (two && !ok1) { const long long t = tick(); ptx::mbar_spin_cluster(W_FULL(st1), ph1); w3 += tick() - t; }
              ptx::tc_fence_after();
              // probe the coming trip's barriers now: the polls' latency hides behind the MMA issue below
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
            if (j == 0) a_ph ^= ((1u << k_ranges) - 1u) << (4 * src);
          }
        }
      }
    }
    if (prof && lane == 0) { pc[0] = tick() - t_begin; pc[1] = w1; pc[2] = w2; pc[3] = w3; }
    }
  } else if constexpr (X2 == 0) {
    // ===================== epilogue: 8 warps; thread = (row, 64-column half of the chunk) ==========
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int h = (warp - 2) >> 2;       // columns [64 h, 64 h + 64) of every 128-column chunk
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;     // 0..255 among epilogue threads
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t a_row = sA + row * 128;
    const int sw = row & 7;
    const float slope = P.head.slope;
    const uint64_t slope2 = ptx::pack_f32x2(slope, slope);
    const uint32_t lead_aready0 = ptx::mapa(A_READY(0, 0), 0);
    const uint32_t lead_accempty0 = ptx::mapa(ACC_EMPTY(0), 0);
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
        const Net& N = P.nets[net];
        // ---- layer 0 goes to the tensor core as well: after folding the latent into the per-shape constant c it
        // is c + W_xyz . (x, y, z), K = 3.  The point is split hi + lo into 16-bit pieces and multiplied against
        // [W_hi | W_hi | W_lo] (K = 16 operand [p_hi | p_lo | p_hi | 0], error ~2^-17 of |W p|, far below the
        // rounding of the activation it produces); c is the layer's bias.  The operand goes into the first 32 B of
        // this point's shared-memory row (free: the last readers of the shared-memory activations, the MMAs of the
        // previous layer that read them, completed before the accumulator barrier this thread passed last).
        const long long t_l0 = tick();
        tc::epi_bar();  // everybody is past the previous use of sBias
        {
          const float4* t = P.tables + N.l0_table * 512;
          reinterpret_cast<float2*>(sBias + bias_buf * 512)[et] = make_float2(__ldg(&t[2 * et].w), __ldg(&t[2 * et + 1].w));
        }
        if (h == 0) {
          float hi[3], lo[3];
          const float p3[3] = {x, y, z};
#pragma unroll
          for (int i = 0; i < 3; i++) {
            hi[i] = ptx::round_op16<F16>(p3[i]);
            lo[i] = p3[i] - hi[i];
          }
          const float k[16] = {hi[0], hi[1], hi[2], lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; i++) w[i] = ptx::pack_op16x2<F16>(k[2 * i], k[2 * i + 1]);
          tc::st_shared_v4(a_row + ((0 ^ sw) << 4), w[0], w[1], w[2], w[3]);
          tc::st_shared_v4(a_row + ((1 ^ sw) << 4), w[4], w[5], w[6], w[7]);
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(lead_aready0);  // A_READY(shared memory, K range 0)
        w2 += tick() - t_l0;

        for (int l = 0; l < N.n_layers; l++) {
          const Layer L = N.layers[l];
          if (L.kind == KIND_HEAD) {
            const uint32_t buf = acc_it & 1u;
            { const long long t = tick(); ptx::mbar_wait(ACC_FULL(buf), (acc_it >> 1) & 1u, P.err, 0x500 | l); w3 += tick() - t; }
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
              float* dst = net == 0 ? yc : yt;
#pragma unroll
              for (int i = 0; i < 5; i++) {
                const float v = __uint_as_float(r[i]) + __ldg(P.bias + L.bias_off + i);
                dst[i] = N.head_leaky ? tc::leaky(v, slope) : v;
              }
            }
            continue;
          }
          const int dsti = (l + 1) & 1;  // where the next layer reads its A operand: 0 shared memory, 1 tensor memory
          { const long long t = tick(); tc::epi_bar(); w1 += tick() - t; }  // all epilogue warps finished the previous drain: smem tables may be rewritten
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

          // drain of one accumulator chunk, specialised on (layer kind, destination) so that the 64 columns
          // of a thread are straight-line code
          auto drain = [&](auto kind_c, auto dst_c, int j) {
            constexpr int KIND = decltype(kind_c)::value;
            constexpr int DST = decltype(dst_c)::value;
            const uint32_t buf = acc_it & 1u;
            { const long long t = tick(); ptx::mbar_wait(ACC_FULL(buf), (acc_it >> 1) & 1u, P.err, 0x600 | j); w0 += tick() - t; }
            acc_it++;
            ptx::tc_fence_after();
            uint32_t r[64];
            ptx::tmem_ld64(tmem_lane + TM_ACC + buf * NCH + h * 64, r);
            ptx::tmem_wait_ld();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(lead_accempty0 + 8u * buf);  // accumulator buffer is free again
            const int col0 = j * NCH + h * 64;  // output column of this layer
            const int c = 2 * j + h;            // 64-wide K chunk of the next layer this thread writes
            uint32_t pk[32];
#pragma unroll
            for (int gq = 0; gq < 8; gq++) {
              if (KIND == KIND_HIDDEN_XYZ) {
#pragma unroll
                for (int f = 0; f < 8; f += 2) {
                  const float4 t0 = sTab[col0 + gq * 8 + f], t1 = sTab[col0 + gq * 8 + f + 1];
                  const float v0 = tc::leaky(__uint_as_float(r[gq * 8 + f]) + fmaf(t0.x, x, fmaf(t0.y, y, fmaf(t0.z, z, t0.w))), slope);
                  const float v1 = tc::leaky(__uint_as_float(r[gq * 8 + f + 1]) + fmaf(t1.x, x, fmaf(t1.y, y, fmaf(t1.z, z, t1.w))), slope);
                  pk[gq * 4 + (f >> 1)] = ptx::pack_op16x2<F16>(v0, v1);
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
                  pk[gq * 4 + f] = ptx::pack_op16x2<F16>(fmaxf(ptx::f32x2_lo(t), ptx::f32x2_lo(u)), fmaxf(ptx::f32x2_hi(t), ptx::f32x2_hi(u)));
                }
              }
              if (DST == 0)
                tc::st_shared_v4(a_row + c * A_CHUNK_BYTES + ((gq ^ sw) << 4), pk[gq * 4], pk[gq * 4 + 1], pk[gq * 4 + 2],
                                 pk[gq * 4 + 3]);
            }
            if (DST == 0) {
              ptx::fence_proxy_async_smem();
            } else {
              // bf16 pair (k, k+1) of row m -> TMEM lane m, column k/2: the layout tcgen05.mma reads A in
              ptx::tmem_st32(tmem_lane + TM_A + (uint32_t)(c * (KCH / 2)), pk);
              ptx::tmem_wait_st();
              ptx::tc_fence_before();
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(lead_aready0 + 8u * (4 * DST + j));
          };
          using I0 = std::integral_constant<int, 0>;
          using I1 = std::integral_constant<int, 1>;
          if (L.kind == KIND_HIDDEN_XYZ) {
            if (dsti) for (int j = 0; j < L.n_chunks; j++) drain(I1{}, I1{}, j);
            else for (int j = 0; j < L.n_chunks; j++) drain(I1{}, I0{}, j);
          } else {  // KIND_HIDDEN and KIND_L0TC: bias + LeakyReLU
            if (dsti) for (int j = 0; j < L.n_chunks; j++) drain(I0{}, I1{}, j);
            else for (int j = 0; j < L.n_chunks; j++) drain(I0{}, I0{}, j);
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
    if (prof && warp == 2 && lane == 0) { pc[4] = tick() - t_begin; pc[5] = w0; pc[6] = w1; pc[7] = w2; pc[8] = w3; }
  } else if constexpr (X2 == 1) {
    // ===================== epilogue, split-operand mode (X2) ==========
    // TMEM lane quarter q = warp % 4: quarters 0, 1 hold the a_hi rows of the tile's 64 points ("high" warps),
    // quarters 2, 3 the a_lo rows of the same points ("low" warps).  High warp (q, h) and low warp (q + 2, h) drain
    // the same 64 columns of the same 32 points and form a pair with a 4 KB exchange slot and three named barriers:
    //   B_FULL: low -> high, "my partial sums are in the slot";  B_FREE: high -> low, "the slot may be rewritten";
    //   B_BACK: high -> low, "the a_lo words you have to put into your TMEM lanes are in the slot".
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    const bool lo_warp = q >= 2;
    const int prow = (q & 1) * 32 + lane;  // point of the tile, 0..63
    const int row = q * 32 + lane;         // operand row / TMEM lane of this thread: prow (a_hi) or prow + 64 (a_lo)
    const int et = threadIdx.x - 64;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t a_row = sA + prow * 128;  // a_hi row; the a_lo row of the same point is 64 rows = 8192 B further
    const int sw = prow & 7;
    const uint32_t pair = (uint32_t)((q & 1) * 2 + h);
    const uint32_t xs = sX + pair * 4096u + (uint32_t)lane * 16u;  // 16-byte granule g of this lane at xs + 512 g
    const uint32_t B_FULL = 2u + pair, B_FREE = 6u + pair, B_BACK = 10u + pair;
    const float slope = P.head.slope;
    const uint64_t slope2 = ptx::pack_f32x2(slope, slope);
    const uint32_t lead_aready0 = ptx::mapa(A_READY(0, 0), 0);
    const uint32_t lead_accempty0 = ptx::mapa(ACC_EMPTY(0), 0);
    uint32_t acc_it = 0;
    int bias_buf = 0;
    if (!lo_warp) ptx::bar_arrive_n(B_FREE, 64);  // the slot starts out free

    for (int64_t it = 0; it < iters; it++) {
      const int64_t tile = it * gridDim.x + blockIdx.x;
      const int64_t g = tile * TILE_PTS + prow;
      float x = 0.f, y = 0.f, z = 0.f;
      if (!lo_warp && g < P.n) {
        if (P.grid_mode) grid_point(P, P.r_lo + g, x, y, z);
        else { x = __ldg(P.xyz + 3 * g); y = __ldg(P.xyz + 3 * g + 1); z = __ldg(P.xyz + 3 * g + 2); }
      }
      float yc[5] = {0, 0, 0, 0, 0}, yt[5] = {0, 0, 0, 0, 0};

      for (int net = 0; net < 2; net++) {
        if (!((P.net_mask >> net) & 1)) continue;
        const Net& N = P.nets[net];
        // ---- layer 0: K = 16 operand [p_hi | p_lo | p_hi | 0] in the point's a_hi row (the tiles hold
        // [W_hi | W_hi | W_lo]: three of the four partial products, the fourth is below 2^-22); zeros in the a_lo row
        tc::epi_bar();  // everybody is past the previous use of sBias
        {
          const float4* t = P.tables + N.l0_table * 512;
          reinterpret_cast<float2*>(sBias + bias_buf * 512)[et] = make_float2(__ldg(&t[2 * et].w), __ldg(&t[2 * et + 1].w));
        }
        if (h == 0) {
          uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          if (!lo_warp) {
            float hi[3], lo[3];
            const float p3[3] = {x, y, z};
#pragma unroll
            for (int i = 0; i < 3; i++) {
              hi[i] = ptx::round_op16<1>(p3[i]);
              lo[i] = p3[i] - hi[i];
            }
            const float k[16] = {hi[0], hi[1], hi[2], lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 8; i++) w[i] = ptx::pack_op16x2<1>(k[2 * i], k[2 * i + 1]);
          }
          const uint32_t my_row = sA + row * 128;
          tc::st_shared_v4(my_row + ((0 ^ sw) << 4), w[0], w[1], w[2], w[3]);
          tc::st_shared_v4(my_row + ((1 ^ sw) << 4), w[4], w[5], w[6], w[7]);
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(lead_aready0);  // A_READY(shared memory, K range 0)

        for (int l = 0; l < N.n_layers; l++) {
          const Layer L = N.layers[l];
          const float inv = L.inv_scale;
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
                uint32_t s[8];
                ptx::ld_shared_v4u(xs, s[0], s[1], s[2], s[3]);
                ptx::ld_shared_v4u(xs + 512, s[4], s[5], s[6], s[7]);
                float* dst = net == 0 ? yc : yt;
#pragma unroll
                for (int i = 0; i < 5; i++) {
                  const float v = fmaf(__uint_as_float(r[i]) + __uint_as_float(s[i]), inv, __ldg(P.bias + L.bias_off + i));
                  dst[i] = N.head_leaky ? tc::leaky(v, slope) : v;
                }
              }
              ptx::bar_arrive_n(B_FREE, 64);
            }
            continue;
          }
          const int dsti = (l + 1) & 1;
          tc::epi_bar();
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
          const uint64_t inv2 = ptx::pack_f32x2(inv, inv);

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
            if (lane == 0) ptx::mbar_arrive_cluster(lead_accempty0 + 8u * buf);  // accumulator buffer is free again
            const int col0 = j * NCH + h * 64;
            const int c = 2 * j + h;  // 64-wide K chunk of the next layer this warp pair produces
            if (lo_warp) {
              // a_lo . W of my 64 columns -> bf16 pairs -> the pair's slot
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
              if (DST == 1) {
                // the a_lo halves of the next layer's operand belong in MY TMEM lanes: the high warp hands them back
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
              ptx::bar_sync_n(B_FULL, 64);
              uint32_t pkh[32];
#pragma unroll
              for (int gq = 0; gq < 8; gq++) {
                uint32_t w[4], pl[4];
                ptx::ld_shared_v4u(xs + 512 * gq, w[0], w[1], w[2], w[3]);
                float v[8];
                if (KIND == KIND_HIDDEN_XYZ) {
#pragma unroll
                  for (int f = 0; f < 4; f++) {
                    const uint64_t acc = (uint64_t)r[gq * 8 + 2 * f] | ((uint64_t)r[gq * 8 + 2 * f + 1] << 32);
                    const uint64_t t = ptx::add_f32x2(acc, ptx::bf16x2_to_f32x2(w[f]));
                    const float4 t0 = sTab[col0 + gq * 8 + 2 * f], t1 = sTab[col0 + gq * 8 + 2 * f + 1];
                    const float s0 = fmaf(t0.x, x, fmaf(t0.y, y, fmaf(t0.z, z, t0.w)));
                    const float s1 = fmaf(t1.x, x, fmaf(t1.y, y, fmaf(t1.z, z, t1.w)));
                    v[2 * f] = tc::leaky(fmaf(ptx::f32x2_lo(t), inv, s0), slope);
                    v[2 * f + 1] = tc::leaky(fmaf(ptx::f32x2_hi(t), inv, s1), slope);
                  }
                } else {
                  const float4 b0 = *reinterpret_cast<const float4*>(bias + col0 + gq * 8);
                  const float4 b1 = *reinterpret_cast<const float4*>(bias + col0 + gq * 8 + 4);
                  const uint64_t bb[4] = {ptx::pack_f32x2(b0.x, b0.y), ptx::pack_f32x2(b0.z, b0.w),
                                          ptx::pack_f32x2(b1.x, b1.y), ptx::pack_f32x2(b1.z, b1.w)};
#pragma unroll
                  for (int f = 0; f < 4; f++) {
                    const uint64_t acc = (uint64_t)r[gq * 8 + 2 * f] | ((uint64_t)r[gq * 8 + 2 * f + 1] << 32);
                    const uint64_t t = ptx::fma_f32x2(ptx::add_f32x2(acc, ptx::bf16x2_to_f32x2(w[f])), inv2, bb[f]);
                    const uint64_t u = ptx::mul_f32x2(t, slope2);
                    v[2 * f] = fmaxf(ptx::f32x2_lo(t), ptx::f32x2_lo(u));
                    v[2 * f + 1] = fmaxf(ptx::f32x2_hi(t), ptx::f32x2_hi(u));
                  }
                }
                if (DBG && (debug & 64)) {  // per-layer checksum of the activations (tools/x3_diag.py)
                  float cs = 0.f, cw = 0.f;
#pragma unroll
                  for (int f = 0; f < 8; f++) { cs += v[f]; cw += v[f] * (float)(col0 + gq * 8 + f + 1); }
                  atomicAdd(reinterpret_cast<double*>(P.prof) + net * 16 + l, (double)cs);
                  atomicAdd(reinterpret_cast<double*>(P.prof) + 32 + net * 16 + l, (double)cw);
                }
#pragma unroll
                for (int f = 0; f < 4; f++) {
                  const uint32_t ph = ptx::pack_f16x2(v[2 * f], v[2 * f + 1]);
                  const float2 back = ptx::f16x2_to_float2(ph);
                  pkh[gq * 4 + f] = ph;
                  pl[f] = ptx::pack_f16x2(v[2 * f] - back.x, v[2 * f + 1] - back.y);
                }
                if (DST == 0) {
                  const uint32_t dsta = a_row + c * A_CHUNK_BYTES + ((gq ^ sw) << 4);
                  tc::st_shared_v4(dsta, pkh[gq * 4], pkh[gq * 4 + 1], pkh[gq * 4 + 2], pkh[gq * 4 + 3]);
                  tc::st_shared_v4(dsta + 64 * 128, pl[0], pl[1], pl[2], pl[3]);
                } else {
                  ptx::st_shared_v4u(xs + 512 * gq, pl[0], pl[1], pl[2], pl[3]);  // read above by this same thread
                }
              }
              if (DST == 0) {
                ptx::fence_proxy_async_smem();
              } else {
                __threadfence_block();
                ptx::bar_arrive_n(B_BACK, 64);
                ptx::tmem_st32(tmem_lane + TM_A + (uint32_t)(c * (KCH / 2)), pkh);
                ptx::tmem_wait_st();
                ptx::tc_fence_before();
              }
              ptx::bar_arrive_n(B_FREE, 64);
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(lead_aready0 + 8u * (4 * DST + j));
          };
          using I0 = std::integral_constant<int, 0>;
          using I1 = std::integral_constant<int, 1>;
          if (L.kind == KIND_HIDDEN_XYZ) {
            if (dsti) for (int j = 0; j < L.n_chunks; j++) drain(I1{}, I1{}, j);
            else for (int j = 0; j < L.n_chunks; j++) drain(I1{}, I0{}, j);
          } else {
            if (dsti) for (int j = 0; j < L.n_chunks; j++) drain(I0{}, I1{}, j);
            else for (int j = 0; j < L.n_chunks; j++) drain(I0{}, I0{}, j);
          }
          if (next_kind == KIND_HIDDEN) {
            bias_buf ^= 1;
            reinterpret_cast<float2*>(sBias + bias_buf * 512)[et] = pre_b;
          } else if (next_kind == KIND_HIDDEN_XYZ) {
            sTab[et] = pre_t0;
            sTab[et + 256] = pre_t1;
          }
        }
      }
      if (!lo_warp && h == 0 && g < P.n) write_heads(P.head, yc, yt, P.out, g);
    }
    if (lo_warp) ptx::bar_sync_n(B_FREE, 64);  // consume the last hand-back so that no barrier is left half-arrived
  } else {
    // ===================== epilogue, three-pass split-operand mode (X2 = 2) ==========
    // An accumulator chunk is 64 points x 256 output columns: TMEM lanes [0,64) hold columns [0,128) of the chunk,
    // lanes [64,128) columns [128,256) (the 128 x 256 cta_group::2 accumulator layout, tools/probe_m128).  Warp
    // (q = warp % 4, h): points 32 (q & 1) .. + 31, columns 128 (q >> 1) + 64 h .. + 63 of the chunk -- every warp
    // holds complete sums, applies scale / bias / LeakyReLU, splits the result into a_hi + a_lo and writes both into
    // the shared-memory tiles of K range R = 2 j + (q >> 1), k-chunk h of the next layer's operand.  Ranges 0 and 1
    // (drain of chunk 0) are still being read by this layer's chunk 1: their stores wait for R_FREE(R); ranges 2 and
    // 3 are written by the layer's last drain, when nothing reads the old ones any more.
    // With EW = 16 four warps share a lane quarter and take 32 columns each (hq = 0..3) instead of 64 (hq = h = 0, 1).
    constexpr int NQ = EW / 4;        // warps per TMEM lane quarter
    constexpr int CW = 128 / NQ;      // accumulator columns per warp
    constexpr int NG = CW / 8;        // 8-column (16-byte) groups per warp
    constexpr int NEPI_T = 32 * EW;   // epilogue threads
    auto epi_bar = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(NEPI_T) : "memory"); };
    const int q = warp & 3;
    const int hq = (warp - 2) >> 2;            // which CW columns of the quarter's 128
    const int h = (hq * CW) >> 6;              // k-chunk of the next layer's K range
    const int g0 = ((hq * CW) & 63) >> 3;      // first 16-byte group inside that k-chunk's 128-byte rows
    const int half = q >> 1;
    const int prow = (q & 1) * 32 + lane;  // point of the tile, 0..63
    const int et = threadIdx.x - 64;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t row_off = (uint32_t)(prow * 128);
    const int sw = prow & 7;
    const float slope = P.head.slope;
    const uint64_t slope2 = ptx::pack_f32x2(slope, slope);
    const uint32_t lead_aready0 = ptx::mapa(A_READY(0, 0), 0);
    const uint32_t lead_accempty0 = ptx::mapa(ACC_EMPTY(0), 0);
    const uint32_t R_FREE = A_READY(1, half);
    uint32_t acc_it = 0, rf_ph = 0;
    int bias_buf = 0;

    for (int64_t it = 0; it < iters; it++) {
      const int64_t tile = it * gridDim.x + blockIdx.x;
      const int64_t g = tile * TILE_PTS + prow;
      float x = 0.f, y = 0.f, z = 0.f;
      if (g < P.n) {
        if (P.grid_mode) grid_point(P, P.r_lo + g, x, y, z);
        else { x = __ldg(P.xyz + 3 * g); y = __ldg(P.xyz + 3 * g + 1); z = __ldg(P.xyz + 3 * g + 2); }
      }
      float yc[5] = {0, 0, 0, 0, 0}, yt[5] = {0, 0, 0, 0, 0};

      for (int net = 0; net < 2; net++) {
        if (!((P.net_mask >> net) & 1)) continue;
        const Net& N = P.nets[net];
        // ---- layer 0: K = 16 operand [p_hi | p_lo | p_hi | 0] in the first 32 B of the point's a_hi row of K range 0
        // (buffer 0); the tiles hold [W_hi | W_hi | W_lo]
        epi_bar();  // everybody is past the previous use of sBias
        {
          const float4* t = P.tables + N.l0_table * 512;
          for (int i = et; i < 512; i += NEPI_T) sBias[bias_buf * 512 + i] = __ldg(&t[i].w);
        }
        if (half == 0) {
          if (hq == 0) {
            float hi[3], lo[3];
            const float p3[3] = {x, y, z};
#pragma unroll
            for (int i = 0; i < 3; i++) {
              hi[i] = ptx::round_op16<1>(p3[i]);
              lo[i] = p3[i] - hi[i];
            }
            const float k[16] = {hi[0], hi[1], hi[2], lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; i++) w[i] = ptx::pack_op16x2<1>(k[2 * i], k[2 * i + 1]);
            const uint32_t my_row = sA + row_off;
            tc::st_shared_v4(my_row + ((0 ^ sw) << 4), w[0], w[1], w[2], w[3]);
            tc::st_shared_v4(my_row + ((1 ^ sw) << 4), w[4], w[5], w[6], w[7]);
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(lead_aready0);  // K range 0 is owned by the quarters 0, 1
        }

        for (int l = 0; l < N.n_layers; l++) {
          const Layer L = N.layers[l];
          const float inv = L.inv_scale;
          if (L.kind == KIND_HEAD) {
            const uint32_t buf = acc_it & 1u;
            { const long long t = tick(); ptx::mbar_wait(ACC_FULL(buf), (acc_it >> 1) & 1u, P.err, 0x500 | l); w3 += tick() - t; }
            acc_it++;
            ptx::tc_fence_after();
            uint32_t r[8], rs[8];
            if (half == 0 && hq == 0) {
              ptx::tmem_ld8(tmem_lane + buf * 256u, r);
              ptx::tmem_ld8(tmem_lane + buf * 256u + 128u, rs);
              ptx::tmem_wait_ld();
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(lead_accempty0 + 8u * buf);
            if (half == 0 && hq == 0) {
              float* dst = net == 0 ? yc : yt;
#pragma unroll
              for (int i = 0; i < 5; i++) {
                const float v = fmaf(__uint_as_float(r[i]) + __uint_as_float(rs[i]), inv, __ldg(P.bias + L.bias_off + i));
                dst[i] = N.head_leaky ? tc::leaky(v, slope) : v;
              }
            }
            continue;
          }
          epi_bar();
          const int next_kind = (l + 1 < N.n_layers) ? N.layers[l + 1].kind : -1;
          const int next_kr = (l + 1 < N.n_layers) ? N.layers[l + 1].k_ranges : 0;
          constexpr int PER = 512 / NEPI_T;  // table entries per thread: 2 or 1
          float pre_b[PER];
          float4 pre_t[PER];
          if (next_kind == KIND_HIDDEN) {
#pragma unroll
            for (int i = 0; i < PER; i++) pre_b[i] = __ldg(P.bias + N.layers[l + 1].bias_off + et + i * NEPI_T);
          } else if (next_kind == KIND_HIDDEN_XYZ) {
            const float4* t = P.tables + N.layers[l + 1].table * 512;
#pragma unroll
            for (int i = 0; i < PER; i++) pre_t[i] = __ldg(t + et + i * NEPI_T);
          }
          const uint32_t bias_s = base + SM_BIAS + (uint32_t)(bias_buf * 2048);
          const uint64_t inv2 = ptx::pack_f32x2(inv, inv);

          auto drain = [&](auto kind_c, int j) {
            constexpr int KIND = decltype(kind_c)::value;
            const uint32_t buf = acc_it & 1u;
            const int col0 = j * X3_NCHUNK + half * 128 + hq * CW;
            // the first half of this warp's bias values is fetched while the accumulator is still being computed: the
            // tensor core's operand reads keep the shared-memory pipe busy, a load issued inside the drain waits long
            uint32_t bpre[CW / 2];
            if (KIND != KIND_HIDDEN_XYZ) {
#pragma unroll
              for (int i = 0; i < CW / 8; i++)
                ptx::ld_shared_v4u(bias_s + (uint32_t)((col0 + 4 * i) * 4), bpre[4 * i], bpre[4 * i + 1], bpre[4 * i + 2], bpre[4 * i + 3]);
            }
            { const long long t = tick(); ptx::mbar_wait(ACC_FULL(buf), (acc_it >> 1) & 1u, P.err, 0x600 | j); w0 += tick() - t; }
            acc_it++;
            ptx::tc_fence_after();
            uint32_t r[CW];
            ptx::tmem_ld_n<CW>(tmem_lane + buf * 256u + hq * CW, r);
            if (L.k_steps != 1) {  // + the accumulator of the small products (layer 0 is a single MMA); in two halves:
              // CW + CW live accumulator registers do not fit next to the rest
#pragma unroll
              for (int hh = 0; hh < 2; hh++) {
                uint32_t rs[CW / 2];
                ptx::tmem_ld_n<CW / 2>(tmem_lane + buf * 256u + 128u + hq * CW + hh * (CW / 2), rs);
                ptx::tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < CW / 2; i += 2) {
                  const uint64_t t = ptx::add_f32x2((uint64_t)r[hh * (CW / 2) + i] | ((uint64_t)r[hh * (CW / 2) + i + 1] << 32),
                                                    (uint64_t)rs[i] | ((uint64_t)rs[i + 1] << 32));
                  r[hh * (CW / 2) + i] = (uint32_t)t;
                  r[hh * (CW / 2) + i + 1] = (uint32_t)(t >> 32);
                }
              }
            } else {
              ptx::tmem_wait_ld();
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(lead_accempty0 + 8u * buf);  // accumulator buffer is free again
            const int R = 2 * j + half;  // K range of the next layer this warp produces (k-chunk h of it)
            const bool wait_free = j == 0 && L.n_chunks > 1;
            if (R >= next_kr) {  // padding columns: the next layer has no such K range
              if (wait_free) rf_ph ^= 1u;
              return;
            }
            const uint32_t dst_hi = sA + (uint32_t)(R * X3_RANGE + h * X3_TILE) + row_off;
            uint32_t ph[CW / 2], pl[CW / 2];
#pragma unroll
            for (int gq = 0; gq < NG; gq++) {
              float v[8];
              if (KIND == KIND_HIDDEN_XYZ) {
#pragma unroll
                for (int f = 0; f < 4; f++) {
                  const float4 t0 = sTab[col0 + gq * 8 + 2 * f], t1 = sTab[col0 + gq * 8 + 2 * f + 1];
                  const float s0 = fmaf(t0.x, x, fmaf(t0.y, y, fmaf(t0.z, z, t0.w)));
                  const float s1 = fmaf(t1.x, x, fmaf(t1.y, y, fmaf(t1.z, z, t1.w)));
                  v[2 * f] = tc::leaky(fmaf(__uint_as_float(r[gq * 8 + 2 * f]), inv, s0), slope);
                  v[2 * f + 1] = tc::leaky(fmaf(__uint_as_float(r[gq * 8 + 2 * f + 1]), inv, s1), slope);
                }
              } else {
                uint32_t bw[8];
                if (gq < NG / 2) {
#pragma unroll
                  for (int i = 0; i < 8; i++) bw[i] = bpre[gq * 8 + i];
                } else {  // 128-bit broadcast loads
                  ptx::ld_shared_v4u(bias_s + (uint32_t)((col0 + gq * 8) * 4), bw[0], bw[1], bw[2], bw[3]);
                  ptx::ld_shared_v4u(bias_s + (uint32_t)((col0 + gq * 8 + 4) * 4), bw[4], bw[5], bw[6], bw[7]);
                }
#pragma unroll
                for (int f = 0; f < 4; f++) {
                  const uint64_t acc = (uint64_t)r[gq * 8 + 2 * f] | ((uint64_t)r[gq * 8 + 2 * f + 1] << 32);
                  const uint64_t t = ptx::fma_f32x2(acc, inv2, (uint64_t)bw[2 * f] | ((uint64_t)bw[2 * f + 1] << 32));
                  const uint64_t u = ptx::mul_f32x2(t, slope2);
                  v[2 * f] = fmaxf(ptx::f32x2_lo(t), ptx::f32x2_lo(u));
                  v[2 * f + 1] = fmaxf(ptx::f32x2_hi(t), ptx::f32x2_hi(u));
                }
              }
              if (DBG && (debug & 64)) {  // per-layer checksum of the activations (tools/x3_diag.py)
                float cs = 0.f, cw = 0.f;
#pragma unroll
                for (int f = 0; f < 8; f++) { cs += v[f]; cw += v[f] * (float)(col0 + gq * 8 + f + 1); }
                atomicAdd(reinterpret_cast<double*>(P.prof) + net * 16 + l, (double)cs);
                atomicAdd(reinterpret_cast<double*>(P.prof) + 32 + net * 16 + l, (double)cw);
              }
              // a_hi = the value with its mantissa cut to 10 bits (exactly an IEEE half for |v| >= 2^-14), a_lo = the
              // exact remainder rounded to half: 21 significant bits, one mask + one subtract per value
#pragma unroll
              for (int f = 0; f < 4; f++) {
                const float h0 = __uint_as_float(__float_as_uint(v[2 * f]) & 0xFFFFE000u);
                const float h1 = __uint_as_float(__float_as_uint(v[2 * f + 1]) & 0xFFFFE000u);
                ph[gq * 4 + f] = ptx::pack_f16x2(h0, h1);
                pl[gq * 4 + f] = ptx::pack_f16x2(v[2 * f] - h0, v[2 * f + 1] - h1);
              }
            }
            if (wait_free) {  // the MMAs of chunk 1 that still read the old K range R
              { const long long t = tick(); ptx::mbar_wait(R_FREE, rf_ph, P.err, 0x680 | l); w1 += tick() - t; }
              rf_ph ^= 1u;
            }
#pragma unroll
            for (int gq = 0; gq < NG; gq++) {
              const uint32_t d = dst_hi + (((g0 + gq) ^ sw) << 4);
              tc::st_shared_v4(d, ph[gq * 4], ph[gq * 4 + 1], ph[gq * 4 + 2], ph[gq * 4 + 3]);
              tc::st_shared_v4(d + X3_PART, pl[gq * 4], pl[gq * 4 + 1], pl[gq * 4 + 2], pl[gq * 4 + 3]);
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(lead_aready0 + 8u * R);
          };
          using I0 = std::integral_constant<int, 0>;
          using I1 = std::integral_constant<int, 1>;
          if (L.kind == KIND_HIDDEN_XYZ) for (int j = 0; j < L.n_chunks; j++) drain(I1{}, j);
          else for (int j = 0; j < L.n_chunks; j++) drain(I0{}, j);
          if (next_kind == KIND_HIDDEN) {
            bias_buf ^= 1;
#pragma unroll
            for (int i = 0; i < PER; i++) sBias[bias_buf * 512 + et + i * NEPI_T] = pre_b[i];
          } else if (next_kind == KIND_HIDDEN_XYZ) {
#pragma unroll
            for (int i = 0; i < PER; i++) sTab[et + i * NEPI_T] = pre_t[i];
          }
        }
      }
      if (half == 0 && hq == 0 && g < P.n) write_heads(P.head, yc, yt, P.out, g);
    }
    // counters of warp 2 (lanes [64,96): waits R_FREE(1)) and of warp 4 (lanes [0,32): waits R_FREE(0))
    if (prof && warp == 2 && lane == 0) { pc[4] = tick() - t_begin; pc[5] = w0; pc[6] = w1; pc[8] = w3; }
    if (prof && warp == 4 && lane == 0) { pc[7] = w1; pc[13] = w0; }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();  // nobody exits (or frees tensor memory) while the peer may still use its barriers / TMEM
  if (warp == 1) ptx::tmem_dealloc2(tmem_base, 512);
}

}  // namespace tc2
}  // namespace dj
