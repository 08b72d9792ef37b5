// Constants, layer-table structs and small device helpers shared by the tensor-core kernels
// (mlp_tc2.cuh: fused decoder forward; latent_tc2.cuh: fused forward + loss + backward): tile geometry, the
// shared-memory carve-up, the per-layer table the host builder fills, np.linspace grid axes, LeakyReLU, the
// epilogue warps' named barrier.  (The single-CTA cta_group::1 kernel this file used to hold was the predecessor
// of the CTA-pair kernel; it is gone.)
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
  const uint8_t* stream;  // unused by the pair program (its streams live in tc2::Net)
  TcLayer layers[MAX_LAYERS];
};
__device__ __forceinline__ float lin_axis(int t, int d, double step, double stop, double offs) {
  // np.linspace(0, stop, d)[t] - offs in float64, no fused multiply-add, then f32
  double v = (d > 1 && t == d - 1) ? stop : __dmul_rn((double)t, step);  // np.linspace(0, stop, 1) is [0.0]
  return (float)__dsub_rn(v, offs);
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

}  // namespace tc
}  // namespace dj
