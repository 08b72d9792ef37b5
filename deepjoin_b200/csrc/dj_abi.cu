// C ABI of deepjoin_b200 (see include/deepjoin_b200.h): weight packing, launch orchestration.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <thread>
#include <cerrno>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include "common.cuh"
#include "mlp_simt.cuh"
#include "mlp_tc.cuh"
#include "mlp_tc2.cuh"
#include "latent_tc2.cuh"
#include "mc.cuh"
#include "sampler.cuh"
#include "mesh_merge.cuh"
#include "nn.cuh"
#ifdef DJ_DEBUG
#include "debug_probe.cuh"
#endif

using namespace dj;

constexpr int DJ_MAX_BATCH = 64;  // shapes per launch of the batched latent step

// ====================================================================== pack
// what one call (one stream) writes: per-shape constants of up to DJ_MAX_BATCH codes, the device error word of the
// watchdogs, grow-only workspaces
struct DjCtx {
  float4* tables = nullptr;
  unsigned int* err = nullptr;
  DevBuf ws_fwd, ws_lat, ws_host, ws_mask;
  PinBuf pin_host;  // staging of dj_decode_samples_host
  std::vector<const float*> lat_ptrs;     // the sample-array pointers last uploaded for the latent step ...
  const float** lat_ptrs_dev = nullptr;   // ... and where they went (a block-wise loop uploads them once)
  ~DjCtx() {
    if (tables) cudaFree(tables);
    if (err) cudaFree(err);
    ws_fwd.release(); ws_lat.release(); ws_host.release(); ws_mask.release();
    pin_host.release();
  }
};

struct DjPack {
  DjNetDesc desc{};
  int device = 0;
  int num_sms = 0;
  int L = 0, Lb = 0, H = 512;
  int K3 = 0;  // width of the h part of the re-injection layer's input (hidden - (L+3))
  // fp32 folded weights on the device, nn.Linear layout [out,in]
  std::vector<float*> fW, fB, hW, hB;
  std::vector<int> f_out, f_in, h_out, h_in;
  float* fWx = nullptr;  // [H, K3+3] = [W_li[:, :K3] | W_li[:, K3+L : K3+L+3]] (fp32 forward of the re-injection layer)
  // tensor-core path
  float* bias_dev = nullptr;
  tc::TcNet nets[2];
  // CTA-pair (cta_group::2) program: one weight stream per cluster rank
  uint8_t* stream2[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  uint8_t* stream2h[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // the same stages with IEEE-half weights (DJ_PREC_FP16)
  tc2::Net nets2[2];
  // split-operand program (DJ_PREC_FP16X2): every K = 128 stage twice (W_hi, W_lo; IEEE half, scaled per layer)
  uint8_t* stream2x[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  tc2::Net nets2x[2];
  uint8_t* stream3x[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // three-pass split-operand program
  tc2::Net nets3x[2];
  bool x3_ok = true;  // false: a layer shape the three-pass program does not cover; DJ_PREC_FP16X2 then runs the four-pass one
  unsigned long long* prof = nullptr;  // [num_sms][16] debug cycle counters (dj_debug_read_prof)
  // fused forward+backward program of the latent step (latent_tc2.cuh)
  uint8_t* lat_stream[2] = {nullptr, nullptr};
  uint8_t* lat_stream_h[2] = {nullptr, nullptr};  // forward tiles in IEEE half
  lat2::Layer lat_prog[lat2::MAX_PROG];
  lat2::Layer lat_prog_x[lat2::MAX_PROG];   // split-operand program (DJ_PREC_FP16X2): every stage as W_lo and W_hi
  uint8_t* lat_stream_x[2] = {nullptr, nullptr};
  int lat_n_prog = 0;
  int lat_net1_lo = 0, lat_net1_hi = 0;  // program entries of the second net (forward + backward)
  bool lat_attr_set = false;
  simt::FoldParams fold{};
  bool tc_attr_set2 = false;
  // Everything a call writes lives in a per-stream context (re-entrant across streams, no shared mutable
  // state): calls on one stream are ordered by the stream and share a context; calls on different streams (or host
  // threads driving different streams) get different ones and may overlap.  The pack itself is read-only after
  // dj_pack_create.
  std::mutex mu;
  std::map<cudaStream_t, std::shared_ptr<DjCtx>> ctxs;
  // host-buffer entry points (dj_decode_samples_host*): two private streams sharing one context; concurrent host
  // calls on one pack are serialised by host_mu
  std::mutex host_mu;
  cudaStream_t host_streams[2] = {nullptr, nullptr};
  cudaEvent_t host_events[2] = {nullptr, nullptr};
  DjCtx* cx(cudaStream_t st);
};

DjCtx* DjPack::cx(cudaStream_t st) {
  std::lock_guard<std::mutex> lock(mu);
  auto it = ctxs.find(st);
  if (it != ctxs.end()) return it->second.get();
  auto c = std::make_shared<DjCtx>();
  if (cudaMalloc((void**)&c->tables, (size_t)DJ_MAX_BATCH * 3 * 512 * sizeof(float4)) != cudaSuccess) return nullptr;
  if (cudaMalloc((void**)&c->err, sizeof(unsigned int)) != cudaSuccess || cudaMemset(c->err, 0, sizeof(unsigned int)) != cudaSuccess) {
    cudaFree(c->tables);
    c->tables = nullptr;
    return nullptr;
  }
  ctxs[st] = c;
  return c.get();
}
#define DJ_CTX(var, pack, st)                                                         \
  DjCtx* var = (pack)->cx(st);                                                        \
  if (!var) return dj::fail(DJ_ERR_CUDA, "%s:%d per-stream context allocation failed", __FILE__, __LINE__)

static void free_pack(DjPack* p) {
  if (!p) return;
  DeviceGuard g(p->device);
  for (auto v : {&p->fW, &p->fB, &p->hW, &p->hB})
    for (float* q : *v)
      if (q) cudaFree(q);
  if (p->fWx) cudaFree(p->fWx);
  for (int i = 0; i < 2; i++) {
    for (int r = 0; r < 2; r++)
      if (p->stream2[i][r]) cudaFree(p->stream2[i][r]);
    for (int r = 0; r < 2; r++)
      if (p->stream2h[i][r]) cudaFree(p->stream2h[i][r]);
    for (int r = 0; r < 2; r++)
      if (p->stream2x[i][r]) cudaFree(p->stream2x[i][r]);
    for (int r = 0; r < 2; r++)
      if (p->stream3x[i][r]) cudaFree(p->stream3x[i][r]);
  }
  if (p->bias_dev) cudaFree(p->bias_dev);
  for (int r = 0; r < 2; r++) {
    if (p->lat_stream[r]) cudaFree(p->lat_stream[r]);
    if (p->lat_stream_h[r]) cudaFree(p->lat_stream_h[r]);
    if (p->lat_stream_x[r]) cudaFree(p->lat_stream_x[r]);
  }
  if (p->prof) cudaFree(p->prof);
  for (int i = 0; i < 2; i++) {
    if (p->host_streams[i]) cudaStreamDestroy(p->host_streams[i]);
    if (p->host_events[i]) cudaEventDestroy(p->host_events[i]);
  }
  p->ctxs.clear();
  delete p;
}

// one [rows x 64] bf16 tile, K-major, 128-byte swizzle (the layout the UMMA descriptor
// umma_desc_sw128_kmajor describes): element (r, k) at r*128 + ((k/8) ^ (r%8))*16 + (k%8)*2
static void pack_tile(const float* W, int ldw, int n_real, int k_real, int n0, int k0, int rows, uint8_t* dst, bool f16 = false) {
  for (int r = 0; r < rows; r++)
    for (int k = 0; k < 64; k++) {
      const int gn = n0 + r, gk = k0 + k;
      const float v = (gn < n_real && gk < k_real) ? W[(size_t)gn * ldw + gk] : 0.f;
      const size_t off = (size_t)r * 128 + (size_t)(((k >> 3) ^ (r & 7)) << 4) + (size_t)(k & 7) * 2;
      if (f16) {
        const __half b = __float2half_rn(v);
        memcpy(dst + off, &b, 2);
      } else {
        const __nv_bfloat16 b = __float2bfloat16_rn(v);
        memcpy(dst + off, &b, 2);
      }
    }
}


// layer-0 tile: row n (hidden unit), K = 16 slots [W_xyz(3) | W_xyz(3) | lo(W_xyz)(3) | 0] in the operand type,
// against the point operand [p_hi | p_lo | p_hi | 0] (mlp_tc2.cuh): sum = W_hi (p_hi + p_lo) + W_lo p_hi
static void pack_tile_l0(const float* W, int ldw, int xyz_off, int n0, int rows, uint8_t* dst, bool f16) {
  for (int r = 0; r < rows; r++)
    for (int k = 0; k < 64; k++) {
      float v = 0.f;
      if (k < 9) {
        const float w = W[(size_t)(n0 + r) * ldw + xyz_off + k % 3];
        const float hi = f16 ? __half2float(__float2half_rn(w)) : __bfloat162float(__float2bfloat16_rn(w));
        v = (k < 6) ? hi : (w - hi);
      }
      const size_t off = (size_t)r * 128 + (size_t)(((k >> 3) ^ (r & 7)) << 4) + (size_t)(k & 7) * 2;
      if (f16) {
        const __half b = __float2half_rn(v);
        memcpy(dst + off, &b, 2);
      } else {
        const __nv_bfloat16 b = __float2bfloat16_rn(v);
        memcpy(dst + off, &b, 2);
      }
    }
}

// split-operand tiles (DJ_PREC_FP16X2): the weight times `scale` (a power of two) as IEEE half, part 0 = the rounded
// value, part 1 = the rounding remainder (again rounded to half): hi + lo carries 22 significant bits
static inline float half_part(float w, int part) {
  const float hi = __half2float(__float2half_rn(w));
  return part == 0 ? hi : __half2float(__float2half_rn(w - hi));
}
static void pack_tile_x2(const float* W, int ldw, int n_real, int k_real, int n0, int k0, int rows, uint8_t* dst, float scale, int part) {
  for (int r = 0; r < rows; r++)
    for (int k = 0; k < 64; k++) {
      const int gn = n0 + r, gk = k0 + k;
      const float v = (gn < n_real && gk < k_real) ? W[(size_t)gn * ldw + gk] * scale : 0.f;
      const size_t off = (size_t)r * 128 + (size_t)(((k >> 3) ^ (r & 7)) << 4) + (size_t)(k & 7) * 2;
      const __half b = __float2half_rn(half_part(v, part));
      memcpy(dst + off, &b, 2);
    }
}
// layer-0 tile of the split-operand program: [W_hi | W_hi | W_lo | 0] of the scaled xyz columns
static void pack_tile_l0_x2(const float* W, int ldw, int xyz_off, int n0, int rows, uint8_t* dst, float scale) {
  for (int r = 0; r < rows; r++)
    for (int k = 0; k < 64; k++) {
      float v = 0.f;
      if (k < 9) v = half_part(W[(size_t)(n0 + r) * ldw + xyz_off + k % 3] * scale, k < 6 ? 0 : 1);
      const size_t off = (size_t)r * 128 + (size_t)(((k >> 3) ^ (r & 7)) << 4) + (size_t)(k & 7) * 2;
      const __half b = __float2half_rn(v);
      memcpy(dst + off, &b, 2);
    }
}
// power of two that brings the largest |w| of a layer into [256, 512): the low halves of all but vanishing weights
// then stay normal numbers (IEEE half: 2^-14), and products with activations up to ~1e2 stay far inside fp32
static float x2_scale(const float* W, size_t n) {
  float m = 0.f;
  for (size_t i = 0; i < n; i++) m = std::max(m, fabsf(W[i]));
  if (!(m > 0.f) || !std::isfinite(m)) return 1.f;
  int e = 0;
  frexpf(m, &e);  // m = f * 2^e, f in [0.5, 1)
  return ldexpf(1.f, std::max(-24, std::min(24, 9 - e)));
}

// transposed source: tile row n, column k <- W[k0 + k][n0 + n]  (B operand of a backward GEMM, dX = dY . W)
static void pack_tile_T(const float* W, int ldw, int out_real, int in_real, int n0, int k0, int rows, uint8_t* dst) {
  for (int r = 0; r < rows; r++)
    for (int k = 0; k < 64; k++) {
      const int gn = n0 + r, gk = k0 + k;  // gn: input index of the layer (output column of the backward GEMM), gk: its output index
      const float v = (gn < in_real && gk < out_real) ? W[(size_t)gk * ldw + gn] : 0.f;
      const __nv_bfloat16 b = __float2bfloat16_rn(v);
      const size_t off = (size_t)r * 128 + (size_t)(((k >> 3) ^ (r & 7)) << 4) + (size_t)(k & 7) * 2;
      memcpy(dst + off, &b, 2);
    }
}
// the same, split-operand program: bf16 high part (part 0) or the bf16 rounding remainder (part 1) of the weight
static void pack_tile_T_x2(const float* W, int ldw, int out_real, int in_real, int n0, int k0, int rows, uint8_t* dst, int part) {
  for (int r = 0; r < rows; r++)
    for (int k = 0; k < 64; k++) {
      const int gn = n0 + r, gk = k0 + k;
      const float w = (gn < in_real && gk < out_real) ? W[(size_t)gk * ldw + gn] : 0.f;
      const float hi = __bfloat162float(__float2bfloat16_rn(w));
      const __nv_bfloat16 b = __float2bfloat16_rn(part == 0 ? hi : w - hi);
      const size_t off = (size_t)r * 128 + (size_t)(((k >> 3) ^ (r & 7)) << 4) + (size_t)(k & 7) * 2;
      memcpy(dst + off, &b, 2);
    }
}
// head-gradient tile: row n (hidden unit), K = 16 slots [W_head[0..4][n] | W_head[0..4][n] | lo(W_head[0..4][n]) | 0]
// against the operand [g_hi | g_lo | g_hi | 0] (latent_tc2.cuh:write_head_operand)
static void pack_tile_head_T(const float* Wh, int ldw, int n0, int rows, uint8_t* dst) {
  for (int r = 0; r < rows; r++)
    for (int k = 0; k < 64; k++) {
      float v = 0.f;
      if (k < 15) {
        const float w = Wh[(size_t)(k % 5) * ldw + n0 + r];
        const float hi = __bfloat162float(__float2bfloat16_rn(w));
        v = (k < 10) ? hi : (w - hi);
      }
      const __nv_bfloat16 b = __float2bfloat16_rn(v);
      const size_t off = (size_t)r * 128 + (size_t)(((k >> 3) ^ (r & 7)) << 4) + (size_t)(k & 7) * 2;
      memcpy(dst + off, &b, 2);
    }
}

extern "C" int dj_abi_version(void) { return DJ_ABI_VERSION; }
extern "C" const char* dj_last_error(void) { return last_error_ref().c_str(); }
extern "C" int64_t dj_launch_count(int reset) {
  return reset ? launch_counter().exchange(0) : launch_counter().load();
}
extern "C" int dj_pack_device(const DjPack* p) { return p ? p->device : -1; }

extern "C" int dj_pack_create(const DjNetDesc* d, const float* const* f_w, const float* const* f_b,
                              const float* const* h_w, const float* const* h_b, int device, DjPack** out) {
  if (!d || !f_w || !f_b || !h_w || !h_b || !out) return fail(DJ_ERR_INVALID, "dj_pack_create: null argument");
  *out = nullptr;
  if (d->hidden != 512) return fail(DJ_ERR_UNSUPPORTED, "hidden width %d (only 512 is supported)", d->hidden);
  if (d->f_layers < 4 || d->f_layers > tc::MAX_LAYERS || d->h_layers < 3 || d->h_layers > tc::MAX_LAYERS)
    return fail(DJ_ERR_UNSUPPORTED, "layer counts f=%d h=%d outside the supported range", d->f_layers, d->h_layers);
  if (d->latent_in < 2 || d->latent_in > d->f_layers - 2)
    return fail(DJ_ERR_UNSUPPORTED, "latent_in=%d must name a hidden layer >= 2", d->latent_in);
  const int L = d->latent_size, Lb = d->tool_latent_size, H = d->hidden;
  const int K3 = H - (L + 3);
  if (L < 1 || Lb < 1 || K3 < 64) return fail(DJ_ERR_UNSUPPORTED, "latent sizes %d/%d not supported", L, Lb);
  int ndev = 0;
  DJ_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(DJ_ERR_INVALID, "device %d out of range", device);
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  DJ_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(DJ_ERR_UNSUPPORTED, "device sm_%d%d: this library is built for sm_100a only", prop.major, prop.minor);

  DjPack* p = new DjPack();
  p->desc = *d;
  p->device = device;
  p->num_sms = prop.multiProcessorCount;
  p->L = L; p->Lb = Lb; p->H = H; p->K3 = K3;
  auto bail = [&](int code) { free_pack(p); return code; };

  // ---- layer shapes as Decoder.__init__ builds them (decoder :60-131)
  for (int l = 0; l < d->f_layers; l++) {
    p->f_in.push_back(l == 0 ? L + 3 : H);
    p->f_out.push_back(l == d->f_layers - 1 ? 5 : (l + 1 == d->latent_in ? K3 : H));
  }
  for (int l = 0; l < d->h_layers; l++) {
    p->h_in.push_back(l == 0 ? Lb + 3 : H);
    p->h_out.push_back(l == d->h_layers - 1 ? 5 : H);
  }
  auto upload = [&](const float* src, size_t n, float** dst) -> cudaError_t {
    cudaError_t e = cudaMalloc((void**)dst, n * sizeof(float));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice);
  };
#define DJ_TRY(expr)                                                                                   \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) return bail(fail(DJ_ERR_CUDA, "%s -> %s", #expr, cudaGetErrorString(_e)));  \
  } while (0)
  p->fW.assign(d->f_layers, nullptr); p->fB.assign(d->f_layers, nullptr);
  p->hW.assign(d->h_layers, nullptr); p->hB.assign(d->h_layers, nullptr);
  for (int l = 0; l < d->f_layers; l++) {
    DJ_TRY(upload(f_w[l], (size_t)p->f_out[l] * p->f_in[l], &p->fW[l]));
    DJ_TRY(upload(f_b[l], (size_t)p->f_out[l], &p->fB[l]));
  }
  for (int l = 0; l < d->h_layers; l++) {
    DJ_TRY(upload(h_w[l], (size_t)p->h_out[l] * p->h_in[l], &p->hW[l]));
    DJ_TRY(upload(h_b[l], (size_t)p->h_out[l], &p->hB[l]));
  }
  const int li = d->latent_in;
  {  // fp32 re-injection layer with the latent columns removed: [h (K3) | xyz (3)]
    std::vector<float> wx((size_t)H * (K3 + 3));
    for (int j = 0; j < H; j++) {
      for (int k = 0; k < K3; k++) wx[(size_t)j * (K3 + 3) + k] = f_w[li][(size_t)j * H + k];
      for (int k = 0; k < 3; k++) wx[(size_t)j * (K3 + 3) + K3 + k] = f_w[li][(size_t)j * H + K3 + L + k];
    }
    DJ_TRY(upload(wx.data(), wx.size(), &p->fWx));
  }
  // ---- latent fold descriptors (tables 0: lin0, 1: lin<latent_in>, 2: hnet_lin0)
  p->fold.src[0] = {p->fW[0], p->fB[0], L + 3, 0, L, L, 0};
  p->fold.src[1] = {p->fW[li], p->fB[li], H, K3, L, K3 + L, 0};
  p->fold.src[2] = {p->hW[0], p->hB[0], Lb + 3, 0, Lb, Lb, L};
  DJ_TRY(cudaMalloc((void**)&p->prof, (size_t)prop.multiProcessorCount * 16 * sizeof(unsigned long long)));
  DJ_TRY(cudaMemset(p->prof, 0, (size_t)prop.multiProcessorCount * 16 * sizeof(unsigned long long)));

  // ---- tensor-core program + weight stream
  std::vector<float> bias_host;
  for (int net = 0; net < 2; net++) {
    const int nl = net == 0 ? d->f_layers : d->h_layers;
    const float* const* B = net == 0 ? f_b : h_b;
    const std::vector<int>& outs = net == 0 ? p->f_out : p->h_out;
    const std::vector<int>& ins = net == 0 ? p->f_in : p->h_in;
    tc::TcNet& N = p->nets[net];
    memset(&N, 0, sizeof N);
    N.n_layers = nl - 1;  // layer 0 runs on CUDA cores
    N.l0_table = net == 0 ? 0 : 2;
    N.head_leaky = net == 1;  // decoder :259-270
    int prev_n_chunks = 4;  // layer 0 writes all 8 k-chunks
    for (int l = 1; l < nl; l++) {
      tc::TcLayer& T = N.layers[l - 1];
      const bool head = (l == nl - 1);
      const bool reinj = (net == 0 && l == li);
      const int n_real = outs[l];
      const int k_real = reinj ? K3 : ins[l];
      T.k_chunks = (int16_t)(2 * prev_n_chunks);
      if (k_real > T.k_chunks * 64) return bail(fail(DJ_ERR_UNSUPPORTED, "layer %d: K=%d exceeds the previous layer's width", l, k_real));
      T.n_chunks = (int16_t)(head ? 1 : (n_real + 127) / 128);
      T.mma_n = (int16_t)(head ? 16 : 128);
      T.kind = (int16_t)(head ? tc::KIND_HEAD : (reinj ? tc::KIND_HIDDEN_XYZ : tc::KIND_HIDDEN));
      T.table = reinj ? 1 : 0;
      T.stage_bytes = T.mma_n * 64 * 2;
      T.bias_off = (int32_t)bias_host.size();
      if (!reinj) {
        for (int j = 0; j < 512; j++) bias_host.push_back(j < n_real ? B[l][j] : 0.f);
      }
      prev_n_chunks = T.n_chunks;
    }
  }
  // ---- CTA-pair program: layer 0 as a K = 16 tensor-core layer (point split hi/lo), then the same layers / bias
  // offsets; every 128-row output chunk is split 64/64 between the two CTAs of a pair and a ring stage carries
  // K = 128 (two [rows x 64] tiles)
  for (int net = 0; net < 2; net++) {
    const int nl = net == 0 ? d->f_layers : d->h_layers;
    const float* const* W = net == 0 ? f_w : h_w;
    const std::vector<int>& outs = net == 0 ? p->f_out : p->h_out;
    const std::vector<int>& ins = net == 0 ? p->f_in : p->h_in;
    const tc::TcNet& N1 = p->nets[net];
    tc2::Net& N = p->nets2[net];
    memset(&N, 0, sizeof N);
    N.n_layers = N1.n_layers + 1;
    N.l0_table = N1.l0_table;
    N.head_leaky = N1.head_leaky;
    if (N.n_layers > tc2::MAX_LAYERS) return bail(fail(DJ_ERR_UNSUPPORTED, "too many layers for the CTA-pair program"));
    std::vector<uint8_t> blob[2], blobh[2];
    {
      tc2::Layer& T = N.layers[0];
      T.n_chunks = 4; T.k_stages = 1; T.k_steps = 1; T.mma_n = 128; T.kind = tc2::KIND_L0TC; T.inv_scale = 1.f; T.k_ranges = 1;
      T.korder[0] = T.korder[1] = 0x76543210u;
      T.tile_bytes = 64 * 128; T.stage_bytes = T.tile_bytes;
      const int xyz_off = (net == 0 ? L : Lb);
      for (int j = 0; j < 4; j++)
        for (int r = 0; r < 2; r++) {
          const size_t o = blob[r].size();
          blob[r].resize(o + T.tile_bytes);
          blobh[r].resize(o + T.tile_bytes);
          pack_tile_l0(W[0], ins[0], xyz_off, j * 128 + r * 64, 64, blob[r].data() + o, false);
          pack_tile_l0(W[0], ins[0], xyz_off, j * 128 + r * 64, 64, blobh[r].data() + o, true);
        }
    }
    for (int l = 1; l < nl; l++) {
      const tc::TcLayer& T1 = N1.layers[l - 1];
      tc2::Layer& T = N.layers[l];
      const bool head = (l == nl - 1);
      const bool reinj = (net == 0 && l == li);
      const int n_real = outs[l];
      const int k_real = reinj ? K3 : ins[l];
      if (T1.k_chunks % 2) return bail(fail(DJ_ERR_UNSUPPORTED, "layer %d: K chunks must be even", l));
      T.n_chunks = T1.n_chunks;
      T.k_stages = (int16_t)(T1.k_chunks / 2);
      T.k_ranges = T.k_stages;
      T.korder[0] = T.korder[1] = 0x76543210u;
      T.k_steps = 8;
      T.inv_scale = 1.f;
      T.mma_n = (int16_t)(head ? 32 : 128);
      T.kind = T1.kind;
      T.bias_off = T1.bias_off;
      T.table = T1.table;
      const int rows = T.mma_n / 2;  // weight rows per CTA and tile
      T.tile_bytes = rows * 128;
      T.stage_bytes = 2 * T.tile_bytes;
      for (int j = 0; j < T.n_chunks; j++)
        for (int st = 0; st < T.k_stages; st++)
          for (int r = 0; r < 2; r++)
            for (int kc = 0; kc < 2; kc++) {
              const size_t o = blob[r].size();
              blob[r].resize(o + T.tile_bytes);
              blobh[r].resize(o + T.tile_bytes);
              pack_tile(W[l], ins[l], n_real, k_real, j * T.mma_n + r * rows, (2 * st + kc) * 64, rows, blob[r].data() + o);
              pack_tile(W[l], ins[l], n_real, k_real, j * T.mma_n + r * rows, (2 * st + kc) * 64, rows, blobh[r].data() + o, true);
            }
    }
    for (int r = 0; r < 2; r++) {
      DJ_TRY(cudaMalloc((void**)&p->stream2[net][r], blob[r].size()));
      DJ_TRY(cudaMemcpy(p->stream2[net][r], blob[r].data(), blob[r].size(), cudaMemcpyHostToDevice));
      DJ_TRY(cudaMalloc((void**)&p->stream2h[net][r], blobh[r].size()));
      DJ_TRY(cudaMemcpy(p->stream2h[net][r], blobh[r].data(), blobh[r].size(), cudaMemcpyHostToDevice));
      N.stream[r] = p->stream2[net][r];
    }
  }
  // ---- split-operand program: the pair program with every stage doubled (W_hi, W_lo) and per-layer scales
  for (int net = 0; net < 2; net++) {
    const int nl = net == 0 ? d->f_layers : d->h_layers;
    const float* const* W = net == 0 ? f_w : h_w;
    const std::vector<int>& outs = net == 0 ? p->f_out : p->h_out;
    const std::vector<int>& ins = net == 0 ? p->f_in : p->h_in;
    tc2::Net& N = p->nets2x[net];
    N = p->nets2[net];
    std::vector<uint8_t> blob[2];
    {
      tc2::Layer& T = N.layers[0];
      const float sc = x2_scale(W[0], (size_t)outs[0] * ins[0]);
      T.inv_scale = 1.f / sc;
      const int xyz_off = (net == 0 ? L : Lb);
      for (int j = 0; j < 4; j++)
        for (int r = 0; r < 2; r++) {
          const size_t o = blob[r].size();
          blob[r].resize(o + T.tile_bytes);
          pack_tile_l0_x2(W[0], ins[0], xyz_off, j * 128 + r * 64, 64, blob[r].data() + o, sc);
        }
    }
    for (int l = 1; l < nl; l++) {
      tc2::Layer& T = N.layers[l];
      const bool reinj = (net == 0 && l == li);
      const int n_real = outs[l];
      const int k_real = reinj ? K3 : ins[l];
      const float sc = x2_scale(W[l], (size_t)outs[l] * ins[l]);
      T.inv_scale = 1.f / sc;
      const int k_ranges = T.k_ranges;
      T.k_stages = (int16_t)(2 * k_ranges);
      const int rows = T.mma_n / 2;
      // stage order of a chunk as (part, K range): W_lo stages before W_hi stages (mlp_tc2.cuh); chunk 0 defers its
      // last K range, the one the previous layer's drain publishes last
      std::vector<std::pair<int, int>> order[2];
      for (int st = 0; st + 1 < k_ranges; st++) order[0].push_back({1, st});
      for (int st = 0; st + 1 < k_ranges; st++) order[0].push_back({0, st});
      order[0].push_back({1, k_ranges - 1});
      order[0].push_back({0, k_ranges - 1});
      for (int part = 1; part >= 0; part--)
        for (int st = 0; st < k_ranges; st++) order[1].push_back({part, st});
      for (int c = 0; c < 2; c++) {
        T.korder[c] = 0;
        for (size_t i = 0; i < order[c].size(); i++) T.korder[c] |= (uint32_t)order[c][i].second << (4 * i);
      }
      for (int j = 0; j < T.n_chunks; j++)
        for (const auto& ps : order[j == 0 ? 0 : 1])
          for (int r = 0; r < 2; r++)
            for (int kc = 0; kc < 2; kc++) {
              const size_t o = blob[r].size();
              blob[r].resize(o + T.tile_bytes);
              pack_tile_x2(W[l], ins[l], n_real, k_real, j * T.mma_n + r * rows, (2 * ps.second + kc) * 64, rows, blob[r].data() + o, sc, ps.first);
            }
    }
    for (int r = 0; r < 2; r++) {
      DJ_TRY(cudaMalloc((void**)&p->stream2x[net][r], blob[r].size()));
      DJ_TRY(cudaMemcpy(p->stream2x[net][r], blob[r].data(), blob[r].size(), cudaMemcpyHostToDevice));
      N.stream[r] = p->stream2x[net][r];
    }
  }
  // ---- three-pass split-operand program (tc2_forward_kernel<.., 2>): 128 x 256 x 16 MMAs, 256 output columns per
  // accumulator chunk (128 weight rows per CTA), one ring stage = one [128 x 64] tile; per K range and k-chunk the
  // W_lo tile, then the W_hi tile
  for (int net = 0; net < 2; net++) {
    const int nl = net == 0 ? d->f_layers : d->h_layers;
    const float* const* W = net == 0 ? f_w : h_w;
    const std::vector<int>& outs = net == 0 ? p->f_out : p->h_out;
    const std::vector<int>& ins = net == 0 ? p->f_in : p->h_in;
    tc2::Net& N = p->nets3x[net];
    N = p->nets2x[net];
    std::vector<uint8_t> blob[2];
    {
      tc2::Layer& T = N.layers[0];
      const float sc = 1.f / T.inv_scale;
      T.n_chunks = 2; T.mma_n = 256; T.tile_bytes = 128 * 128; T.stage_bytes = T.tile_bytes;
      const int xyz_off = (net == 0 ? L : Lb);
      for (int j = 0; j < 2; j++)
        for (int r = 0; r < 2; r++) {
          const size_t o = blob[r].size();
          blob[r].resize(o + T.tile_bytes);
          pack_tile_l0_x2(W[0], ins[0], xyz_off, j * 256 + r * 128, 128, blob[r].data() + o, sc);
        }
    }
    for (int l = 1; l < nl; l++) {
      tc2::Layer& T = N.layers[l];
      const bool head = (l == nl - 1);
      const bool reinj = (net == 0 && l == li);
      const int n_real = outs[l];
      const int k_real = reinj ? K3 : ins[l];
      const float sc = 1.f / T.inv_scale;
      const int k_ranges = T.k_ranges;
      if (k_ranges > 4) p->x3_ok = false;
      T.mma_n = (int16_t)(head ? 32 : 256);
      T.n_chunks = (int16_t)(head ? 1 : (n_real + 255) / 256);
      if (T.n_chunks > 2 || (T.n_chunks > 1 && k_ranges < 2)) p->x3_ok = false;
      if (!p->x3_ok) break;
      T.k_steps = 4;
      T.k_stages = (int16_t)(4 * k_ranges);
      const int rows = T.mma_n / 2;
      T.tile_bytes = rows * 128;
      T.stage_bytes = T.tile_bytes;
      // K range order: chunk 0 in the order the previous layer's drains publish them; chunk 1 reads range 1 first
      // so that the drain of chunk 0 may overwrite it early (R1_FREE)
      std::vector<int> order[2];
      for (int r = 0; r < k_ranges; r++) order[0].push_back(r);
      if (k_ranges > 1) order[1].push_back(1);
      for (int r = 0; r < k_ranges; r++)
        if (r != 1) order[1].push_back(r);
      for (int c = 0; c < 2; c++) {
        T.korder[c] = 0;
        for (size_t i = 0; i < order[c].size(); i++) T.korder[c] |= (uint32_t)order[c][i] << (4 * i);
      }
      for (int j = 0; j < T.n_chunks; j++)
        for (int r : order[j == 0 ? 0 : 1])
          for (int kc = 0; kc < 2; kc++)
            for (int part = 1; part >= 0; part--)
              for (int c = 0; c < 2; c++) {
                const size_t o = blob[c].size();
                blob[c].resize(o + T.tile_bytes);
                pack_tile_x2(W[l], ins[l], n_real, k_real, j * T.mma_n + c * rows, r * 128 + kc * 64, rows, blob[c].data() + o, sc, part);
              }
    }
    if (!p->x3_ok) break;
    for (int r = 0; r < 2; r++) {
      DJ_TRY(cudaMalloc((void**)&p->stream3x[net][r], blob[r].size()));
      DJ_TRY(cudaMemcpy(p->stream3x[net][r], blob[r].data(), blob[r].size(), cudaMemcpyHostToDevice));
      N.stream[r] = p->stream3x[net][r];
    }
  }
  // ---- latent-step program (latent_tc2.cuh): forward, loss, backward in one pass over the tile
  {
    const int nlf = d->f_layers, nlh = d->h_layers;
    if ((nlf - 1) + (nlh - 1) + 2 > lat2::MASK_LAYERS + 2 || nlf - 1 + nlh - 1 != lat2::MASK_LAYERS)
      return bail(fail(DJ_ERR_UNSUPPORTED, "latent step: %d + %d hidden activations (built for %d)", nlf - 1, nlh - 1, lat2::MASK_LAYERS));
    std::vector<uint8_t> blob[2], blobh[2];  // blobh: forward tiles in IEEE half (DJ_PREC_FP16), backward tiles bf16 as in blob
    // split-operand program: per entry, how to make the (part 0 = high, 1 = low) tile of (n0, k0), and the layer's scale
    std::vector<std::function<void(int, int, int, uint8_t*, int)>> gen_x(lat2::MAX_PROG);
    std::vector<float> scale_x(lat2::MAX_PROG, 1.f);
    int np = 0;
    auto add = [&](int kind, int src, int dst, int n_chunks, int k_stages, int k_steps, int mma_n, int mask_layer, int slot,
                   int bias_off, int after) -> lat2::Layer& {
      lat2::Layer& T = p->lat_prog[np++];
      memset(&T, 0, sizeof T);
      T.kind = (int8_t)kind; T.src = (int8_t)src; T.dst = (int8_t)dst; T.k_steps = (int8_t)k_steps;
      T.n_chunks = (int8_t)n_chunks; T.k_stages = (int8_t)k_stages; T.mask_layer = (int8_t)mask_layer; T.slot = (int8_t)slot;
      T.mma_n = (int16_t)mma_n; T.after = (int16_t)after; T.bias_off = bias_off;
      T.tile_bytes = (mma_n / 2) * 128;
      T.stage_bytes = (k_steps == 1 ? 1 : 2) * T.tile_bytes;
      return T;
    };
    // stages of one layer: chunk-major, then K; each CTA rank gets rows [r*mma_n/2, ...) of every chunk
    auto emit = [&](const lat2::Layer& T, auto&& tile_fn) {
      const int rows = T.mma_n / 2;
      for (int j = 0; j < T.n_chunks; j++)
        for (int st = 0; st < T.k_stages; st++)
          for (int r = 0; r < 2; r++)
            for (int kc = 0; kc < (T.k_steps == 1 ? 1 : 2); kc++) {
              const size_t o = blob[r].size();
              blob[r].resize(o + T.tile_bytes);
              blobh[r].resize(o + T.tile_bytes);
              tile_fn(j * T.mma_n + r * rows, (2 * st + kc) * 64, rows, blob[r].data() + o, false);
              tile_fn(j * T.mma_n + r * rows, (2 * st + kc) * 64, rows, blobh[r].data() + o, true);
            }
    };
    int buf = 0;  // where the current activations live: 0 shared memory, 1 tensor memory
    int mask = 0;
    for (int net = 0; net < 2; net++) {
      const int nl = net == 0 ? nlf : nlh;
      const float* const* W = net == 0 ? f_w : h_w;
      const std::vector<int>& outs = net == 0 ? p->f_out : p->h_out;
      const std::vector<int>& ins = net == 0 ? p->f_in : p->h_in;
      const tc2::Net& N2 = p->nets2[net];
      if (net == 1) p->lat_net1_lo = np;
      add(lat2::K_L0, 2, 0, 0, 0, 0, 0, mask++, net == 0 ? 0 : 2, 0, 0);
      buf = 0;
      for (int l = 1; l < nl; l++) {
        const tc2::Layer& F = N2.layers[l];
        const bool head = (l == nl - 1);
        const bool reinj = (net == 0 && l == li);
        const int k_real = reinj ? K3 : ins[l];
        const int kind = head ? (net == 0 ? lat2::K_HEAD_F : lat2::K_HEAD_H) : (reinj ? lat2::K_FWD_XYZ : lat2::K_FWD);
        lat2::Layer& T = add(kind, buf, head ? 2 : 1 - buf, F.n_chunks, F.k_stages, 8, F.mma_n, head ? 0 : mask, reinj ? 1 : 0,
                             F.bias_off, 0);
        emit(T, [&](int n0, int k0, int rows, uint8_t* dst, bool f16) { pack_tile(W[l], ins[l], outs[l], k_real, n0, k0, rows, dst, f16); });
        {
          const float sc = x2_scale(W[l], (size_t)outs[l] * ins[l]);
          const float* Wl = W[l];
          const int ldw = ins[l], n_real = outs[l];
          scale_x[np - 1] = sc;
          gen_x[np - 1] = [=](int n0, int k0, int rows, uint8_t* dst, int part) { pack_tile_x2(Wl, ldw, n_real, k_real, n0, k0, rows, dst, sc, part); };
        }
        if (!head) { mask++; buf = 1 - buf; }
      }
    }
    // backward: hnet first (its head operand is written right after the loss), then fnet
    for (int net = 1; net >= 0; net--) {
      const int nl = net == 0 ? nlf : nlh;
      const float* const* W = net == 0 ? f_w : h_w;
      const std::vector<int>& outs = net == 0 ? p->f_out : p->h_out;
      const std::vector<int>& ins = net == 0 ? p->f_in : p->h_in;
      const int mask0 = net == 0 ? 0 : nlf - 1;  // mask layer of this net's layer-0 output
      if (net == 0) p->lat_net1_hi = np;
      // head: K = 16 operand in shared memory -> gradient of the last hidden activation, into tensor memory
      {
        lat2::Layer& T = add(lat2::K_BWD, 0, 1, 4, 1, 1, 128, mask0 + nl - 2, 0, 0, 0);
        emit(T, [&](int n0, int, int rows, uint8_t* dst, bool) { pack_tile_head_T(W[nl - 1], ins[nl - 1], n0, rows, dst); });
        {
          const float* Wh = W[nl - 1];
          const int ldw = ins[nl - 1];
          gen_x[np - 1] = [=](int n0, int, int rows, uint8_t* dst, int) { pack_tile_head_T(Wh, ldw, n0, rows, dst); };
        }
      }
      buf = 1;
      for (int l = nl - 2; l >= 1; l--) {
        // gradient wrt the input of layer l (= output of layer l-1, mask layer mask0 + l - 1)
        const bool reinj = (net == 0 && l == li);
        const int in_real = reinj ? K3 : ins[l];              // columns of W_l that multiply hidden activations
        const int out_real = outs[l];                          // contraction length
        const int n_chunks = (in_real + 127) / 128, k_stages = ((out_real + 127) / 128);
        const bool last = (l == 1);
        const bool summed = last || (net == 0 && l - 1 == li);  // dPre of layer 0 / of the re-injection layer
        const int kind = last ? lat2::K_SUM : (summed ? lat2::K_BWD_SUM : lat2::K_BWD);
        const int slot = last ? (net == 0 ? 0 : 2) : 1;
        lat2::Layer& T = add(kind, buf, last ? 2 : 1 - buf, n_chunks, k_stages, 8, 128, mask0 + l - 1, slot, 0,
                             (last && net == 1) ? 1 : 0);
        emit(T, [&](int n0, int k0, int rows, uint8_t* dst, bool) { pack_tile_T(W[l], ins[l], out_real, in_real, n0, k0, rows, dst); });
        {
          const float* Wl = W[l];
          const int ldw = ins[l];
          gen_x[np - 1] = [=](int n0, int k0, int rows, uint8_t* dst, int part) { pack_tile_T_x2(Wl, ldw, out_real, in_real, n0, k0, rows, dst, part); };
        }
        buf = 1 - buf;
      }
    }
    p->lat_n_prog = np;
    {  // ---- the split-operand program: same entries, every K = 128 stage as (W_lo, W_hi), stage order as in the forward kernel
      std::vector<uint8_t> blobx[2];
      for (int l = 0; l < np; l++) {
        lat2::Layer& T = p->lat_prog_x[l];
        T = p->lat_prog[l];
        T.inv_scale = 1.f;
        T.k_ranges = T.k_stages;
        T.korder[0] = T.korder[1] = 0x76543210u;
        if (T.n_chunks == 0) continue;  // epilogue-only entry
        const int rows = T.mma_n / 2;
        if (T.k_steps == 1) {  // K = 16 head-gradient operand: already a compensated product, one stage per chunk
          for (int j = 0; j < T.n_chunks; j++)
            for (int r = 0; r < 2; r++) {
              const size_t o = blobx[r].size();
              blobx[r].resize(o + T.tile_bytes);
              gen_x[l](j * T.mma_n + r * rows, 0, rows, blobx[r].data() + o, 0);
            }
          continue;
        }
        const bool fwd_l = T.kind <= lat2::K_HEAD_H;
        if (fwd_l) T.inv_scale = 1.f / scale_x[l];
        const int k_ranges = T.k_ranges;
        if (2 * k_ranges > 8) return bail(fail(DJ_ERR_UNSUPPORTED, "latent step: more than 4 K ranges per layer"));
        T.k_stages = (int8_t)(2 * k_ranges);
        std::vector<std::pair<int, int>> order[2];
        for (int st = 0; st + 1 < k_ranges; st++) order[0].push_back({1, st});
        for (int st = 0; st + 1 < k_ranges; st++) order[0].push_back({0, st});
        order[0].push_back({1, k_ranges - 1});
        order[0].push_back({0, k_ranges - 1});
        for (int part = 1; part >= 0; part--)
          for (int st = 0; st < k_ranges; st++) order[1].push_back({part, st});
        for (int c = 0; c < 2; c++) {
          T.korder[c] = 0;
          for (size_t i = 0; i < order[c].size(); i++) T.korder[c] |= (uint32_t)order[c][i].second << (4 * i);
        }
        for (int j = 0; j < T.n_chunks; j++)
          for (const auto& ps : order[j == 0 ? 0 : 1])
            for (int r = 0; r < 2; r++)
              for (int kc = 0; kc < 2; kc++) {
                const size_t o = blobx[r].size();
                blobx[r].resize(o + T.tile_bytes);
                gen_x[l](j * T.mma_n + r * rows, (2 * ps.second + kc) * 64, rows, blobx[r].data() + o, ps.first);
              }
      }
      for (int r = 0; r < 2; r++) {
        DJ_TRY(cudaMalloc((void**)&p->lat_stream_x[r], blobx[r].size()));
        DJ_TRY(cudaMemcpy(p->lat_stream_x[r], blobx[r].data(), blobx[r].size(), cudaMemcpyHostToDevice));
      }
    }
    for (int r = 0; r < 2; r++) {
      DJ_TRY(cudaMalloc((void**)&p->lat_stream[r], blob[r].size()));
      DJ_TRY(cudaMemcpy(p->lat_stream[r], blob[r].data(), blob[r].size(), cudaMemcpyHostToDevice));
      DJ_TRY(cudaMalloc((void**)&p->lat_stream_h[r], blobh[r].size()));
      DJ_TRY(cudaMemcpy(p->lat_stream_h[r], blobh[r].data(), blobh[r].size(), cudaMemcpyHostToDevice));
    }
  }
  DJ_TRY(upload(bias_host.data(), bias_host.size(), &p->bias_dev));
#undef DJ_TRY
  *out = p;
  return DJ_OK;
}

extern "C" int dj_pack_destroy(DjPack* p) {
  free_pack(p);
  return DJ_OK;
}

// ====================================================================== forward
struct GridSpec {
  int on = 0;
  int d0 = 0, d1 = 0, d2 = 0;
  double stop = 0, offs = 0;
  int64_t r_lo = 0;
  double step(int d) const { return d > 1 ? stop / (double)(d - 1) : 0.0; }
};

static int check_err_word(DjPack* p, cudaStream_t st) {
  unsigned int h = 0;
  DJ_CTX(c, p, st);
  DJ_CUDA(cudaMemcpyAsync(&h, c->err, sizeof h, cudaMemcpyDeviceToHost, st));
  DJ_CUDA(cudaStreamSynchronize(st));
  if (h) return fail(DJ_ERR_DEVICE, "device watchdog fired: 0x%08x", h);
  return DJ_OK;
}

static int launch_fold(DjPack* p, const float* code_dev, cudaStream_t st) {
  DJ_CTX(c, p, st);
  simt::fold_kernel<<<(3 * 512 * 32 + 255) / 256, 256, 0, st>>>(p->fold, code_dev, c->tables);
  DJ_LAUNCHED();
  return DJ_OK;
}

template <bool TB>
static int launch_sgemm(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int64_t M, int N, int K,
                        const float* bias, int bstride, int act, float slope, const float* dsrc, int ldd, cudaStream_t st) {
  dim3 grid((N + simt::BN - 1) / simt::BN, (unsigned)((M + simt::BM - 1) / simt::BM));
  simt::sgemm_kernel<TB><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, (int)M, N, K, bias, bstride, act, slope, dsrc, ldd);
  DJ_LAUNCHED();
  return DJ_OK;
}

// fp32 forward of both nets for n points.  fH[l] / hH[l]: output buffer of layer l ([n,512], ld 512);
// yC / yT: [n,8].  Distinct buffers per layer when the activations are kept for the backward.
static int fp32_forward(DjPack* p, const float* xyz, int64_t n, int net_mask, float* const* fH, float* const* hH,
                        float* yC, float* yT, cudaStream_t st) {
  const float slope = p->desc.slope;
  DJ_CTX(c, p, st);
  const float* tabf = reinterpret_cast<const float*>(c->tables);
  const int H = p->H, K3 = p->K3, li = p->desc.latent_in;
  int rc;
  if (net_mask & 1) {
    const int nl = p->desc.f_layers;
    if ((rc = launch_sgemm<true>(xyz, 3, tabf, 4, fH[0], H, n, H, 3, tabf + 3, 4, 1, slope, nullptr, 0, st))) return rc;
    for (int l = 1; l < nl; l++) {
      const float* in = fH[l - 1];
      if (l == nl - 1) {
        if ((rc = launch_sgemm<true>(in, H, p->fW[l], H, yC, 8, n, 5, H, p->fB[l], 1, 0, slope, nullptr, 0, st))) return rc;
      } else if (l == li) {
        simt::put_xyz_cols_kernel<<<ceil_div(n, 256), 256, 0, st>>>(xyz, fH[l - 1], H, K3, n);
        DJ_LAUNCHED();
        if ((rc = launch_sgemm<true>(in, H, p->fWx, K3 + 3, fH[l], H, n, H, K3 + 3, tabf + 4 * 512 + 3, 4, 1, slope, nullptr, 0, st))) return rc;
      } else {
        if ((rc = launch_sgemm<true>(in, H, p->fW[l], H, fH[l], H, n, p->f_out[l], H, p->fB[l], 1, 1, slope, nullptr, 0, st))) return rc;
      }
    }
  }
  if (net_mask & 2) {
    const int nl = p->desc.h_layers;
    const float* t2 = tabf + 2 * 4 * 512;
    if ((rc = launch_sgemm<true>(xyz, 3, t2, 4, hH[0], H, n, H, 3, t2 + 3, 4, 1, slope, nullptr, 0, st))) return rc;
    for (int l = 1; l < nl; l++) {
      if (l == nl - 1) {
        if ((rc = launch_sgemm<true>(hH[l - 1], H, p->hW[l], H, yT, 8, n, 5, H, p->hB[l], 1, 1, slope, nullptr, 0, st))) return rc;
      } else {
        if ((rc = launch_sgemm<true>(hH[l - 1], H, p->hW[l], H, hH[l], H, n, H, H, p->hB[l], 1, 1, slope, nullptr, 0, st))) return rc;
      }
    }
  }
  return DJ_OK;
}

// epilogue warps of the three-pass split-operand forward (mlp_tc2.cuh)
#ifndef DJ_X3_EW
#define DJ_X3_EW 16
#endif
static constexpr int X3_EW = DJ_X3_EW;

// do_fold = false: the caller already folded this code into the pack's tables (several launches with one code)
static int run_forward(DjPack* p, const float* code_dev, const float* xyz_dev, const GridSpec& G, int64_t n,
                       const HeadCfg& head, int net_mask, int precision, float* out_dev, cudaStream_t st, bool do_fold = true) {
  if (n <= 0) return DJ_OK;
  int rc;
  DJ_CTX(cx, p, st);
  if (do_fold && (rc = launch_fold(p, code_dev, st))) return rc;
  if (precision == DJ_PREC_BF16 || precision == DJ_PREC_FP16 || precision == DJ_PREC_FP16X2) {
    const bool f16 = precision == DJ_PREC_FP16;
    const bool x2 = precision == DJ_PREC_FP16X2;
    tc2::Params P;
    memset(&P, 0, sizeof P);
    // DJ_PREC_FP16X2 runs the three-pass program; DEEPJOIN_B200_X2_FOURPASS=1 selects the four-pass one (diagnostics)
#ifdef DJ_DEBUG
    const bool four_pass = getenv("DEEPJOIN_B200_X2_FOURPASS") != nullptr;
#else
    static const bool four_pass = getenv("DEEPJOIN_B200_X2_FOURPASS") != nullptr;
#endif
    const bool x3 = x2 && !four_pass && p->x3_ok;
    P.nets[0] = x3 ? p->nets3x[0] : (x2 ? p->nets2x[0] : p->nets2[0]);
    P.nets[1] = x3 ? p->nets3x[1] : (x2 ? p->nets2x[1] : p->nets2[1]);
    if (f16)
      for (int net = 0; net < 2; net++)
        for (int r = 0; r < 2; r++) P.nets[net].stream[r] = p->stream2h[net][r];
    P.net_mask = net_mask;
    P.grid_mode = G.on;
#ifdef DJ_DEBUG
    if (const char* e = getenv("DEEPJOIN_B200_DEBUG")) P.debug = atoi(e);
#endif
    P.bias = p->bias_dev;
    P.tables = cx->tables;
    P.xyz = xyz_dev;
    P.n = n;
    P.d0 = G.d0; P.d1 = G.d1; P.d2 = G.d2;
    P.step0 = G.step(G.d0); P.step1 = G.step(G.d1); P.step2 = G.step(G.d2);
    P.stop = G.stop; P.offs = G.offs; P.r_lo = G.r_lo;
    P.head = head;
    P.out = out_dev;
    P.err = cx->err;
    P.prof = p->prof;
    const int tile_pts = x2 ? tc2::TILE_PTS_X2 : tc2::TILE_M;
    const int64_t ntiles = (n + tile_pts - 1) / tile_pts;
    const int grid = (int)std::min<int64_t>((ntiles + 1) / 2 * 2, p->num_sms / 2 * 2);  // whole CTA pairs
    if (!p->tc_attr_set2) {
      DJ_CUDA(cudaFuncSetAttribute(tc2::tc2_forward_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES));
#ifdef DJ_DEBUG
      DJ_CUDA(cudaFuncSetAttribute(tc2::tc2_forward_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES));
#endif
      DJ_CUDA(cudaFuncSetAttribute(tc2::tc2_forward_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES));
      DJ_CUDA(cudaFuncSetAttribute(tc2::tc2_forward_kernel<0, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES));
      DJ_CUDA(cudaFuncSetAttribute(tc2::tc2_forward_kernel<0, 1, 2, X3_EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES));
      p->tc_attr_set2 = true;
    }
#ifdef DJ_DEBUG
    if (x2 && (P.debug & 64)) DJ_CUDA(cudaMemsetAsync(p->prof, 0, 64 * sizeof(unsigned long long), st));
    if (x2 && !x3 && P.debug) {
      DJ_CUDA(cudaFuncSetAttribute(tc2::tc2_forward_kernel<1, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES));
      tc2::tc2_forward_kernel<1, 1, 1><<<grid, tc2::NTHREADS, tc2::SMEM_BYTES, st>>>(P);
    } else
    if (x3 && P.debug) {
      DJ_CUDA(cudaFuncSetAttribute(tc2::tc2_forward_kernel<1, 1, 2, X3_EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES));
      tc2::tc2_forward_kernel<1, 1, 2, X3_EW><<<grid, 64 + 32 * X3_EW, tc2::SMEM_BYTES, st>>>(P);
    } else
#endif
    if (x3) tc2::tc2_forward_kernel<0, 1, 2, X3_EW><<<grid, 64 + 32 * X3_EW, tc2::SMEM_BYTES, st>>>(P);
    else if (x2) tc2::tc2_forward_kernel<0, 1, 1><<<grid, tc2::NTHREADS, tc2::SMEM_BYTES, st>>>(P);
    else if (f16) tc2::tc2_forward_kernel<0, 1><<<grid, tc2::NTHREADS, tc2::SMEM_BYTES, st>>>(P);
#ifdef DJ_DEBUG
    else if (P.debug) tc2::tc2_forward_kernel<1, 0><<<grid, tc2::NTHREADS, tc2::SMEM_BYTES, st>>>(P);
#endif
    else tc2::tc2_forward_kernel<0, 0><<<grid, tc2::NTHREADS, tc2::SMEM_BYTES, st>>>(P);
    DJ_LAUNCHED();
    return DJ_OK;
  }
  if (precision != DJ_PREC_FP32) return fail(DJ_ERR_INVALID, "unknown precision %d", precision);
  // fp32: layer-by-layer SGEMMs over chunks, two ping-pong activation buffers
  const int64_t CH = 65536;
  const int64_t c = std::min(CH, n);
  const size_t act = (size_t)c * 512 * sizeof(float);
  const size_t need = 2 * act + 2 * (size_t)c * 8 * sizeof(float) + (size_t)c * 3 * sizeof(float);
  DJ_CUDA(cx->ws_fwd.reserve(need));
  uint8_t* w = cx->ws_fwd.as<uint8_t>();
  float* bufA = reinterpret_cast<float*>(w);
  float* bufB = reinterpret_cast<float*>(w + act);
  float* yC = reinterpret_cast<float*>(w + 2 * act);
  float* yT = yC + c * 8;
  float* gx = yT + c * 8;
  std::vector<float*> fH(p->desc.f_layers), hH(p->desc.h_layers);
  for (int l = 0; l < p->desc.f_layers; l++) fH[l] = (l & 1) ? bufB : bufA;
  for (int l = 0; l < p->desc.h_layers; l++) hH[l] = (l & 1) ? bufB : bufA;
  for (int64_t s = 0; s < n; s += CH) {
    const int64_t m = std::min(CH, n - s);
    const float* xyz = xyz_dev ? xyz_dev + 3 * s : gx;
    if (G.on) {
      simt::grid_points_kernel<<<ceil_div(m, 256), 256, 0, st>>>(gx, m, G.r_lo + s, G.d0, G.d1, G.d2, G.step(G.d0),
                                                                G.step(G.d1), G.step(G.d2), G.stop, G.offs);
      DJ_LAUNCHED();
      xyz = gx;
    }
    if ((rc = fp32_forward(p, xyz, m, net_mask, fH.data(), hH.data(), yC, yT, st))) return rc;
    float* o = out_dev + (head.out_mode == OUT_ALL20 ? 20 * s : s);
    simt::heads_kernel<<<ceil_div(m, 256), 256, 0, st>>>(head, (net_mask & 1) ? yC : nullptr, (net_mask & 2) ? yT : nullptr, 8, o, m);
    DJ_LAUNCHED();
  }
  return DJ_OK;
}

static int net_mask_for(int use_net) { return use_net == 0 ? 1 : (use_net == 3 ? 2 : 3); }

extern "C" int dj_query_points(DjPack* p, const float* code_dev, const float* xyz_dev, int64_t n, int use_net,
                               int return_occ, int apply_sigmoid, int precision, float* out_dev, void* stream) {
  if (!p || !code_dev || (!xyz_dev && n > 0) || (!out_dev && n > 0) || n < 0) return fail(DJ_ERR_INVALID, "dj_query_points: bad argument");
  if (use_net < 0 || use_net > 3) return fail(DJ_ERR_BAD_NET, "Requested output from non-existent network: %d", use_net);
  DeviceGuard g(p->device);
  HeadCfg H{OUT_SINGLE, use_net, return_occ, apply_sigmoid, p->desc.slope};
  return run_forward(p, code_dev, xyz_dev, GridSpec{}, n, H, net_mask_for(use_net), precision, out_dev, (cudaStream_t)stream);
}

extern "C" int dj_query_all(DjPack* p, const float* code_dev, const float* xyz_dev, int64_t n, int precision,
                            float* out_dev, void* stream) {
  if (!p || !code_dev || (!xyz_dev && n > 0) || (!out_dev && n > 0) || n < 0) return fail(DJ_ERR_INVALID, "dj_query_all: bad argument");
  DeviceGuard g(p->device);
  HeadCfg H{OUT_ALL20, 0, 0, 0, p->desc.slope};
  return run_forward(p, code_dev, xyz_dev, GridSpec{}, n, H, 3, precision, out_dev, (cudaStream_t)stream);
}

extern "C" int dj_query_grid(DjPack* p, const float* code_dev, int d0, int d1, int d2, double stop, double offs,
                             int64_t r_lo, int64_t r_hi, int use_net, int return_occ, int apply_sigmoid, int precision,
                             float* out_dev, void* stream) {
  if (!p || !code_dev || !out_dev) return fail(DJ_ERR_INVALID, "dj_query_grid: null argument");
  if (d0 < 1 || d1 < 1 || d2 < 1) return fail(DJ_ERR_INVALID, "dj_query_grid: dims must be >= 1");
  const int64_t tot = (int64_t)d0 * d1 * d2;
  if (r_lo < 0 || r_hi > tot || r_lo > r_hi) return fail(DJ_ERR_INVALID, "dj_query_grid: row range [%lld,%lld) outside [0,%lld)", (long long)r_lo, (long long)r_hi, (long long)tot);
  if (use_net < 0 || use_net > 3) return fail(DJ_ERR_BAD_NET, "Requested output from non-existent network: %d", use_net);
  DeviceGuard g(p->device);
  GridSpec G;
  G.on = 1; G.d0 = d0; G.d1 = d1; G.d2 = d2; G.stop = stop; G.offs = offs; G.r_lo = r_lo;
  HeadCfg H{OUT_SINGLE, use_net, return_occ, apply_sigmoid, p->desc.slope};
  return run_forward(p, code_dev, nullptr, G, r_hi - r_lo, H, net_mask_for(use_net), precision, out_dev, (cudaStream_t)stream);
}

// memcpy split over a few host threads (one core moves ~10 GB/s; the staging copies of a 256^3 query are 0.5 GB)
static void par_memcpy(void* dst, const void* src, size_t bytes) {
  const int nt = bytes >= ((size_t)8 << 20) ? 4 : 1;
  if (nt == 1) { memcpy(dst, src, bytes); return; }
  std::vector<std::thread> th;
  const size_t per = ((bytes / nt) + 4095) & ~(size_t)4095;
  for (int t = 0; t < nt; t++) {
    const size_t o = (size_t)t * per;
    if (o >= bytes) break;
    const size_t m = std::min(per, bytes - o);
    th.emplace_back([=] { memcpy((uint8_t*)dst + o, (const uint8_t*)src + o, m); });
  }
  for (auto& t : th) t.join();
}
static bool is_pinned_host(const void* q) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, q) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

extern "C" int dj_decode_samples_host_strided(DjPack* p, const float* code_host, const double* pts_host, int64_t n,
                                              int64_t pt_stride, int64_t dim_stride, int use_net, int return_occ,
                                              int apply_sigmoid, int precision, double* out_host) {
  if (!p || !code_host || (!pts_host && n > 0) || (!out_host && n > 0) || n < 0) return fail(DJ_ERR_INVALID, "dj_decode_samples_host: bad argument");
  // element (i, d) of the points at pts_host[i*pt_stride + d*dim_stride]: rows [n,3] (3, 1), or three coordinate
  // arrays (1, S >= n) -- what core.reconstruct.get_query_points returns (np.vstack(...).T, core/reconstruct.py:25-30)
  const bool soa = (pt_stride == 1 && dim_stride >= n && n > 1);
  if (!soa && !(pt_stride == 3 && dim_stride == 1) && n > 0)
    return fail(DJ_ERR_UNSUPPORTED, "dj_decode_samples_host: point strides (%lld, %lld) are neither rows nor coordinate arrays", (long long)pt_stride, (long long)dim_stride);
  if (use_net < 0 || use_net > 3) return fail(DJ_ERR_BAD_NET, "Requested output from non-existent network: %d", use_net);
  DeviceGuard g(p->device);
  // Software pipeline over chunks of 2^21 points, two slots.  Pageable host arrays are staged through pinned
  // memory by this thread (a cudaMemcpyAsync on pageable memory blocks the caller and serialises the chunks);
  // arrays that are already pinned (the Python wrapper takes its output from the pinned pool) are DMA targets as
  // they are.  Slot s is reused by chunk k+2 after chunk k's event: its download is then complete.
  const int64_t CH = 1 << 21;
  const int64_t c = std::min<int64_t>(CH, std::max<int64_t>(n, 1));
  const int nc = p->L + p->Lb;
  // device: [code | 2 x (pts f64 | pts f32 | out f32 | out f64)]
  const size_t o_p64 = 0, o_p32 = o_p64 + (size_t)c * 24, o_o32 = o_p32 + (size_t)c * 12,
               o_o64 = (o_o32 + (size_t)c * 4 + 15) & ~(size_t)15, half = (o_o64 + (size_t)c * 8 + 255) & ~(size_t)255;
  // one host-buffer call at a time per pack (private streams, staging buffers); device-pointer calls on other
  // streams are not affected: they work in their own contexts
  std::lock_guard<std::mutex> host_lock(p->host_mu);
  for (int i = 0; i < 2; i++) {
    if (!p->host_streams[i]) DJ_CUDA(cudaStreamCreateWithFlags(&p->host_streams[i], cudaStreamNonBlocking));
    if (!p->host_events[i]) DJ_CUDA(cudaEventCreateWithFlags(&p->host_events[i], cudaEventDisableTiming));
  }
  cudaStream_t sts[2] = {p->host_streams[0], p->host_streams[1]};
  cudaEvent_t done[2] = {p->host_events[0], p->host_events[1]};
  DJ_CTX(hc, p, sts[0]);
  {  // both private streams work on the same per-shape tables (folded once, below) and staging buffers
    std::lock_guard<std::mutex> lock(p->mu);
    p->ctxs[sts[1]] = p->ctxs[sts[0]];
  }
  DJ_CUDA(hc->ws_host.reserve(1024 + 2 * half));
  uint8_t* w0 = hc->ws_host.as<uint8_t>();
  const bool stage_in = n > 0 && !is_pinned_host(pts_host), stage_out = n > 0 && !is_pinned_host(out_host);
  // pinned: [code 1 KB | 2 x (pts f64 | out f64)]
  const size_t ph = (size_t)c * 32;
  DJ_CUDA(hc->pin_host.reserve(1024 + 2 * ph));
  uint8_t* h0 = hc->pin_host.as<uint8_t>();
  memcpy(h0, code_host, nc * sizeof(float));
  auto cleanup = [&](int rc) {
    for (int i = 0; i < 2; i++) cudaStreamSynchronize(sts[i]);
    return rc;
  };
  if (cudaMemcpyAsync(w0, h0, nc * sizeof(float), cudaMemcpyHostToDevice, sts[0]) != cudaSuccess)
    return cleanup(fail(DJ_ERR_CUDA, "dj_decode_samples_host: code upload failed"));
  {  // the fold writes the pack's shared tables: once, before either stream launches a decoder kernel
    int rc = launch_fold(p, (const float*)w0, sts[0]);
    if (rc) return cleanup(rc);
    if (cudaStreamSynchronize(sts[0]) != cudaSuccess) return cleanup(fail(DJ_ERR_CUDA, "dj_decode_samples_host: fold failed"));
  }
  HeadCfg H{OUT_SINGLE, use_net, return_occ, apply_sigmoid, p->desc.slope};
  const int64_t nchunks = (n + CH - 1) / CH;
  // chunk j's results leave its slot (after its event)
  auto retire = [&](int64_t j) -> int {
    const int slot = (int)(j & 1);
    if (cudaEventSynchronize(done[slot]) != cudaSuccess) return fail(DJ_ERR_CUDA, "dj_decode_samples_host: chunk %lld failed: %s", (long long)j, cudaGetErrorString(cudaGetLastError()));
    if (stage_out) par_memcpy(out_host + j * CH, h0 + 1024 + (size_t)slot * ph + (size_t)c * 24, (size_t)std::min(CH, n - j * CH) * 8);
    return DJ_OK;
  };
  for (int64_t k = 0; k < nchunks; k++) {
    const int64_t s = k * CH, m = std::min(CH, n - s);
    const int slot = (int)(k & 1);
    int rc;
    if (k >= 2 && (rc = retire(k - 2))) return cleanup(rc);
    // the fp32 path works in one shared activation workspace: it stays on one stream
    cudaStream_t st = sts[(precision != DJ_PREC_FP32) ? slot : 0];
    uint8_t* w = w0 + 1024 + (size_t)slot * half;
    uint8_t* hs = h0 + 1024 + (size_t)slot * ph;
    if (soa) {  // three coordinate runs of m doubles -> device [3][m]
      for (int dd = 0; dd < 3; dd++) {
        const void* src = pts_host + (size_t)dd * dim_stride + s;
        if (stage_in) { par_memcpy(hs + (size_t)dd * m * 8, src, (size_t)m * 8); src = hs + (size_t)dd * m * 8; }
        if (cudaMemcpyAsync(w + o_p64 + (size_t)dd * m * 8, src, (size_t)m * 8, cudaMemcpyHostToDevice, st) != cudaSuccess)
          return cleanup(fail(DJ_ERR_CUDA, "dj_decode_samples_host: upload failed"));
      }
      simt::soa_f64_to_xyz_f32_kernel<<<ceil_div(3 * m, 256), 256, 0, st>>>((const double*)(w + o_p64), (float*)(w + o_p32), m);
    } else {
      const void* src = pts_host + 3 * s;
      if (stage_in) { par_memcpy(hs, src, (size_t)m * 24); src = hs; }
      if (cudaMemcpyAsync(w + o_p64, src, (size_t)m * 24, cudaMemcpyHostToDevice, st) != cudaSuccess)
        return cleanup(fail(DJ_ERR_CUDA, "dj_decode_samples_host: upload failed"));
      simt::f64_to_f32_kernel<<<ceil_div(3 * m, 256), 256, 0, st>>>((const double*)(w + o_p64), (float*)(w + o_p32), 3 * m);
    }
    DJ_LAUNCHED();
    rc = run_forward(p, (const float*)w0, (const float*)(w + o_p32), GridSpec{}, m, H, net_mask_for(use_net),
                     precision, (float*)(w + o_o32), st, false);
    if (rc) return cleanup(rc);
    simt::f32_to_f64_kernel<<<ceil_div(m, 256), 256, 0, st>>>((const float*)(w + o_o32), (double*)(w + o_o64), m);
    DJ_LAUNCHED();
    void* dst = stage_out ? (void*)(hs + (size_t)c * 24) : (void*)(out_host + s);
    if (cudaMemcpyAsync(dst, w + o_o64, (size_t)m * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess)
      return cleanup(fail(DJ_ERR_CUDA, "dj_decode_samples_host: download failed"));
    if (cudaEventRecord(done[slot], st) != cudaSuccess) return cleanup(fail(DJ_ERR_CUDA, "dj_decode_samples_host: event failed"));
  }
  for (int64_t j = std::max<int64_t>(0, nchunks - 2); j < nchunks; j++) {
    int rc = retire(j);
    if (rc) return cleanup(rc);
  }
  int rc = cleanup(DJ_OK);
  if (rc) return rc;
  if (precision != DJ_PREC_FP32) return check_err_word(p, sts[0]);
  return DJ_OK;
}

extern "C" int dj_decode_samples_host(DjPack* p, const float* code_host, const double* pts_host, int64_t n, int use_net,
                                      int return_occ, int apply_sigmoid, int precision, double* out_host) {
  return dj_decode_samples_host_strided(p, code_host, pts_host, n, 3, 1, use_net, return_occ, apply_sigmoid, precision, out_host);
}

// ====================================================================== latent optimisation
struct LatWs {
  std::vector<float*> fH, hH;
  float *yC, *yT, *dyC, *dyT, *G0, *G1, *colsum, *loss_acc, *grad, *terms, *bx, *bs, *bn;
};

// small: the tensor-core path keeps no activations in global memory
static int carve_latent(DjPack* p, int64_t n, LatWs& W, cudaStream_t st, bool small = false) {
  DJ_CTX(c, p, st);
  const int nf = p->desc.f_layers - 1, nh = p->desc.h_layers - 1;
  const size_t act = small ? 0 : (size_t)n * 512;
  const size_t floats = (size_t)(nf + nh + 2) * act + 4 * (size_t)n * 8 + 3 * 512 + 8 + 256 + 8 + (size_t)n * 7 + 64;
  DJ_CUDA(c->ws_lat.reserve(floats * sizeof(float)));
  c->lat_ptrs_dev = nullptr;  // this layout overwrites the tensor-core step's pointer table
  float* q = c->ws_lat.as<float>();
  W.fH.resize(nf + 1); W.hH.resize(nh + 1);
  for (int l = 0; l < nf; l++) { W.fH[l] = q; q += act; }
  for (int l = 0; l < nh; l++) { W.hH[l] = q; q += act; }
  W.fH[nf] = nullptr; W.hH[nh] = nullptr;
  W.G0 = q; q += act;
  W.G1 = q; q += act;
  W.yC = q; q += n * 8; W.yT = q; q += n * 8; W.dyC = q; q += n * 8; W.dyT = q; q += n * 8;
  W.colsum = q; q += 3 * 512;
  W.loss_acc = q; q += 8;
  W.grad = q; q += 256;
  W.terms = q; q += 8;
  W.bx = q; q += 3 * n; W.bs = q; q += n; W.bn = q; q += 3 * n;
  return DJ_OK;
}

// forward (activations kept) + loss + backward to dL/dcode.  Leaves grad in W.grad, terms in W.terms.
static int latent_fwd_bwd(DjPack* p, const float* code_dev, const float* xyz, const float* sdf, const float* nrm, int64_t n,
                          const DjLatentCfg* cfg, LatWs& W, cudaStream_t st) {
  int rc;
  const float slope = p->desc.slope;
  const int H = p->H, K3 = p->K3, li = p->desc.latent_in, nlf = p->desc.f_layers, nlh = p->desc.h_layers;
  if ((rc = launch_fold(p, code_dev, st))) return rc;
  if ((rc = fp32_forward(p, xyz, n, 3, W.fH.data(), W.hH.data(), W.yC, W.yT, st))) return rc;
  DJ_CUDA(cudaMemsetAsync(W.colsum, 0, (3 * 512 + 8) * sizeof(float), st));  // colsum + loss_acc are contiguous
  simt::LossCfg LC{cfg->lambda1, cfg->lambda3, cfg->clamp_dist, slope, 1.0f / (float)n, cfg->loss_kind};
  simt::loss_grad_kernel<<<ceil_div(n, 256), 256, 0, st>>>(LC, W.yC, W.yT, 8, sdf, nrm, W.dyC, W.dyT, W.loss_acc, n);
  DJ_LAUNCHED();
  auto colsum = [&](const float* G, int slot) -> int {
    dim3 grid(16, 128), block(32, 8);
    simt::colsum_kernel<<<grid, block, 0, st>>>(G, H, n, H, W.colsum + slot * 512);
    DJ_LAUNCHED();
    return DJ_OK;
  };
  // ---- fnet: dPre_l = (dPre_{l+1} . W_{l+1}) * phi'(H_{l+1}),  H_{l+1} = output buffer of layer l
  {
    const float* g = W.dyC;
    int ldg = 8, kred = 5;
    float* pp[2] = {W.G0, W.G1};
    int t = 0;
    for (int l = nlf - 1; l >= 1; l--) {
      // gradient wrt the output of layer l-1 (width f_out[l-1]); W_l is [f_out[l], f_in[l]], first f_out[l-1] columns
      const int width = p->f_out[l - 1];
      if ((rc = launch_sgemm<false>(g, ldg, p->fW[l], p->f_in[l], pp[t], H, n, width, kred, nullptr, 0, 0, slope, W.fH[l - 1], H, st))) return rc;
      g = pp[t]; ldg = H; kred = width; t ^= 1;
      if (l - 1 == li) { if ((rc = colsum(g, 1))) return rc; }
      if (l - 1 == 0) { if ((rc = colsum(g, 0))) return rc; }
    }
  }
  {
    const float* g = W.dyT;
    int ldg = 8, kred = 5;
    float* pp[2] = {W.G0, W.G1};
    int t = 0;
    for (int l = nlh - 1; l >= 1; l--) {
      if ((rc = launch_sgemm<false>(g, ldg, p->hW[l], p->h_in[l], pp[t], H, n, H, kred, nullptr, 0, 0, slope, W.hH[l - 1], H, st))) return rc;
      g = pp[t]; ldg = H; kred = H; t ^= 1;
      if (l - 1 == 0) { if ((rc = colsum(g, 2))) return rc; }
    }
  }
  simt::GradFinishParams GP;
  for (int i = 0; i < 3; i++) { GP.src[i] = p->fold.src[i]; GP.colsum[i] = W.colsum + i * 512; }
  GP.L = p->L; GP.Lb = p->Lb;
  GP.reg_lambda = cfg->l2reg ? cfg->code_reg_lambda : 0.f;
  simt::grad_finish_kernel<<<p->L + p->Lb, 128, 0, st>>>(GP, code_dev, W.grad, W.loss_acc);
  DJ_LAUNCHED();
  simt::loss_finish_kernel<<<1, 32, 0, st>>>(W.loss_acc, W.terms);
  DJ_LAUNCHED();
  (void)K3;
  return DJ_OK;
}

// Workspace of the tensor-core latent step for a batch of S shapes: accumulators [S][ACC_STRIDE] (column sums +
// loss terms), grad [S][256], terms [S][8], device pointer table [3][S].
struct LatTcWs {
  float *acc, *grad, *terms;
  const float** ptrs;  // xyz[S] | sdf[S] | nrm[S]
};
static int carve_latent_tc(DjPack* p, int S, LatTcWs& W, cudaStream_t st) {
  DJ_CTX(c, p, st);
  const size_t floats = (size_t)S * (lat2::ACC_STRIDE + 256 + 8) + 64;
  DJ_CUDA(c->ws_lat.reserve(floats * sizeof(float) + (size_t)3 * S * sizeof(void*) + 256));
  float* q = c->ws_lat.as<float>();
  W.acc = q; q += (size_t)S * lat2::ACC_STRIDE;
  W.grad = q; q += (size_t)S * 256;
  W.terms = q; q += (size_t)S * 8;
  W.ptrs = reinterpret_cast<const float**>((reinterpret_cast<uintptr_t>(q) + 63) & ~(uintptr_t)63);
  return DJ_OK;
}
// upload the per-shape sample array pointers (once per optimisation, not per iteration)
static int upload_ptrs(DjPack* p, LatTcWs& W, int S, const float* const* xyz, const float* const* sdf, const float* const* nrm, cudaStream_t st) {
  DJ_CTX(c, p, st);
  std::vector<const float*> h((size_t)3 * S);
  for (int s = 0; s < S; s++) { h[s] = xyz[s]; h[S + s] = sdf[s]; h[2 * S + s] = nrm[s]; }
  // the next block of a block-wise loop (DjLatentCfg.iter_offset > 0) finds them in place: no copy, and no
  // synchronisation, so the block is enqueued while the previous one still runs
  if (c->lat_ptrs_dev == W.ptrs && c->lat_ptrs == h) return DJ_OK;
  DJ_CUDA(cudaStreamSynchronize(st));  // a loop still reading the old pointers
  c->lat_ptrs = h;
  c->lat_ptrs_dev = W.ptrs;
  DJ_CUDA(cudaMemcpyAsync(W.ptrs, c->lat_ptrs.data(), h.size() * sizeof(void*), cudaMemcpyHostToDevice, st));
  DJ_CUDA(cudaStreamSynchronize(st));
  return DJ_OK;
}

// tensor-core forward + loss + backward (latent_tc2.cuh) for S shapes in ONE launch: codes [S][C] -> grad [S][256],
// terms [S][8].  idx (optional): this iteration's rows of every shape's pool ([S] rows of n, idx_stride apart),
// gathered inside the kernel.
static int latent_fwd_bwd_tc(DjPack* p, int S, const float* codes_dev, const int32_t* idx, int64_t idx_stride, int64_t n,
                             const DjLatentCfg* cfg, LatTcWs& W, cudaStream_t st) {
  const int C = p->L + p->Lb;
  DJ_CTX(c, p, st);
  simt::fold_kernel<<<dim3((3 * 512 * 32 + 255) / 256, S), 256, 0, st>>>(p->fold, codes_dev, c->tables, C);
  DJ_LAUNCHED();
  DJ_CUDA(cudaMemsetAsync(W.acc, 0, (size_t)S * lat2::ACC_STRIDE * sizeof(float), st));
  const bool x2 = cfg->precision == DJ_PREC_FP16X2;  // split operands, forward and backward (64 samples per CTA)
  const int tile_pts = x2 ? tc2::TILE_PTS_X2 : lat2::TILE_M;
  const int64_t tps = (n + tile_pts - 1) / tile_pts;
  const int64_t ntiles = tps * S;
  const int grid = (int)std::min<int64_t>((ntiles + 1) / 2 * 2, p->num_sms / 2 * 2);
  DJ_CUDA(c->ws_mask.reserve((size_t)grid * lat2::MASK_LAYERS * 8 * 128 * sizeof(unsigned long long)));
  lat2::Params P;
  memset(&P, 0, sizeof P);
  memcpy(P.prog, x2 ? p->lat_prog_x : p->lat_prog, sizeof(lat2::Layer) * p->lat_n_prog);
  P.n_prog = p->lat_n_prog;
  const bool half_fwd = cfg->precision == DJ_PREC_FP16;  // forward in IEEE half (8x finer than bf16: the loss gates follow
                                                         // the fp32 forward far more often); backward bf16 (range)
  P.stream[0] = x2 ? p->lat_stream_x[0] : (half_fwd ? p->lat_stream_h[0] : p->lat_stream[0]);
  P.stream[1] = x2 ? p->lat_stream_x[1] : (half_fwd ? p->lat_stream_h[1] : p->lat_stream[1]);
  P.bias = p->bias_dev;
  P.tables = c->tables;
  P.xyz = W.ptrs; P.sdf = W.ptrs + S; P.nrm = W.ptrs + 2 * S;
  P.idx = idx; P.idx_stride = idx_stride;
  P.n = n;
  P.n_shapes = S; P.tiles_per_shape = (int32_t)tps;
  P.loss = simt::LossCfg{cfg->lambda1, cfg->lambda3, cfg->clamp_dist, p->desc.slope, 1.0f / (float)n, cfg->loss_kind};
  if (cfg->loss_kind == DJ_LOSS_DEEPSDF) { P.skip_lo = p->lat_net1_lo; P.skip_hi = p->lat_net1_hi; }
  P.masks = c->ws_mask.as<unsigned long long>();
  P.acc = W.acc;
  P.err = c->err;
  if (!p->lat_attr_set) {
    DJ_CUDA(cudaFuncSetAttribute(lat2::latent_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lat2::SMEM_BYTES));
    DJ_CUDA(cudaFuncSetAttribute(lat2::latent_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lat2::SMEM_BYTES));
    DJ_CUDA(cudaFuncSetAttribute(lat2::latent_tc2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lat2::SMEM_BYTES));
    p->lat_attr_set = true;
  }
  if (x2) lat2::latent_tc2_kernel<true, true><<<grid, lat2::NTHREADS, lat2::SMEM_BYTES, st>>>(P);
  else if (half_fwd) lat2::latent_tc2_kernel<true><<<grid, lat2::NTHREADS, lat2::SMEM_BYTES, st>>>(P);
  else lat2::latent_tc2_kernel<false><<<grid, lat2::NTHREADS, lat2::SMEM_BYTES, st>>>(P);
  DJ_LAUNCHED();
  simt::GradFinishParams GP;
  for (int i = 0; i < 3; i++) { GP.src[i] = p->fold.src[i]; GP.colsum[i] = W.acc + i * 512; }
  GP.L = p->L; GP.Lb = p->Lb;
  GP.reg_lambda = cfg->l2reg ? cfg->code_reg_lambda : 0.f;
  simt::grad_finish_kernel<<<dim3(C, S), 128, 0, st>>>(GP, codes_dev, W.grad, W.acc + 3 * 512, lat2::ACC_STRIDE, C, 256);
  DJ_LAUNCHED();
  simt::loss_finish_kernel<<<S, 32, 0, st>>>(W.acc + 3 * 512, W.terms, lat2::ACC_STRIDE, 8);
  DJ_LAUNCHED();
  return DJ_OK;
}

static int check_latent_cfg(DjPack* p, const DjLatentCfg* cfg) {
  if (!cfg) return fail(DJ_ERR_INVALID, "null DjLatentCfg");
  if (cfg->precision != DJ_PREC_FP32 && cfg->precision != DJ_PREC_BF16 && cfg->precision != DJ_PREC_FP16 && cfg->precision != DJ_PREC_FP16X2)
    return fail(DJ_ERR_INVALID, "latent gradient: unknown precision %d", cfg->precision);
  if (p->L + p->Lb > 256) return fail(DJ_ERR_UNSUPPORTED, "code length > 256");
  if (cfg->loss_kind != DJ_LOSS_DEEPJOIN && cfg->loss_kind != DJ_LOSS_DEEPSDF)
    return fail(DJ_ERR_INVALID, "latent gradient: unknown loss_kind %d", cfg->loss_kind);
  return DJ_OK;
}

extern "C" int dj_latent_grad(DjPack* p, const float* code_dev, const float* xyz_dev, const float* sdf_dev,
                              const float* nrm_dev, int64_t n, const DjLatentCfg* cfg, float* grad_dev,
                              float* loss_terms_dev, void* stream) {
  if (!p || !code_dev || !xyz_dev || !sdf_dev || !nrm_dev || !grad_dev || n <= 0) return fail(DJ_ERR_INVALID, "dj_latent_grad: bad argument");
  int rc;
  if ((rc = check_latent_cfg(p, cfg))) return rc;
  DeviceGuard g(p->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (cfg->precision != DJ_PREC_FP32) {  // tensor cores (bf16 operands; gradients underflow in IEEE half)
    LatTcWs T;
    if ((rc = carve_latent_tc(p, 1, T, st))) return rc;
    if ((rc = upload_ptrs(p, T, 1, &xyz_dev, &sdf_dev, &nrm_dev, st))) return rc;
    if ((rc = latent_fwd_bwd_tc(p, 1, code_dev, nullptr, 0, n, cfg, T, st))) return rc;
    if ((rc = check_err_word(p, st))) return rc;
    DJ_CUDA(cudaMemcpyAsync(grad_dev, T.grad, (p->L + p->Lb) * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (loss_terms_dev) DJ_CUDA(cudaMemcpyAsync(loss_terms_dev, T.terms, 5 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return DJ_OK;
  }
  LatWs W;
  if ((rc = carve_latent(p, n, W, st))) return rc;
  if ((rc = latent_fwd_bwd(p, code_dev, xyz_dev, sdf_dev, nrm_dev, n, cfg, W, st))) return rc;
  DJ_CUDA(cudaMemcpyAsync(grad_dev, W.grad, (p->L + p->Lb) * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (loss_terms_dev) DJ_CUDA(cudaMemcpyAsync(loss_terms_dev, W.terms, 5 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return DJ_OK;
}

extern "C" int dj_latent_optimize_batch(DjPack* p, int n_shapes, float* codes_dev, float* adam_m_dev, float* adam_v_dev,
                                        const float* const* pool_xyz_dev, const float* const* pool_sdf_dev,
                                        const float* const* pool_nrm_dev, const int64_t* pool_n, const int32_t* idx_dev,
                                        const DjLatentCfg* cfg, float* log_dev, void* stream) {
  if (!p || !codes_dev || !adam_m_dev || !adam_v_dev || !pool_xyz_dev || !pool_sdf_dev || !pool_nrm_dev || !pool_n || !idx_dev)
    return fail(DJ_ERR_INVALID, "dj_latent_optimize_batch: null argument");
  if (n_shapes < 1 || n_shapes > DJ_MAX_BATCH) return fail(DJ_ERR_INVALID, "dj_latent_optimize_batch: 1..%d shapes per call (got %d)", DJ_MAX_BATCH, n_shapes);
  for (int s = 0; s < n_shapes; s++)
    if (!pool_xyz_dev[s] || !pool_sdf_dev[s] || !pool_nrm_dev[s] || pool_n[s] <= 0) return fail(DJ_ERR_INVALID, "dj_latent_optimize_batch: bad pool %d", s);
  int rc;
  if ((rc = check_latent_cfg(p, cfg))) return rc;
  if (cfg->num_iterations < 1 || cfg->num_samples < 1) return fail(DJ_ERR_INVALID, "dj_latent_optimize: iterations/samples must be >= 1");
  DeviceGuard g(p->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = cfg->num_samples;
  const int C = p->L + p->Lb;
  const int log_every = cfg->log_every > 0 ? cfg->log_every : 10;
  const int total = cfg->total_iterations > 0 ? cfg->total_iterations : cfg->num_iterations;
  const int e0 = cfg->iter_offset;
  if (e0 < 0 || e0 + cfg->num_iterations > total) return fail(DJ_ERR_INVALID, "dj_latent_optimize: iterations [%d, %d) outside the loop of %d", e0, e0 + cfg->num_iterations, total);
  const int64_t log_rows = (total + log_every - 1) / log_every;
  simt::AdamCfg A{cfg->lr, cfg->beta1, cfg->beta2, cfg->eps, total, cfg->loss_kind == DJ_LOSS_DEEPSDF ? 1 : 0};
  if (cfg->precision != DJ_PREC_FP32) {
    // all shapes in one launch per iteration; the kernel gathers every shape's mini-batch rows itself
    LatTcWs T;
    if ((rc = carve_latent_tc(p, n_shapes, T, st))) return rc;
    if ((rc = upload_ptrs(p, T, n_shapes, pool_xyz_dev, pool_sdf_dev, pool_nrm_dev, st))) return rc;
    const int64_t idx_stride = (int64_t)cfg->num_iterations * n;
    for (int el = 0; el < cfg->num_iterations; el++) {
      const int e = e0 + el;
      if ((rc = latent_fwd_bwd_tc(p, n_shapes, codes_dev, idx_dev + (size_t)el * n, idx_stride, n, cfg, T, st))) return rc;
      float* row = (log_dev && e % log_every == 0) ? log_dev + (size_t)(e / log_every) * 8 : nullptr;
      simt::adam_kernel<<<n_shapes, 256, 0, st>>>(A, e, codes_dev, adam_m_dev, adam_v_dev, T.grad, C, T.terms, row, p->L, C, 256, 8,
                                                  (long long)log_rows * 8);
      DJ_LAUNCHED();
    }
    // a device watchdog that fired during the loop must not go unnoticed: checked after the LAST block of the loop
    // (this synchronises the stream; the caller reads the codes right after anyway).  Earlier blocks of a block-wise
    // loop return as soon as they are enqueued, so that the host can stage the next block meanwhile.
    if (e0 + cfg->num_iterations < total) return DJ_OK;
    return check_err_word(p, st);
  }
  // fp32 (reference arithmetic): shape after shape
  LatWs W;
  if ((rc = carve_latent(p, n, W, st))) return rc;
  for (int s = 0; s < n_shapes; s++) {
    float* code = codes_dev + (size_t)s * C;
    const int32_t* idx = idx_dev + (size_t)s * cfg->num_iterations * n;
    for (int el = 0; el < cfg->num_iterations; el++) {
      const int e = e0 + el;
      simt::gather_kernel<<<ceil_div(n, 256), 256, 0, st>>>(pool_xyz_dev[s], pool_sdf_dev[s], pool_nrm_dev[s], idx + (size_t)el * n, (int)n, W.bx, W.bs, W.bn);
      DJ_LAUNCHED();
      if ((rc = latent_fwd_bwd(p, code, W.bx, W.bs, W.bn, n, cfg, W, st))) return rc;
      float* row = (log_dev && e % log_every == 0) ? log_dev + ((size_t)s * log_rows + e / log_every) * 8 : nullptr;
      simt::adam_kernel<<<1, 256, 0, st>>>(A, e, code, adam_m_dev + (size_t)s * C, adam_v_dev + (size_t)s * C, W.grad, C, W.terms, row, p->L);
      DJ_LAUNCHED();
    }
  }
  return DJ_OK;
}

extern "C" int dj_latent_optimize(DjPack* p, float* code_dev, float* adam_m_dev, float* adam_v_dev,
                                  const float* pool_xyz_dev, const float* pool_sdf_dev, const float* pool_nrm_dev,
                                  int64_t pool_n, const int32_t* idx_dev, const DjLatentCfg* cfg, float* log_dev,
                                  void* stream) {
  return dj_latent_optimize_batch(p, 1, code_dev, adam_m_dev, adam_v_dev, &pool_xyz_dev, &pool_sdf_dev, &pool_nrm_dev, &pool_n,
                                  idx_dev, cfg, log_dev, stream);
}

// ---- host: mesh writers.  The batch meshing worker (core/utils_multiprocessing.py:74-114) exports every mesh it
// extracts; numpy's record packing took longer (80 ms for 445 k vertices / 857 k faces) than the whole GPU pipeline
// (69 ms), a Python OBJ loop 4 s.  Same bytes as the Python writers in deepjoin_b200/mesh.py.
extern "C" int dj_host_write_ply(const char* path, const double* verts, int64_t nv, const int64_t* faces, int64_t nf) {
  if (!path || nv < 0 || nf < 0 || (nv && !verts) || (nf && !faces)) return fail(DJ_ERR_INVALID, "dj_host_write_ply: bad argument");
  char header[512];
  const int hl = snprintf(header, sizeof header,
                          "ply\nformat binary_little_endian 1.0\ncomment deepjoin_b200\nelement vertex %lld\nproperty float x\n"
                          "property float y\nproperty float z\nelement face %lld\nproperty list uchar int vertex_indices\nend_header\n",
                          (long long)nv, (long long)nf);
  std::vector<uint8_t> buf((size_t)hl + (size_t)nv * 12 + (size_t)nf * 13);
  memcpy(buf.data(), header, hl);
  float* vf = reinterpret_cast<float*>(buf.data() + hl);  // header length is not a multiple of 4: write through memcpy
  uint8_t* q = buf.data() + hl;
  for (int64_t i = 0; i < 3 * nv; i++) {
    const float x = (float)verts[i];
    memcpy(q + 4 * i, &x, 4);
  }
  (void)vf;
  q += (size_t)nv * 12;
  for (int64_t i = 0; i < nf; i++) {
    q[0] = 3;
    const int32_t t[3] = {(int32_t)faces[3 * i], (int32_t)faces[3 * i + 1], (int32_t)faces[3 * i + 2]};
    memcpy(q + 1, t, 12);
    q += 13;
  }
  FILE* fh = fopen(path, "wb");
  if (!fh) return fail(errno == EACCES ? DJ_ERR_PERMISSION : DJ_ERR_IO, "dj_host_write_ply: cannot open %s: %s", path, strerror(errno));
  const size_t wr = fwrite(buf.data(), 1, buf.size(), fh);
  const int cl = fclose(fh);
  if (wr != buf.size() || cl != 0) return fail(DJ_ERR_IO, "dj_host_write_ply: short write to %s", path);
  return DJ_OK;
}

extern "C" int dj_host_write_obj(const char* path, const double* verts, int64_t nv, const int64_t* faces, int64_t nf) {
  if (!path || nv < 0 || nf < 0 || (nv && !verts) || (nf && !faces)) return fail(DJ_ERR_INVALID, "dj_host_write_obj: bad argument");
  FILE* fh = fopen(path, "w");
  if (!fh) return fail(errno == EACCES ? DJ_ERR_PERMISSION : DJ_ERR_IO, "dj_host_write_obj: cannot open %s: %s", path, strerror(errno));
  std::vector<char> big(1 << 22);
  setvbuf(fh, big.data(), _IOFBF, big.size());
  bool ok = true;
  for (int64_t i = 0; i < nv && ok; i++) ok = fprintf(fh, "v %.8f %.8f %.8f\n", verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]) > 0;
  for (int64_t i = 0; i < nf && ok; i++)
    ok = fprintf(fh, "f %lld %lld %lld\n", (long long)faces[3 * i] + 1, (long long)faces[3 * i + 1] + 1, (long long)faces[3 * i + 2] + 1) > 0;
  const int cl = fclose(fh);
  if (!ok || cl != 0) return fail(DJ_ERR_IO, "dj_host_write_obj: short write to %s", path);
  return DJ_OK;
}

// (the host-side sampler on numpy's MT19937 stream lives in host_sampler.cpp)
extern "C" int dj_sample_schedule(const float* pool_sdf_dev, int64_t pool_n, int num_iterations, int num_samples,
                                  float uniform_ratio, uint64_t seed, int32_t* idx_dev, void* stream) {
  return sampler::sample_schedule(pool_sdf_dev, pool_n, num_iterations, num_samples, uniform_ratio, seed, idx_dev, (cudaStream_t)stream);
}

// ====================================================================== marching cubes
struct DjMcWork {
  int device = 0;
  mc::Dims D{};
  DevBuf records, counts, offs, edge, misc, verts, faces, merge, wide, volbuf;
  int64_t n_rec = 0;        // cut cells of the last dj_mc_count
  int arena_cap = 0;        // records per arena of the record list (mc::NARENA arenas)
  int span = 32;            // the fullest arena of the last count, rounded up to 32: what the emit kernels walk
  // host-facing create_mesh entry points (dj_create_mesh_count / fetch): this workspace's own stream, so that two
  // host threads meshing with one DjPack (each with its own DjMcWork) neither share per-shape tables nor order
  // against each other
  cudaStream_t own_stream = nullptr;
  double* v64 = nullptr;  // create_mesh: final float64 vertices on the device (inside verts / merge)
  int64_t nblocks = 0, nv = 0, nf = 0;
  int bpp = 0;  // blocks per cell layer (grid.x of the marching-cubes kernels; grid.y = cell layers)
  const float* vol = nullptr;
  bool counted = false;
  // create_mesh post-processing
  double spacing[3] = {1, 1, 1}, centre = 0;
};

extern "C" int dj_mc_create(int device, DjMcWork** out) {
  if (!out) return fail(DJ_ERR_INVALID, "dj_mc_create: null out");
  int ndev = 0;
  DJ_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(DJ_ERR_INVALID, "device %d out of range", device);
  DjMcWork* w = new DjMcWork();
  w->device = device;
  {
    DeviceGuard g(device);
    if (cudaStreamCreateWithFlags(&w->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete w;
      return fail(DJ_ERR_CUDA, "dj_mc_create: cannot create a stream on device %d", device);
    }
  }
  *out = w;
  return DJ_OK;
}
extern "C" int dj_mc_destroy(DjMcWork* w) {
  if (!w) return DJ_OK;
  DeviceGuard g(w->device);
  for (DevBuf* b : {&w->records, &w->counts, &w->offs, &w->edge, &w->misc, &w->verts, &w->faces, &w->merge, &w->wide, &w->volbuf}) b->release();
  if (w->own_stream) cudaStreamDestroy(w->own_stream);
  delete w;
  return DJ_OK;
}

extern "C" int dj_mc_count(DjMcWork* w, const float* vol_dev, int n0, int n1, int n2, int z_off, int seam, double level,
                           int check_range, int64_t* n_verts, int64_t* n_faces, void* stream) {
  if (!w || !vol_dev || !n_verts || !n_faces) return fail(DJ_ERR_INVALID, "dj_mc_count: null argument");
  if (n0 < 2 || n1 < 2 || n2 < 2) return fail(DJ_ERR_INVALID, "dj_mc_count: volume must be at least 2x2x2 (got %dx%dx%d)", n0, n1, n2);
  DeviceGuard g(w->device);
  cudaStream_t st = (cudaStream_t)stream;
  w->counted = false;
  mc::Dims& D = w->D;
  D.n0 = n0; D.n1 = n1; D.n2 = n2; D.z_off = z_off; D.seam = seam ? 1 : 0;
  D.c1 = n1 - 1; D.c2 = n2 - 1;
  D.ncells = (int64_t)(n0 - 1) * D.c1 * D.c2;
  D.level = level;
  {  // largest float32 <= level (classification compares float32 samples in float32, exactly as the double test would)
    float lf = (float)level;
    if ((double)lf > level) lf = nextafterf(lf, -INFINITY);
    if (lf == 0.f) lf = 0.f;  // +0: the kernel reads the sign bit of (level_lo - v)
    D.level_lo = lf;
  }
  w->vol = vol_dev;
  if (n0 - 1 > 65535 || (n1 - 1 + mc::ROWS - 1) / mc::ROWS > 65535) return fail(DJ_ERR_UNSUPPORTED, "dj_mc_count: more than 65535 cell layers / row groups");
  if (D.ncells >= 0x7FFFFFFFLL) return fail(DJ_ERR_UNSUPPORTED, "dj_mc_count: more than 2^31 - 1 cells in one slab");
  D.wpr = (D.c2 + mc::UC - 1) / mc::UC;
  const int bpr = (D.wpr + mc::CB / 32 - 1) / (mc::CB / 32);                 // blocks along x
  w->bpp = ((D.c1 + mc::ROWS - 1) / mc::ROWS) * bpr;                         // blocks per cell layer
  w->nblocks = (int64_t)(n0 - 1) * D.c1 * D.wpr;                             // scan units: 31-cell chunks in sweep order
  if (w->nblocks >= 0x7FFFFFFFLL) return fail(DJ_ERR_UNSUPPORTED, "dj_mc_count: too many scan units");
  DJ_CUDA(w->counts.reserve((size_t)w->nblocks * sizeof(uint2)));
  DJ_CUDA(w->offs.reserve((size_t)w->nblocks * sizeof(int2)));
  uint2* units = w->counts.as<uint2>();
  const int64_t ngroups = (w->nblocks + mc::SG - 1) / mc::SG;
  DJ_CUDA(w->misc.reserve(64 + mc::NARENA * sizeof(int) + (size_t)ngroups * sizeof(longlong2)));
  long long* totals = w->misc.as<long long>();     // [0] vertices [1] triangles [2] (min, max) ordered ints
  int* mm = reinterpret_cast<int*>(totals + 2);
  int* cursors = reinterpret_cast<int*>(totals + 8);  // [NARENA] records per arena
  longlong2* group_off = reinterpret_cast<longlong2*>(cursors + mc::NARENA);
  const dim3 mc_grid((unsigned)bpr, (unsigned)((D.c1 + mc::ROWS - 1) / mc::ROWS), (unsigned)(n0 - 1));
  // The record list (one entry per cut cell, ~3 % of the cells): NARENA arenas of arena_cap entries.  It starts at
  // 1/16 of the cells and grows on demand: the kernel counts past the capacity, the sweep is then repeated once
  // with arenas that fit (the workspace keeps the larger size).
  w->arena_cap = (int)std::max<int64_t>(std::max(w->arena_cap, 64), D.ncells / 16 / mc::NARENA);
  long long h[3];
  std::vector<int> hc(mc::NARENA);
  for (int attempt = 0; attempt < 2; attempt++) {
    DJ_CUDA(w->records.reserve((size_t)w->arena_cap * mc::NARENA * sizeof(mc::Record)));
    DJ_CUDA(cudaMemsetAsync(cursors, 0, mc::NARENA * sizeof(int), st));
    DJ_CUDA(cudaMemsetAsync(units, 0, (size_t)w->nblocks * sizeof(uint2), st));  // only units with cut cells are written
    mc::classify_kernel<<<mc_grid, mc::CB, 0, st>>>(D, vol_dev, units, w->records.as<mc::Record>(), w->arena_cap, cursors);
    DJ_LAUNCHED();
    mc::scan_groups_kernel<<<(unsigned)ngroups, mc::SG, 0, st>>>(units, w->nblocks, w->offs.as<int2>(), group_off);
    DJ_LAUNCHED();
    mc::scan_totals_kernel<<<1, mc::SG, 0, st>>>(group_off, ngroups, totals);
    DJ_LAUNCHED();
    DJ_CUDA(cudaMemcpyAsync(h, totals, sizeof h, cudaMemcpyDeviceToHost, st));
    DJ_CUDA(cudaMemcpyAsync(hc.data(), cursors, mc::NARENA * sizeof(int), cudaMemcpyDeviceToHost, st));
    DJ_CUDA(cudaStreamSynchronize(st));
    int most = 0;
    w->n_rec = 0;
    for (int a = 0; a < mc::NARENA; a++) { most = std::max(most, hc[a]); w->n_rec += hc[a]; }
    w->span = std::max(32, (most + 31) / 32 * 32);
    if (most <= w->arena_cap) break;
    if (attempt == 1) return fail(DJ_ERR_DEVICE, "dj_mc_count: record arena overflow after growing (%d > %d)", most, w->arena_cap);
    w->arena_cap = most + most / 4 + 64;
  }
  if (check_range && h[1] == 0) {
    // skimage's ValueError: level outside [min, max].  A cut cell has samples on both sides of the level, so the range
    // only has to be looked at when nothing was cut (then this pass over the volume decides which error it is)
    const int init[2] = {0x7FFFFFFF, (int)0x80000000};  // min, max (ordered ints)
    DJ_CUDA(cudaMemcpyAsync(mm, init, sizeof init, cudaMemcpyHostToDevice, st));
    mc::minmax_kernel<<<1184, 256, 0, st>>>(vol_dev, (int64_t)n0 * n1 * n2, mm);
    DJ_LAUNCHED();
    int hm[2];
    DJ_CUDA(cudaMemcpyAsync(hm, mm, sizeof hm, cudaMemcpyDeviceToHost, st));
    DJ_CUDA(cudaStreamSynchronize(st));
    auto unord = [](int i) { int b = i >= 0 ? i : i ^ 0x7FFFFFFF; float f; memcpy(&f, &b, 4); return f; };
    const float lo = unord(hm[0]), hi = unord(hm[1]);
    if (level < (double)lo || level > (double)hi)
      return fail(DJ_ERR_LEVEL_RANGE, "Surface level must be within volume data range. (level %g, range [%g, %g])", level, lo, hi);
  }
  w->nv = h[0];
  w->nf = h[1];
  if (w->nv >= 0x7FFFFFFFLL || w->nf >= 0x7FFFFFFFLL / 3) return fail(DJ_ERR_UNSUPPORTED, "mesh too large for int32 ids");
  *n_verts = w->nv;
  *n_faces = w->nf;
  if (w->nf == 0 && !seam) return fail(DJ_ERR_NO_SURFACE, "No surface found at the given iso value.");
  w->counted = true;
  return DJ_OK;
}

extern "C" int dj_mc_fill(DjMcWork* w, int descent, float* verts_dev, int32_t* faces_dev, void* stream) {
  if (!w || !w->counted) return fail(DJ_ERR_INVALID, "dj_mc_fill: call dj_mc_count first");
  if ((w->nv && !verts_dev) || (w->nf && !faces_dev)) return fail(DJ_ERR_INVALID, "dj_mc_fill: null output");
  DeviceGuard g(w->device);
  cudaStream_t st = (cudaStream_t)stream;
  const mc::Dims& D = w->D;
  const int64_t nsamp = (int64_t)D.n0 * D.n1 * D.n2;
  DJ_CUDA(w->edge.reserve((size_t)nsamp * mc::SLOTS * sizeof(int)));
  if (D.seam) {
    const int64_t np = (int64_t)D.n1 * D.n2;
    mc::seam_fill_kernel<<<ceil_div(np * 2, 256), 256, 0, st>>>(w->edge.as<int>(), np);
    DJ_LAUNCHED();
  }
  const int* cursors = reinterpret_cast<const int*>(w->misc.as<long long>() + 8);
  const longlong2* group_off = reinterpret_cast<const longlong2*>(cursors + mc::NARENA);
  if (w->n_rec > 0) {
    const unsigned grid = (unsigned)(((int64_t)w->span * mc::NARENA + mc::EB - 1) / mc::EB);
    mc::emit_vertices_kernel<<<grid, mc::EB, 0, st>>>(D, w->records.as<mc::Record>(), w->span, w->arena_cap, cursors, w->offs.as<int2>(), group_off,
                                                     w->edge.as<int>(), verts_dev);
    DJ_LAUNCHED();
    mc::emit_faces_kernel<<<grid, mc::EB, 0, st>>>(D, w->records.as<mc::Record>(), w->span, w->arena_cap, cursors, w->offs.as<int2>(), group_off,
                                                  w->edge.as<int>(), descent, faces_dev);
    DJ_LAUNCHED();
  }
  return DJ_OK;
}

extern "C" int dj_mc_top_plane(DjMcWork* w, int32_t* ids_dev, void* stream) {
  if (!w || !w->counted || !ids_dev) return fail(DJ_ERR_INVALID, "dj_mc_top_plane: bad argument");
  DeviceGuard g(w->device);
  const mc::Dims& D = w->D;
  const int64_t np = (int64_t)D.n1 * D.n2;
  mc::top_plane_kernel<<<ceil_div(np * 3, 256), 256, 0, (cudaStream_t)stream>>>(w->edge.as<int>() + (size_t)(D.n0 - 1) * np * mc::SLOTS, ids_dev, np);
  DJ_LAUNCHED();
  return DJ_OK;
}

// vertices of create_mesh: f32 index units -> f64, * spacing, swap columns 0/1, - centre
// (core/reconstruct.py:172,179-187; float64 multiply then subtract, no fused multiply-add)
__global__ void mesh_post_kernel(const float* __restrict__ v, double* __restrict__ out, int64_t nv, double s0, double s1,
                                 double s2, double centre) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv) return;
  out[3 * i + 0] = __dsub_rn(__dmul_rn((double)v[3 * i + 1], s1), centre);
  out[3 * i + 1] = __dsub_rn(__dmul_rn((double)v[3 * i + 0], s0), centre);
  out[3 * i + 2] = __dsub_rn(__dmul_rn((double)v[3 * i + 2], s2), centre);
}

// Device-side Mesh.merge_vertices (csrc/mesh_merge.cuh).  v64 [nv,3] is left untouched; when vertices
// merged, *vout points at the compacted copy inside w->merge and faces are re-indexed in place.
// Synchronises the stream (the caller needs the count).
static int merge_vertices_dev(DjMcWork* w, const double* v64, int64_t nv, int32_t* faces, int64_t nf, cudaStream_t st,
                              double** vout_p, int64_t* uniq_p) {
  *vout_p = const_cast<double*>(v64);
  *uniq_p = nv;
  if (nv == 0) return DJ_OK;
  if (nv > 0x7ffffff0ll) return fail(DJ_ERR_UNSUPPORTED, "mesh with %lld vertices", (long long)nv);
  uint32_t tsize = 1024;
  while ((int64_t)tsize < 2 * nv) tsize <<= 1;
  const int64_t nsb = ceil_div(nv, (int64_t)meshmerge::SCAN_BLOCK);
  // layout (int32 words): [table tsize][canon nv][newid nv][block sums nsb][counter][vout 3 nv doubles]
  const size_t off_canon = (size_t)tsize, off_newid = off_canon + nv, off_bs = off_newid + nv;
  const size_t off_cnt = (off_bs + nsb + 1) & ~(size_t)1, off_vout = off_cnt + 2;
  DJ_CUDA(w->merge.reserve(off_vout * sizeof(int32_t) + (size_t)nv * 3 * sizeof(double)));
  int32_t* mt = w->merge.as<int32_t>();
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(mt + off_cnt);
  meshmerge::fill_kernel<<<(unsigned)std::min<int64_t>(ceil_div((int64_t)tsize, (int64_t)256), 4096), 256, 0, st>>>(mt, tsize, meshmerge::EMPTY);
  DJ_LAUNCHED();
  DJ_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), st));
  meshmerge::insert_kernel<<<(unsigned)ceil_div(nv, (int64_t)256), 256, 0, st>>>(v64, nv, mt, tsize - 1);
  DJ_LAUNCHED();
  meshmerge::lookup_kernel<<<(unsigned)ceil_div(nv, (int64_t)256), 256, 0, st>>>(v64, nv, mt, tsize - 1, mt + off_canon, cnt);
  DJ_LAUNCHED();
  unsigned long long uniq = 0;
  DJ_CUDA(cudaMemcpyAsync(&uniq, cnt, sizeof uniq, cudaMemcpyDeviceToHost, st));
  DJ_CUDA(cudaStreamSynchronize(st));
  if ((int64_t)uniq == nv) return DJ_OK;
  double* vout = reinterpret_cast<double*>(mt + off_vout);
  meshmerge::block_count_kernel<<<(unsigned)nsb, meshmerge::SCAN_BLOCK, 0, st>>>(mt + off_canon, nv, mt + off_bs);
  DJ_LAUNCHED();
  meshmerge::scan_sums_kernel<<<1, meshmerge::SCAN_BLOCK, 0, st>>>(mt + off_bs, (int)nsb);
  DJ_LAUNCHED();
  meshmerge::compact_kernel<<<(unsigned)nsb, meshmerge::SCAN_BLOCK, 0, st>>>(v64, mt + off_canon, nv, mt + off_bs, mt + off_newid, vout);
  DJ_LAUNCHED();
  if (nf > 0) {
    meshmerge::remap_faces_kernel<<<(unsigned)ceil_div(nf * 3, (int64_t)256), 256, 0, st>>>(faces, nf * 3, mt + off_canon, mt + off_newid);
    DJ_LAUNCHED();
  }
  *vout_p = vout;
  *uniq_p = (int64_t)uniq;
  return DJ_OK;
}

/* replaces: trimesh's merge_vertices inside trimesh.Trimesh(process=True) (core/reconstruct.py:189). */
extern "C" int dj_mesh_merge_vertices(DjMcWork* w, double* verts_dev, int64_t n_verts, int32_t* faces_dev, int64_t n_faces,
                                      int64_t* n_verts_out, void* stream) {
  if (!w || !n_verts_out || n_verts < 0 || n_faces < 0 || (n_verts && !verts_dev) || (n_faces && !faces_dev))
    return fail(DJ_ERR_INVALID, "dj_mesh_merge_vertices: bad argument");
  DeviceGuard g(w->device);
  cudaStream_t st = (cudaStream_t)stream;
  double* vout = verts_dev;
  int64_t uniq = n_verts;
  int rc = merge_vertices_dev(w, verts_dev, n_verts, faces_dev, n_faces, st, &vout, &uniq);
  if (rc) return rc;
  if (vout != verts_dev) DJ_CUDA(cudaMemcpyAsync(verts_dev, vout, (size_t)uniq * 3 * sizeof(double), cudaMemcpyDeviceToDevice, st));
  *n_verts_out = uniq;
  return DJ_OK;
}

extern "C" int dj_create_mesh_count(DjPack* p, DjMcWork* w, const float* code_host, int d0, int d1, int d2, double padding,
                                    int use_net, int return_occ, int apply_sigmoid, double level, int descent, int precision,
                                    int64_t* n_verts, int64_t* n_faces) {
  if (!p || !w || !code_host || !n_verts || !n_faces) return fail(DJ_ERR_INVALID, "dj_create_mesh_count: null argument");
  if (p->device != w->device) return fail(DJ_ERR_INVALID, "pack and marching-cubes workspace live on different devices");
  DeviceGuard g(p->device);
  cudaStream_t st = w->own_stream;
  const int64_t tot = (int64_t)d0 * d1 * d2;
  const int nc = p->L + p->Lb;
  DJ_CUDA(w->volbuf.reserve((size_t)tot * sizeof(float) + 1024));
  float* code_dev = w->volbuf.as<float>();
  float* vol = code_dev + 256;
  DJ_CUDA(cudaMemcpyAsync(code_dev, code_host, nc * sizeof(float), cudaMemcpyHostToDevice, st));
  const double stop = 1.0 + (padding * 2), offs = 0.5 + padding;
  int rc = dj_query_grid(p, code_dev, d0, d1, d2, stop, offs, 0, tot, use_net, return_occ, apply_sigmoid, precision, vol, st);
  if (rc) return rc;
  // values.reshape(dims): the flat rows are reinterpreted as a C-contiguous [d0, d1, d2] volume (core/reconstruct.py:162)
  rc = dj_mc_count(w, vol, d0, d1, d2, 0, 0, level, 1, n_verts, n_faces, st);
  if (precision != DJ_PREC_FP32) {
    int rc2 = check_err_word(p, st);
    if (rc2) return rc2;
  }
  if (rc) return rc;
  DJ_CUDA(w->verts.reserve((size_t)w->nv * 3 * (sizeof(float) + sizeof(double)) + 16));
  DJ_CUDA(w->faces.reserve((size_t)w->nf * 3 * sizeof(int)));
  rc = dj_mc_fill(w, descent, w->verts.as<float>(), w->faces.as<int>(), st);
  if (rc) return rc;
  w->spacing[0] = (1 + (padding * 2)) / d0;
  w->spacing[1] = (1 + (padding * 2)) / d1;
  w->spacing[2] = (1 + (padding * 2)) / d2;
  w->centre = 0.5 + padding;
  // world-space float64 vertices, then the duplicate-vertex merge trimesh's process=True performs (:189)
  const int64_t nv = w->nv;
  double* v64 = reinterpret_cast<double*>(w->verts.as<float>() + (((size_t)nv * 3 + 1) & ~(size_t)1));
  mesh_post_kernel<<<ceil_div(nv, 256), 256, 0, st>>>(w->verts.as<float>(), v64, nv, w->spacing[0], w->spacing[1], w->spacing[2], w->centre);
  DJ_LAUNCHED();
  w->v64 = v64;
  int64_t uniq = nv;
  double* vfinal = v64;
  if ((rc = merge_vertices_dev(w, v64, nv, w->faces.as<int>(), w->nf, st, &vfinal, &uniq))) return rc;
  w->v64 = vfinal;
  w->nv = uniq;
  *n_verts = uniq;
  return DJ_OK;
}

extern "C" int dj_create_mesh_fetch(DjMcWork* w, double* verts_host, int32_t* faces_host) {
  if (!w || !w->counted || !w->v64 || !verts_host || !faces_host) return fail(DJ_ERR_INVALID, "dj_create_mesh_fetch: bad argument");
  DeviceGuard g(w->device);
  cudaStream_t st = w->own_stream;
  DJ_CUDA(cudaMemcpyAsync(verts_host, w->v64, (size_t)w->nv * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  DJ_CUDA(cudaMemcpyAsync(faces_host, w->faces.as<int>(), (size_t)w->nf * 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
  DJ_CUDA(cudaStreamSynchronize(st));
  return DJ_OK;
}

// faces widened to the int64 a trimesh.Trimesh holds, and a non-finite-vertex flag, both produced on the device so that the
// host wrapper neither converts nor scans the arrays (core/reconstruct.py:189: Trimesh(vertices, faces))
__global__ void mesh_widen_kernel(const int32_t* __restrict__ f32, int64_t* __restrict__ f64, int64_t nf3, const double* __restrict__ v,
                                  int64_t nv3, int* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nf3) f64[i] = (int64_t)f32[i];
  bool bad = false;
  if (i < nv3) bad = !isfinite(v[i]);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

extern "C" int dj_create_mesh_fetch64(DjMcWork* w, double* verts_host, int64_t* faces_host, int* nonfinite) {
  if (!w || !w->counted || !w->v64 || !verts_host || !faces_host || !nonfinite)
    return fail(DJ_ERR_INVALID, "dj_create_mesh_fetch64: bad argument");
  DeviceGuard g(w->device);
  cudaStream_t st = w->own_stream;
  const int64_t nf3 = w->nf * 3, nv3 = w->nv * 3;
  DJ_CUDA(w->wide.reserve((size_t)nf3 * sizeof(int64_t) + 16));
  int64_t* f64 = w->wide.as<int64_t>();
  int* flag = reinterpret_cast<int*>(f64 + nf3);
  DJ_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  const int64_t n = nf3 > nv3 ? nf3 : nv3;
  if (n) {
    mesh_widen_kernel<<<(unsigned)ceil_div(n, (int64_t)256), 256, 0, st>>>(w->faces.as<int>(), f64, nf3, w->v64, nv3, flag);
    DJ_LAUNCHED();
  }
  DJ_CUDA(cudaMemcpyAsync(verts_host, w->v64, (size_t)nv3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  DJ_CUDA(cudaMemcpyAsync(faces_host, f64, (size_t)nf3 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  DJ_CUDA(cudaMemcpyAsync(nonfinite, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  DJ_CUDA(cudaStreamSynchronize(st));
  return DJ_OK;
}

// ====================================================================== evaluation metrics: nearest neighbour
extern "C" int dj_nn_query(const float* queries_dev, int64_t n, const float* targets_dev, int64_t m, float* dist2_dev,
                           int32_t* idx_dev, void* stream) {
  if (n < 0 || m < 1 || (n && (!queries_dev || !dist2_dev || !idx_dev)) || !targets_dev || m > 0x7fffffffll)
    return fail(DJ_ERR_INVALID, "dj_nn_query: bad argument");
  if (n == 0) return DJ_OK;
  nn::nn_kernel<<<(unsigned)ceil_div(n, (int64_t)nn::QB), nn::QB, 0, (cudaStream_t)stream>>>(queries_dev, targets_dev, n, m, dist2_dev, idx_dev);
  DJ_LAUNCHED();
  return DJ_OK;
}

#ifdef DJ_DEBUG
// ====================================================================== debug probe (not product path)
// A_host [128,K], W_host [128,K] fp32 (nn.Linear layout); D_host [128,128] = bf16(A) . bf16(W)^T.
// desc_tmpl == 0 / k_adv_bytes == 0 / idesc == 0 select the encodings the product kernel uses.
extern "C" int dj_debug_read_prof(DjPack* p, unsigned long long* out_host, int n_words) {
  if (!p || !out_host || n_words < 0 || n_words > p->num_sms * 16) return fail(DJ_ERR_INVALID, "dj_debug_read_prof: bad argument");
  DeviceGuard g(p->device);
  DJ_CUDA(cudaDeviceSynchronize());
  DJ_CUDA(cudaMemcpy(out_host, p->prof, (size_t)n_words * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return DJ_OK;
}

extern "C" int dj_debug_umma_probe(const float* A_host, const float* W_host, int K, int use_bulk, int k_adv_bytes,
                                   uint32_t idesc, uint64_t desc_tmpl, float* D_host) {
  if (!A_host || !W_host || !D_host || K < 64 || K > 256 || K % 64) return fail(DJ_ERR_INVALID, "dj_debug_umma_probe: bad argument");
  float *A = nullptr, *D = nullptr;
  uint8_t* B = nullptr;
  unsigned int* err = nullptr;
  std::vector<uint8_t> blob((size_t)(K / 64) * 16384);
  for (int c = 0; c < K / 64; c++) pack_tile(W_host, K, 128, K, 0, c * 64, 128, blob.data() + (size_t)c * 16384);
  DJ_CUDA(cudaMalloc((void**)&A, 128 * K * sizeof(float)));
  DJ_CUDA(cudaMalloc((void**)&D, 128 * 128 * sizeof(float)));
  DJ_CUDA(cudaMalloc((void**)&B, blob.size()));
  DJ_CUDA(cudaMalloc((void**)&err, 4));
  DJ_CUDA(cudaMemset(err, 0, 4));
  DJ_CUDA(cudaMemset(D, 0xff, 128 * 128 * sizeof(float)));
  DJ_CUDA(cudaMemcpy(A, A_host, 128 * K * sizeof(float), cudaMemcpyHostToDevice));
  DJ_CUDA(cudaMemcpy(B, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  probe::ProbeParams P;
  P.A = A; P.Bblob = B; P.D = D; P.K = K; P.use_bulk = use_bulk;
  P.k_adv_bytes = k_adv_bytes ? k_adv_bytes : 32;
  P.idesc = idesc ? idesc : ptx::umma_idesc_bf16(128, 128);
  P.desc_tmpl = desc_tmpl ? desc_tmpl : (((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61));
  P.err = err;
  const int smem = 8 * 16384 + 1024 + 64;
  DJ_CUDA(cudaFuncSetAttribute(probe::umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe::umma_probe_kernel<<<1, 128, smem>>>(P);
  DJ_LAUNCHED();
  cudaError_t e = cudaDeviceSynchronize();
  unsigned int h = 0;
  if (e == cudaSuccess) cudaMemcpy(&h, err, 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(D_host, D, 128 * 128 * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(A); cudaFree(D); cudaFree(B); cudaFree(err);
  if (e != cudaSuccess) return fail(DJ_ERR_CUDA, "umma probe: %s (err word 0x%08x)", cudaGetErrorString(e), h);
  if (h) return fail(DJ_ERR_DEVICE, "umma probe watchdog 0x%08x", h);
  return DJ_OK;
}
#endif  // DJ_DEBUG
