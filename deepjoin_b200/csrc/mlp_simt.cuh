// fp32 CUDA-core kernels: the reference's own arithmetic (torch fp32 SGEMM + elementwise),
// used as the DJ_PREC_FP32 path of the forward and for the latent-gradient backward, plus the
// small per-shape kernels every precision shares (latent fold, grid points, heads, Adam).
//
// replaces: networks/decoder_z_lb_both_leaky_n_sdf.py:207-276 (ff_fnet / ff_hnet) layer by layer,
//           reconstruct.py:147-224 (loss, backward, Adam step).
#pragma once
#include "common.cuh"
#include "heads.cuh"

namespace dj {
namespace simt {

// ------------------------------------------------------------------ latent fold
// tables[t][j] = (w_x, w_y, w_z, b[j] + sum_i W[j, lat_off + i] * code[code_off + i])
// t = 0: fnet lin0, 1: fnet lin<latent_in>, 2: hnet lin0.   One warp per (t, j).
struct FoldSrc {
  const float* W;   // [512, ld]
  const float* b;   // [512]
  int ld;
  int lat_off;      // first latent column
  int lat_n;        // number of latent columns
  int xyz_off;      // first xyz column
  int code_off;     // offset into the code vector
};
struct FoldParams {
  FoldSrc src[3];
};

// blockIdx.y = shape of a batch: code + y*code_stride -> tables + y*3*512 (strides 0 / grid.y 1 for one shape)
__global__ void fold_kernel(FoldParams P, const float* __restrict__ code, float4* __restrict__ tables, int code_stride = 0) {
  code += (size_t)blockIdx.y * code_stride;
  tables += (size_t)blockIdx.y * 3 * 512;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= 3 * 512) return;
  const int t = w / 512, j = w % 512;
  const FoldSrc S = P.src[t];
  const float* row = S.W + (size_t)j * S.ld;
  float acc = 0.f;
  for (int i = lane; i < S.lat_n; i += 32) acc = fmaf(row[S.lat_off + i], code[S.code_off + i], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) tables[t * 512 + j] = make_float4(row[S.xyz_off], row[S.xyz_off + 1], row[S.xyz_off + 2], acc + S.b[j]);
}

// ------------------------------------------------------------------ grid points
// replaces np.meshgrid(...)/vstack of decode_shape (core/reconstruct.py:68-71), rows [r_lo, r_lo+n)
__global__ void grid_points_kernel(float* __restrict__ xyz, int64_t n, int64_t r_lo, int d0, int d1, int d2,
                                   double step0, double step1, double step2, double stop, double offs) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const int64_t r = r_lo + g;
  const int k = (int)(r % d2);
  const int64_t q = r / d2;
  const int j = (int)(q % d0);
  const int i = (int)(q / d0);
  auto lin = [&](int t, int d, double step) {
    double v = (d > 1 && t == d - 1) ? stop : __dmul_rn((double)t, step);  // np.linspace(0, stop, 1) is [0.0]
    return (float)__dsub_rn(v, offs);
  };
  xyz[3 * g + 0] = lin(j, d0, step0);
  xyz[3 * g + 1] = lin(i, d1, step1);
  xyz[3 * g + 2] = lin(k, d2, step2);
}

__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) out[g] = (float)in[g];
}
// points stored as three coordinate arrays [3][m] (the memory behind np.vstack(...).T) -> [m,3] float32
__global__ void soa_f64_to_xyz_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int64_t m) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < 3 * m) {
    const int64_t i = g / 3;
    const int d = (int)(g - 3 * i);
    out[g] = (float)in[(int64_t)d * m + i];
  }
}
__global__ void f32_to_f64_kernel(const float* __restrict__ in, double* __restrict__ out, int64_t n) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) out[g] = (double)in[g];
}

// copy xyz into columns [col, col+3) of a row-major activation buffer
__global__ void put_xyz_cols_kernel(const float* __restrict__ xyz, float* __restrict__ buf, int ld, int col, int64_t n) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  buf[g * ld + col + 0] = xyz[3 * g + 0];
  buf[g * ld + col + 1] = xyz[3 * g + 1];
  buf[g * ld + col + 2] = xyz[3 * g + 2];
}

// ------------------------------------------------------------------ SGEMM
// C[M,N] = epi( A[M,K] * op(B) ),  op(B) = B^T for TRANSB (B is [N,K] row-major, nn.Linear layout)
//                                   op(B) = B   otherwise  (B is [K,N] row-major)
// epi: + bias[n*bias_stride]; LeakyReLU if act; * phi'(dsrc[m, n]) if dsrc (phi' = 1 if >0 else slope)
constexpr int BM = 128, BN = 128, BK = 16;

template <bool TRANSB>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B,
                                                    int ldb, float* __restrict__ C, int ldc, int M, int N, int K,
                                                    const float* __restrict__ bias, int bias_stride, int act,
                                                    float slope, const float* __restrict__ dsrc, int ldd) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // A tile: 128 rows x 16 k; thread loads 8 consecutive k of one row
    {
      const int r = t >> 1, kk = (t & 1) * 8;
      const int gm = m0 + r;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int gk = k0 + kk + i;
        As[kk + i][r] = (gm < M && gk < K) ? A[(size_t)gm * lda + gk] : 0.f;
      }
    }
    if (TRANSB) {
      const int r = t >> 1, kk = (t & 1) * 8;
      const int gn = n0 + r;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int gk = k0 + kk + i;
        Bs[kk + i][r] = (gn < N && gk < K) ? B[(size_t)gn * ldb + gk] : 0.f;
      }
    } else {
      const int kk = t >> 4, nn = (t & 15) * 8;
      const int gk = k0 + kk;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int gn = n0 + nn + i;
        Bs[kk][nn + i] = (gk < K && gn < N) ? B[(size_t)gk * ldb + gn] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; k++) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[(size_t)gn * bias_stride];
      if (act) v = v > 0.f ? v : v * slope;
      if (dsrc) v *= (dsrc[(size_t)gm * ldd + gn] > 0.f) ? 1.0f : slope;
      C[(size_t)gm * ldc + gn] = v;
    }
  }
}

// ------------------------------------------------------------------ heads from stored head rows
__global__ void heads_kernel(HeadCfg H, const float* __restrict__ yC, const float* __restrict__ yT, int ld,
                             float* __restrict__ out, int64_t n) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  float yc[5] = {0, 0, 0, 0, 0}, yt[5] = {0, 0, 0, 0, 0};
  if (yC)
#pragma unroll
    for (int i = 0; i < 5; i++) yc[i] = yC[g * ld + i];
  if (yT)
#pragma unroll
    for (int i = 0; i < 5; i++) yt[i] = yT[g * ld + i];
  write_heads(H, yc, yt, out, g);
}

// ------------------------------------------------------------------ loss + head gradients
// replaces reconstruct.py:166-201 and autograd through the heads / B select (decoder :299-313).
// loss_acc[0..2] += (sdf, occ, norm) partial sums (already scaled by lambda / N).
struct LossCfg {
  float lambda1, lambda3, clamp_dist, slope;
  float inv_n;
  int kind;  // DJ_LOSS_DEEPJOIN | DJ_LOSS_DEEPSDF
};

// Loss of ONE sample and its gradient wrt the 2 x 5 head pre-activations (reconstruct.py:166-192):
// yc = fnet head pre-activations, yt = hnet head outputs AFTER its LeakyReLU (decoder :259-270).
// gc / gt: d loss / d (head pre-activations); l3 += (sdf, occ, normal) loss terms of this sample.
__device__ __forceinline__ void loss_row(const LossCfg& L, const float (&yc)[5], const float (&yt)[5], float s,
                                         const float (&n_gt)[3], float (&gc)[5], float (&gt)[5], float (&l3)[3]) {
  if (L.kind == 1) {
    // baseline DeepSDF objective (reconstruct_deepsdf.py:109-127): L1 between the clamped prediction
    // tanh(y0) (decoder_deepsdf.py:103-104) and the clamped target; only the first head row exists
    const float pred = tanhf(yc[0]);
    const float diff = fminf(fmaxf(pred, -L.clamp_dist), L.clamp_dist) - fminf(fmaxf(s, -L.clamp_dist), L.clamp_dist);
    l3[0] += L.inv_n * fabsf(diff);
    const float sg = (diff > 0.f) ? 1.f : (diff < 0.f ? -1.f : 0.f);
    const bool pass = (pred >= -L.clamp_dist) && (pred <= L.clamp_dist);
#pragma unroll
    for (int i = 0; i < 5; i++) { gc[i] = 0.f; gt[i] = 0.f; }
    gc[0] = pass ? L.inv_n * sg * (1.f - pred * pred) : 0.f;
    return;
  }
  const float c_sdf = tanhf(yc[0]), t_sdf = tanhf(yt[0]);
  const float sc = sigmoidf_(yc[1]), st = sigmoidf_(yt[1]);
  float cn[3], tn[3];
#pragma unroll
  for (int i = 0; i < 3; i++) { cn[i] = tanhf(yc[2 + i]); tn[i] = tanhf(yt[2 + i]); }
  const bool sel_b = (st > 0.5f) || (c_sdf < -t_sdf);
  const float b_sdf = sel_b ? -t_sdf : c_sdf;
  const float occ = (s <= 0.f) ? 1.f : 0.f;  // tensor_sdf_to_occ (core/reconstruct.py:33-34)
  // sdf: lambda3 * mean |clamp(b_sdf) - s|
  float d_bsdf = 0.f;
  if (L.lambda3 != 0.f) {
    const float cl = fminf(fmaxf(b_sdf, -L.clamp_dist), L.clamp_dist);
    const float diff = cl - s;
    l3[0] += L.lambda3 * L.inv_n * fabsf(diff);
    const float sg = (diff > 0.f) ? 1.f : (diff < 0.f ? -1.f : 0.f);
    const bool pass = (b_sdf >= -L.clamp_dist) && (b_sdf <= L.clamp_dist);  // torch.clamp backward: closed interval
    d_bsdf = pass ? L.lambda3 * L.inv_n * sg : 0.f;
  }
  // occ: BCELoss(b_occ, occ); log clamped at -100, backward (p - o) / max(p (1-p), 1e-12)
  const float p = sc * (1.f - st);
  {
    const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.f - p), -100.f);
    l3[1] += -L.inv_n * (occ * lp + (1.f - occ) * l1p);
  }
  const float d_p = L.inv_n * (p - occ) / fmaxf(p * (1.f - p), 1e-12f);
  // normals: lambda1 * mean over 3N of (b_n - m)^2
  float d_bn[3] = {0, 0, 0};
  if (L.lambda1 != 0.f) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const float bn = sel_b ? tn[i] : cn[i];
      const float diff = bn - n_gt[i];
      l3[2] += L.lambda1 * L.inv_n * (1.f / 3.f) * diff * diff;
      d_bn[i] = L.lambda1 * L.inv_n * (2.f / 3.f) * diff;
    }
  }
  // route through the select and the head non-linearities
  gc[0] = (sel_b ? 0.f : d_bsdf) * (1.f - c_sdf * c_sdf);
  gt[0] = (sel_b ? -d_bsdf : 0.f) * (1.f - t_sdf * t_sdf);
  gc[1] = d_p * (1.f - st) * sc * (1.f - sc);
  gt[1] = -d_p * sc * st * (1.f - st);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    gc[2 + i] = (sel_b ? 0.f : d_bn[i]) * (1.f - cn[i] * cn[i]);
    gt[2 + i] = (sel_b ? d_bn[i] : 0.f) * (1.f - tn[i] * tn[i]);
  }
#pragma unroll
  for (int i = 0; i < 5; i++) gt[i] *= (yt[i] > 0.f ? 1.f : L.slope);  // LeakyReLU on the hnet head (decoder :259-270)
}

__global__ void loss_grad_kernel(LossCfg L, const float* __restrict__ yC, const float* __restrict__ yT, int ld,
                                 const float* __restrict__ sdf_gt, const float* __restrict__ n_gt,
                                 float* __restrict__ dyC, float* __restrict__ dyT, float* __restrict__ loss_acc,
                                 int64_t n) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float l3[3] = {0.f, 0.f, 0.f};
  if (g < n) {
    float yc[5], yt[5], gc[5], gt[5];
#pragma unroll
    for (int i = 0; i < 5; i++) { yc[i] = yC[g * ld + i]; yt[i] = yT[g * ld + i]; }
    const float m[3] = {n_gt[3 * g], n_gt[3 * g + 1], n_gt[3 * g + 2]};
    loss_row(L, yc, yt, sdf_gt[g], m, gc, gt, l3);
#pragma unroll
    for (int i = 0; i < 5; i++) {
      dyC[g * ld + i] = gc[i];
      dyT[g * ld + i] = gt[i];
    }
  }
  float l_sdf = l3[0], l_occ = l3[1], l_n = l3[2];
  // block reduction -> 3 atomics per block
  __shared__ float red[3][8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    l_sdf += __shfl_xor_sync(0xffffffffu, l_sdf, o);
    l_occ += __shfl_xor_sync(0xffffffffu, l_occ, o);
    l_n += __shfl_xor_sync(0xffffffffu, l_n, o);
  }
  if (lane == 0) { red[0][w] = l_sdf; red[1][w] = l_occ; red[2][w] = l_n; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) s += red[threadIdx.x][i];
    atomicAdd(loss_acc + threadIdx.x, s);
  }
}

// column sums over the points: out[c] += sum_m G[m, c]   (grid: (ceil(ncol/32), nblk_rows), block 32x8)
__global__ void colsum_kernel(const float* __restrict__ G, int ld, int64_t n, int ncol, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < ncol)
    for (int64_t m = (int64_t)blockIdx.y * 8 + threadIdx.y; m < n; m += (int64_t)gridDim.y * 8) s += G[m * ld + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < ncol) {
    for (int i = 1; i < 8; i++) s += red[i][threadIdx.x];
    atomicAdd(out + c, s);
  }
}

// dL/dcode[i] = sum_j colsum_t[j] * W_t[j, lat_off + i] (+ regulariser 2*lambda*code[i]/len)
struct GradFinishParams {
  FoldSrc src[3];          // same descriptors as the fold
  const float* colsum[3];  // [512] each: dPre of fnet lin0, fnet lin<latent_in>, hnet lin0
  int L, Lb;
  float reg_lambda;        // 0 when !l2reg
};
// blockIdx.y = shape of a batch: accumulators (colsum, loss_acc) + y*acc_stride, code + y*code_stride,
// grad + y*grad_stride (all strides 0 with grid.y 1 for one shape)
__global__ void grad_finish_kernel(GradFinishParams P, const float* __restrict__ code, float* __restrict__ grad,
                                   float* __restrict__ loss_acc, int acc_stride = 0, int code_stride = 0, int grad_stride = 0) {
  const int i = blockIdx.x;  // code element
  const int tot = P.L + P.Lb;
  if (i >= tot) return;
  code += (size_t)blockIdx.y * code_stride;
  grad += (size_t)blockIdx.y * grad_stride;
  loss_acc += (size_t)blockIdx.y * acc_stride;
  float acc = 0.f;
  for (int t = 0; t < 3; t++) {
    const FoldSrc S = P.src[t];
    const int li = i - S.code_off;
    if (li < 0 || li >= S.lat_n) continue;
    const float* cs = P.colsum[t] + (size_t)blockIdx.y * acc_stride;
    for (int j = threadIdx.x; j < 512; j += blockDim.x) acc = fmaf(cs[j], S.W[(size_t)j * S.ld + S.lat_off + li], acc);
  }
  __shared__ float red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w];
    const float len = (i < P.L) ? (float)P.L : (float)P.Lb;
    const float c = code[i];
    s += 2.f * P.reg_lambda * c / len;
    grad[i] = s;
    if (P.reg_lambda != 0.f) atomicAdd(loss_acc + 3, P.reg_lambda * c * c / len);
  }
}

// loss_terms[5] = (total, sdf, occ, norm, reg) from the accumulators
__global__ void loss_finish_kernel(const float* __restrict__ loss_acc, float* __restrict__ terms, int acc_stride = 0,
                                   int terms_stride = 0) {
  loss_acc += (size_t)blockIdx.x * acc_stride;  // blockIdx.x = shape of a batch
  terms += (size_t)blockIdx.x * terms_stride;
  if (threadIdx.x == 0) {
    const float s = loss_acc[0], o = loss_acc[1], nn = loss_acc[2], r = loss_acc[3];
    terms[0] = s + o + nn + r;
    terms[1] = s; terms[2] = o; terms[3] = nn; terms[4] = r;
  }
}

// torch.optim.Adam single step (no weight decay / amsgrad), lr schedule of reconstruct.py:38-46
struct AdamCfg {
  float lr, beta1, beta2, eps;
  int num_iterations;
  int mag_after_step;
};
// blockIdx.x = shape of a batch (code / m / v + x*code_stride, grad + x*grad_stride, terms + x*terms_stride,
// log_row + x*log_stride)
__global__ void adam_kernel(AdamCfg A, int e, float* __restrict__ code, float* __restrict__ m, float* __restrict__ v,
                            const float* __restrict__ grad, int n, const float* __restrict__ terms,
                            float* __restrict__ log_row, int L, int code_stride = 0, int grad_stride = 0, int terms_stride = 0,
                            long long log_stride = 0) {
  code += (size_t)blockIdx.x * code_stride;
  m += (size_t)blockIdx.x * code_stride;
  v += (size_t)blockIdx.x * code_stride;
  grad += (size_t)blockIdx.x * grad_stride;
  terms += (size_t)blockIdx.x * terms_stride;
  if (log_row) log_row += (size_t)blockIdx.x * log_stride;
  const int i = threadIdx.x;
  __shared__ float sq[2];
  if (i < 2) sq[i] = 0.f;
  __syncthreads();
  float c_old = 0.f, c_new = 0.f;
  if (i < n) {
    const int half = A.num_iterations / 2;
    float lr = A.lr;
    if (half > 0) {
      const int k = e / half;
      for (int q = 0; q < k; q++) lr *= 0.1f;
    }
    const float g = grad[i];
    const float mi = A.beta1 * m[i] + (1.f - A.beta1) * g;
    const float vi = A.beta2 * v[i] + (1.f - A.beta2) * g * g;
    m[i] = mi;
    v[i] = vi;
    const float step = (float)(e + 1);
    const float bc1 = 1.f - powf(A.beta1, step), bc2 = 1.f - powf(A.beta2, step);
    const float denom = sqrtf(vi) / sqrtf(bc2) + A.eps;
    c_old = code[i];
    c_new = c_old - (lr / bc1) * mi / denom;
    code[i] = c_new;
  }
  if (!log_row) return;
  // |z|, |z_b|: DeepJoin's loop logs them BEFORE the step (reconstruct.py:210-211), the DeepSDF baseline's AFTER
  // optimizer.step() (reconstruct_deepsdf.py:134-141)
  const float c = A.mag_after_step ? c_new : c_old;
  if (i < n) atomicAdd(&sq[i < L ? 0 : 1], c * c);
  __syncthreads();
  if (i == 0) {
    log_row[0] = (float)e;
    for (int k = 0; k < 5; k++) log_row[1 + k] = terms[k];
    log_row[6] = sqrtf(sq[0]);
    log_row[7] = sqrtf(sq[1]);
  }
}

// batch gather from the resident sample pool
__global__ void gather_kernel(const float* __restrict__ pxyz, const float* __restrict__ psdf,
                              const float* __restrict__ pnrm, const int32_t* __restrict__ idx, int n,
                              float* __restrict__ xyz, float* __restrict__ sdf, float* __restrict__ nrm) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const int64_t s = idx[g];
#pragma unroll
  for (int i = 0; i < 3; i++) { xyz[3 * g + i] = pxyz[3 * s + i]; nrm[3 * g + i] = pnrm[3 * s + i]; }
  sdf[g] = psdf[s];
}

}  // namespace simt
}  // namespace dj
