// Output heads of the joint decoder: tanh / sigmoid and the closed-form Broken / Restoration
// combine.  replaces: networks/decoder_z_lb_both_leaky_n_sdf.py:278-453.
#pragma once
#include <cuda_runtime.h>

namespace dj {

enum { OUT_SINGLE = 0, OUT_ALL20 = 1 };

struct HeadCfg {
  int out_mode;       // OUT_SINGLE: one value per point; OUT_ALL20: the use_net=None tuple
  int use_net;        // 0 C, 1 B, 2 R, 3 T            (OUT_SINGLE)
  int return_occ;     // decoder :186-187
  int apply_sigmoid;  // decode_samples :127-128
  float slope;        // LeakyReLU slope (also used by the hidden layers)
};

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

// yc: fnet head pre-activations (no activation on lin8, decoder :229);
// yt: hnet head AFTER its LeakyReLU (decoder :259-270).
__device__ __forceinline__ void write_heads(const HeadCfg& H, const float (&yc)[5], const float (&yt)[5], float* out,
                                            long long g) {
  if (H.out_mode == OUT_SINGLE) {
    float v;
    if (H.return_occ) {
      if (H.use_net == 0) v = yc[1];                                         // logit (:353-354)
      else if (H.use_net == 3) v = yt[1];                                    // logit (:451-452)
      else {
        const float sc = sigmoidf_(yc[1]), st = sigmoidf_(yt[1]);
        v = (H.use_net == 1) ? sc * (1.0f - st) : sc * st;                   // :377, :419
      }
    } else {
      const float c = tanhf(yc[0]), t = tanhf(yt[0]);
      if (H.use_net == 0) v = c;
      else if (H.use_net == 3) v = t;
      else if (H.use_net == 1) v = fmaxf(c, -t);                             // eval: torch.max (:383-385)
      else v = fmaxf(c, t);                                                  // (:425-427)
    }
    if (H.apply_sigmoid) v = sigmoidf_(v);
    out[g] = v;
    return;
  }
  // use_net=None: select form even in eval (:299-344)
  const float c_sdf = tanhf(yc[0]), t_sdf = tanhf(yt[0]);
  const float sc = sigmoidf_(yc[1]), st = sigmoidf_(yt[1]);
  float cn[3], tn[3];
#pragma unroll
  for (int i = 0; i < 3; i++) { cn[i] = tanhf(yc[2 + i]); tn[i] = tanhf(yt[2 + i]); }
  const bool sel_b = (st > 0.5f) || (c_sdf < -t_sdf);
  const bool sel_r = (st <= 0.5f) || (c_sdf < t_sdf);
  float* o = out + g * 20;
  o[0] = yc[1];
  o[1] = sc * (1.0f - st);
  o[2] = sc * st;
  o[3] = yt[1];
  o[4] = c_sdf;
  o[5] = sel_b ? -t_sdf : c_sdf;
  o[6] = sel_r ? t_sdf : c_sdf;
  o[7] = t_sdf;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    o[8 + i] = cn[i];
    o[11 + i] = sel_b ? tn[i] : cn[i];
    o[14 + i] = sel_r ? -tn[i] : cn[i];
    o[17 + i] = tn[i];
  }
}

}  // namespace dj
