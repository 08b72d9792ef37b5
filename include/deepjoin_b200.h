/*
 * deepjoin_b200 -- C ABI of the B200 (sm_100a) DeepJoin inference hot path.
 *
 * Plain C: raw pointers + sizes, no torch types.  Every function returns an
 * int status (DJ_OK == 0) and never throws; dj_last_error() gives the message
 * of the last failure on the calling thread.  The Python host layer
 * (deepjoin_b200/_lib.py, ctypes) maps statuses to the reference's exception
 * types (core/errors.py).  All `*_dev` pointers are CUDA device pointers owned
 * by the caller; `stream` is a cudaStream_t passed as void* (0 = default).
 *
 * Each entry point names the reference interface it replaces
 * (paths relative to the reference's deepjoin/python/).
 */
#ifndef DEEPJOIN_B200_H
#define DEEPJOIN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DJ_ABI_VERSION 4

/* status codes */
#define DJ_OK 0
#define DJ_ERR_INVALID 1      /* bad argument                                     */
#define DJ_ERR_CUDA 2         /* CUDA runtime error (see dj_last_error)           */
#define DJ_ERR_UNSUPPORTED 3  /* architecture outside the supported configuration */
#define DJ_ERR_NO_SURFACE 4   /* skimage RuntimeError  -> IsosurfaceExtractionError (core/reconstruct.py:175) */
#define DJ_ERR_LEVEL_RANGE 5  /* skimage ValueError    -> IsosurfaceExtractionError (core/reconstruct.py:175) */
#define DJ_ERR_NO_SAMPLES 6   /* core/data.py:56-59,66-69 -> NoSamplesError       */
#define DJ_ERR_BAD_NET 7      /* decoder :454-457 RuntimeError("non-existent network") */
#define DJ_ERR_DEVICE 8       /* kernel-side watchdog / device fault flag         */
#define DJ_ERR_PERMISSION 9   /* host file not writable -> PermissionError (utils_multiprocessing.py:99-109) */
#define DJ_ERR_IO 10          /* other host file error -> OSError                 */

/* arithmetic of the hidden layers */
#define DJ_PREC_FP32 0   /* fp32 CUDA-core path: the reference's own arithmetic (torch fp32 SGEMM) */
#define DJ_PREC_BF16 1   /* bf16 operands, fp32 accumulation in TMEM, tcgen05 tensor cores         */
#define DJ_PREC_FP16 2   /* IEEE-half operands on the same kernels (forward only): same speed, 8x finer
                            operand rounding than bf16; the latent step treats it as DJ_PREC_BF16  */
#define DJ_PREC_FP16X2 3 /* split operands on the tensor cores: activations and weights as high + low IEEE-half
                            pairs, a_hi.W_hi + a_hi.W_lo + a_lo.W_hi accumulated in fp32 (3x the MMA work per
                            point in the forward entry points, 4x in the latent step).  The parity mode: ~1e-5 of
                            the fp64 decoder on trained-norm weights, where single-pass bf16 / half are at
                            5e-2 / 8e-3 */

/* use_net selector of Decoder.forward (networks/decoder_z_lb_both_leaky_n_sdf.py:345-453) */
#define DJ_NET_C 0
#define DJ_NET_B 1
#define DJ_NET_R 2
#define DJ_NET_T 3

typedef struct DjPack DjPack;     /* folded + packed decoder weights on one device */
typedef struct DjMcWork DjMcWork; /* marching-cubes workspace                      */

/* Architecture of Decoder.__init__ (networks/decoder_z_lb_both_leaky_n_sdf.py:15-152)
 * restricted to what experiments/<*>/specs.json ships: weight-normed ReLU-family MLPs,
 * one latent re-injection layer, xyz fed to the sub-net. */
typedef struct DjNetDesc {
  int32_t latent_size;       /* L   (CodeLength, 128)            */
  int32_t tool_latent_size;  /* Lb  (BreakCodeLength, 64)        */
  int32_t hidden;            /* 512 (must be 512)                */
  int32_t f_layers;          /* number of fnet linear layers (9) */
  int32_t h_layers;          /* number of hnet linear layers (6) */
  int32_t latent_in;         /* fnet layer whose input is [h ; z ; xyz] (4) */
  float   slope;             /* LeakyReLU negative slope (0.01; decoder :135) */
} DjNetDesc;

int dj_abi_version(void);
const char* dj_last_error(void);

/* ---- weights ------------------------------------------------------------------
 * replaces: Decoder.__init__ + load_state_dict + .cuda()  (core/utils_multiprocessing.py:16-30)
 * f_w[l] / h_w[l]: HOST pointers, weight-norm already folded (W = g*v/||v||_row), fp32
 * row-major [out,in] exactly as nn.Linear stores them; f_b / h_b the biases. */
int dj_pack_create(const DjNetDesc* desc, const float* const* f_w, const float* const* f_b,
                   const float* const* h_w, const float* const* h_b, int device, DjPack** out);
int dj_pack_destroy(DjPack* pack);
int dj_pack_device(const DjPack* pack);

/* ---- decoder forward ----------------------------------------------------------
 * replaces: Decoder.forward(net_input, use_net, return_occ)  (decoder :174-457), one
 * latent code for all rows (what decode_samples and reconstruct() feed it).
 * code_dev [L+Lb] f32; xyz_dev [n,3] f32; out_dev [n] f32.
 * apply_sigmoid: the `if sigmoid:` of decode_samples (core/reconstruct.py:127-128). */
int dj_query_points(DjPack* pack, const float* code_dev, const float* xyz_dev, int64_t n,
                    int use_net, int return_occ, int apply_sigmoid, int precision,
                    float* out_dev, void* stream);

/* use_net=None branch (decoder :299-344): out_dev [n,20] f32, columns
 * 0 c_occ(logit) 1 b_occ 2 r_occ 3 t_occ(logit) 4 c_sdf 5 b_sdf 6 r_sdf 7 t_sdf
 * 8-10 c_n 11-13 b_n 14-16 r_n 17-19 t_n */
int dj_query_all(DjPack* pack, const float* code_dev, const float* xyz_dev, int64_t n,
                 int precision, float* out_dev, void* stream);

/* replaces: decode_shape(..., reshape=False)  (core/reconstruct.py:45-86): the points of
 * np.meshgrid(lin_0, lin_1, lin_2) (default 'xy' indexing) are generated in-kernel from
 * the flat row index r in [r_lo, r_hi): r = (i*d0 + j)*d2 + k -> (lin_0[j], lin_1[i], lin_2[k]),
 * lin_a = np.linspace(0, stop, d_a) - offs evaluated in float64 then cast to f32.
 * out_dev [r_hi - r_lo] f32 (row r at out_dev[r - r_lo]). */
int dj_query_grid(DjPack* pack, const float* code_dev, int d0, int d1, int d2, double stop,
                  double offs, int64_t r_lo, int64_t r_hi, int use_net, int return_occ,
                  int apply_sigmoid, int precision, float* out_dev, void* stream);

/* replaces: decode_samples(vec, decoder, pts, use_net, batch_size, sigmoid)
 * (core/reconstruct.py:89-132) with HOST buffers: pts_host [n,3] float64 (as numpy hands
 * them over), out_host [n] float64.  Copies are chunked and overlapped inside. */
int dj_decode_samples_host(DjPack* pack, const float* code_host, const double* pts_host, int64_t n,
                           int use_net, int return_occ, int apply_sigmoid, int precision,
                           double* out_host);
/* same; element (i, d) of the points at pts_host[i*pt_stride + d*dim_stride] (in doubles): (3, 1) = rows, or
 * (1, S) = three coordinate arrays S apart -- the memory behind get_query_points' np.vstack(...).T
 * (core/reconstruct.py:25-30), which the reference slices row-wise (:113); no host-side transposition. */
int dj_decode_samples_host_strided(DjPack* pack, const float* code_host, const double* pts_host, int64_t n,
                                   int64_t pt_stride, int64_t dim_stride, int use_net, int return_occ,
                                   int apply_sigmoid, int precision, double* out_host);

/* ---- latent-code optimisation ---------------------------------------------------
 * replaces: the body of reconstruct()  (reconstruct.py:107-226) */
typedef struct DjLatentCfg {
  int32_t num_iterations;
  int32_t num_samples;      /* rows per iteration                                   */
  float   lr;               /* initial lr; x0.1 from iteration int(iters/2) (:38-46) */
  float   lambda1;          /* normal loss weight (:185-192)                         */
  float   lambda3;          /* sdf loss weight (:171-178)                            */
  float   clamp_dist;       /* :166                                                  */
  float   code_reg_lambda;  /* :197-201                                              */
  int32_t l2reg;
  float   beta1, beta2, eps;/* torch.optim.Adam defaults .9/.999/1e-8 (:58)          */
  int32_t log_every;        /* 10 (:203)                                             */
  int32_t precision;
  int32_t loss_kind;        /* DJ_LOSS_DEEPJOIN (0) | DJ_LOSS_DEEPSDF (1)            */
  /* continuation of a loop in pieces (the host sampler of the reference-exact path produces the schedule block by
   * block while the device works on the previous block): this call runs iterations iter_offset ..
   * iter_offset + num_iterations - 1 of a loop of total_iterations (0 = num_iterations) -- lr step, Adam bias
   * correction and log rows follow the GLOBAL iteration; adam_m / adam_v / codes / log carry over between calls,
   * idx_dev holds this call's num_iterations rows only.  With the tensor-core precisions a call that does not reach
   * total_iterations returns as soon as its iterations are enqueued (idx_dev must stay valid until the stream has
   * run them); the call that completes the loop synchronises the stream and reports a device fault of any block */
  int32_t iter_offset;
  int32_t total_iterations;
} DjLatentCfg;
/* DJ_LOSS_DEEPSDF: the baseline's objective (reconstruct_deepsdf.py:109-134) on the single-net pack a
 * DeepSDF decoder builds: mean |clamp(tanh(y0), +-d) - clamp(sdf, +-d)| + reg; lambda1 / lambda3 and the
 * normals are ignored, the second net is neither evaluated nor differentiated. */
#define DJ_LOSS_DEEPJOIN 0
#define DJ_LOSS_DEEPSDF 1

/* One forward+backward: loss terms [5] = (loss, sdf, occ, norm, reg) and dL/dcode [L+Lb].
 * xyz_dev [n,3], sdf_dev [n], nrm_dev [n,3] f32. */
int dj_latent_grad(DjPack* pack, const float* code_dev, const float* xyz_dev, const float* sdf_dev,
                   const float* nrm_dev, int64_t n, const DjLatentCfg* cfg, float* grad_dev,
                   float* loss_terms_dev, void* stream);

/* The whole Adam loop on the device.  The sample set (pool_* [pool_n, ...] f32) stays
 * resident; iteration e uses rows idx_dev[e*num_samples .. (e+1)*num_samples) (int32) --
 * the schedule core.data.select_samples would have produced (core/data.py:15-73), either
 * injected by the caller (parity) or drawn by dj_sample_schedule.
 * code_dev [L+Lb] in/out; adam_m/adam_v [L+Lb] in/out (zero them for a fresh run);
 * log_dev [ceil(iters/log_every), 8] = epoch, loss, sdf, occ, norm, reg, |z|, |z_b| (:203-211). */
int dj_latent_optimize(DjPack* pack, float* code_dev, float* adam_m_dev, float* adam_v_dev,
                       const float* pool_xyz_dev, const float* pool_sdf_dev, const float* pool_nrm_dev,
                       int64_t pool_n, const int32_t* idx_dev, const DjLatentCfg* cfg,
                       float* log_dev, void* stream);

/* The same loop for n_shapes (<= 64) independent shapes at once (what a reconstruct_chunk worker does item
 * after item, core/utils_multiprocessing.py:33-71): with DJ_PREC_BF16 every iteration is ONE launch over all
 * shapes' mini-batches, which fills the tile slots a single 30,000-sample batch leaves idle.
 * codes / adam_m / adam_v [n_shapes, L+Lb]; pool_*_dev, pool_n: HOST arrays of n_shapes device pointers / row
 * counts; idx_dev [n_shapes, iterations, num_samples]; log_dev [n_shapes, ceil(iters/log_every), 8]. */
int dj_latent_optimize_batch(DjPack* pack, int n_shapes, float* codes_dev, float* adam_m_dev, float* adam_v_dev,
                             const float* const* pool_xyz_dev, const float* const* pool_sdf_dev,
                             const float* const* pool_nrm_dev, const int64_t* pool_n, const int32_t* idx_dev,
                             const DjLatentCfg* cfg, float* log_dev, void* stream);

/* host-only: mesh export as the batch meshing worker does for every item (core/utils_multiprocessing.py:99-109,
 * core/handler.py:81-126 saver -> mesh.export): binary little-endian PLY (float32 vertices, uchar-counted int32
 * faces) / Wavefront OBJ ("v %.8f", 1-based "f").  verts [nv,3] float64, faces [nf,3] int64 (the arrays a mesh
 * holds).  No GPU. */
int dj_host_write_ply(const char* path, const double* verts, int64_t nv, const int64_t* faces, int64_t nf);
int dj_host_write_obj(const char* path, const double* verts, int64_t nv, const int64_t* faces, int64_t nf);

/* host-only helper of the reference-exact sampler (core/data.py:141-176): numpy's legacy
 * RandomState.permutation(n) -- arange(n) shuffled by mtrand's _shuffle_raw / random_interval -- continuing the
 * MT19937 state (key624, *pos) as np.random.get_state() returns it; *pos is updated.  out [n] int64.  No GPU. */
int dj_host_mt19937_permutation(uint32_t* key624, int32_t* pos, int64_t n, int64_t* out);

/* replaces: the per-iteration `core.data.select_samples(pts, values, num_samples, uniform_ratio)` of the latent loop
 * (reconstruct.py:114-118 -> core/data.py:15-73,122-176) for ALL iterations of one reconstruct() call, host side, on
 * numpy's global MT19937 stream: out [num_iterations][num_samples] int32 row indices, bit-identical to calling the
 * reference sampler num_iterations times from the state (key624, *pos) (np.random.get_state()); *pos / key624 are
 * left where numpy would be.  sdf [n_pts] float32 (one value column).  The calling thread walks the generator,
 * n_threads workers (0 = host cores - 1) finish the shuffles.  progress (optional): *progress = number of leading
 * iterations whose rows are final, updated as the call proceeds, so that another thread can consume them meanwhile.
 * DJ_ERR_NO_SAMPLES where the reference raises NoSamplesError.  No GPU. */
int dj_host_sample_schedule(uint32_t* key624, int32_t* pos, const float* sdf, int64_t n_pts, double uniform_ratio,
                            int num_iterations, int num_samples, int n_threads, int32_t* out, int32_t* progress);

/* replaces: core.data.select_samples(pts, values, num_samples, uniform_ratio)
 * (core/data.py:15-73,122-176) for every iteration at once, on the device, from a counter
 * based RNG (same distribution, not numpy's stream): keep the surface half, a random
 * ratio/(1-ratio) share of the uniform half, permute, then contiguous windows of
 * num_samples/2 positive and the rest non-positive rows.  DJ_ERR_NO_SAMPLES if a class is
 * empty.  idx_dev [num_iterations*num_samples] int32 out. */
int dj_sample_schedule(const float* pool_sdf_dev, int64_t pool_n, int num_iterations, int num_samples,
                       float uniform_ratio, uint64_t seed, int32_t* idx_dev, void* stream);

/* ---- marching cubes ---------------------------------------------------------------
 * replaces: skimage.measure.marching_cubes(values, level, gradient_direction) as called by
 * create_mesh (core/reconstruct.py:168-176).  vol_dev: C-contiguous f32 [n0,n1,n2].
 * Two phases so the caller allocates exact-size outputs:
 *   dj_mc_count  -> *n_verts, *n_faces   (synchronises `stream`)
 *   dj_mc_fill   -> verts_dev [n_verts,3] f32 in index units, columns (axis0,axis1,axis2);
 *                   faces_dev [n_faces,3] int32, ids in sequential-sweep first-use order.
 * descent != 0 reverses every face (skimage gradient_direction='descent').
 * Slab sharding (axis 0): vol_dev points at the slab's first plane, n0 = number of planes the slab
 * holds (its cell layers + 1), z_off = global index of that first plane (added to the axis-0
 * coordinate).  seam = 0: stand-alone volume.  seam = 1: this slab continues a sweep; the edges in
 * its plane 0 belong to the slab below, vertex ids are slab-local and faces referencing a seam edge
 * hold -2 - slot, slot = 3*(y*n2 + x) + dir (dir 0: axis-2 edge, 1: axis-1 edge); the slab below
 * exports its top-plane ids with dj_mc_top_plane so the gather step can patch them. */
int dj_mc_create(int device, DjMcWork** out);
int dj_mc_destroy(DjMcWork* work);
int dj_mc_count(DjMcWork* work, const float* vol_dev, int n0, int n1, int n2, int z_off, int seam,
                double level, int check_range, int64_t* n_verts, int64_t* n_faces, void* stream);
int dj_mc_fill(DjMcWork* work, int descent, float* verts_dev, int32_t* faces_dev, void* stream);
/* ids_dev [n1*n2*3] int32: vertex ids of the edges in the slab's LAST plane (valid where an edge
 * is intersected).  Call after dj_mc_fill. */
int dj_mc_top_plane(DjMcWork* work, int32_t* ids_dev, void* stream);

/* replaces: create_mesh() up to (not including) trimesh.Trimesh  (core/reconstruct.py:135-187)
 * with HOST outputs: grid query + marching cubes + spacing / column swap / centring in f64.
 * Phase 1 returns the counts, phase 2 copies into caller buffers. */
int dj_create_mesh_count(DjPack* pack, DjMcWork* work, const float* code_host, int d0, int d1, int d2,
                         double padding, int use_net, int return_occ, int apply_sigmoid, double level,
                         int descent, int precision, int64_t* n_verts, int64_t* n_faces);
int dj_create_mesh_fetch(DjMcWork* work, double* verts_host, int32_t* faces_host);
/* same, with faces widened on the device to the int64 a trimesh.Trimesh holds (core/reconstruct.py:189) and
 * *nonfinite = 1 when any vertex coordinate is NaN/Inf (what Trimesh(process=True) would have removed);
 * host buffers may be pinned (the Python wrapper takes them from torch's pinned-memory pool). */
int dj_create_mesh_fetch64(DjMcWork* work, double* verts_host, int64_t* faces_host, int* nonfinite);

/* replaces: the duplicate-vertex merge of trimesh.Trimesh(vertices, faces) (default process=True,
 * core/reconstruct.py:189): vertices whose coordinates agree after rounding to 8 decimals collapse onto
 * the one with the smallest index, survivors keep their order, faces are re-indexed.  In place on
 * device arrays (verts float64 [n_verts,3], faces int32 [n_faces,3]); synchronises the stream. */
int dj_mesh_merge_vertices(DjMcWork* work, double* verts_dev, int64_t n_verts, int32_t* faces_dev, int64_t n_faces,
                           int64_t* n_verts_out, void* stream);

/* replaces: scipy.spatial.cKDTree(targets).query(queries) of the evaluation metrics that follow meshing
 * (core/metrics.py:46-50,79): exact nearest neighbour by brute force on the current device.
 * queries [n,3], targets [m,3] f32; dist2 [n] squared distances, idx [n] int32 (smallest index on ties). */
int dj_nn_query(const float* queries_dev, int64_t n, const float* targets_dev, int64_t m, float* dist2_dev,
                int32_t* idx_dev, void* stream);

#ifdef DJ_DEBUG
/* The two entry points below exist only in a library built with -DDJ_DEBUG (python -m deepjoin_b200.build --debug,
 * which writes lib/libdeepjoin_b200_debug.so): diagnostics for tools/, not part of the product library. */
/* DEBUG ONLY (tools/gpu_diag.py): one 128x128xK bf16 UMMA built from the product kernel's pieces;
 * D_host [128,128] = bf16(A_host [128,K]) . bf16(W_host [128,K])^T.  Zero for k_adv_bytes / idesc /
 * desc_tmpl selects the product encodings. */
int dj_debug_umma_probe(const float* A_host, const float* W_host, int K, int use_bulk, int k_adv_bytes,
                        uint32_t idesc, uint64_t desc_tmpl, float* D_host);

/* DEBUG ONLY: per-CTA cycle counters of the last tensor-core forward launched with
 * DEEPJOIN_B200_DEBUG & 8 ([CTA][16] words; layout in csrc/mlp_tc2.cuh). */
int dj_debug_read_prof(DjPack* pack, unsigned long long* out_host, int n_words);
#endif /* DJ_DEBUG */

/* counters for bench.py: kernels launched by this library since the last reset */
int64_t dj_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif
