/* spnerf_b200.h — C ABI of libspnerf_sm100a.so
 *
 * The B200-native replacement for the PyTorch op groups that SP-NeRF's ray-rendering hot path
 * launches (reference: modules/rendering.py:119-218 render_rays, models/spnerf.py:63-159
 * inference, models/spnerf.py:273-369 SPNeRF.forward, modules/metrics.py:17-183 losses).
 * The reference has no FFI of its own: its "plugin API" is the Python call surface, which
 * sp-nerf_b200/{modules,models} mirrors; those Python mirrors bind exactly the entry points
 * declared here (ctypes).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in _host; the library never
 *    allocates, frees or retains caller memory (workspaces are caller-provided);
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *  - return value: 0 on success, SPNERF_ERR_* (> 0) for argument errors, -cudaError_t for
 *    CUDA runtime errors at launch time;
 *  - all arrays are dense row-major; "rays" is the reference's (B,11) fp32 layout
 *    [origin(3) direction(3) near far sun_dir(3)] (datasets/satellite_scene.py:577-592).
 */
#ifndef SPNERF_B200_H
#define SPNERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPNERF_ABI_VERSION 1

enum {
  SPNERF_OK = 0,
  SPNERF_ERR_BAD_ARG = 1,
  SPNERF_ERR_UNSUPPORTED = 2,
  SPNERF_ERR_WORKSPACE = 3
};

/* library / device probes (no compute) */
int spnerf_abi_version(void);
/* nonzero if a bounded device-side wait expired since load (debug aid; 0 in normal operation) */
unsigned int spnerf_watchdog_code(void);
/* sizeof() of every argument struct below, in declaration order (binding self-check) */
void spnerf_struct_sizes(int32_t* out10_host);

/* ---------------------------------------------------------------------------------------------
 * Tensor-core bring-up check (tests only): D[128,n] = A * B^T from caller-built shared-memory
 * images and caller-built UMMA descriptors.  No reference counterpart.
 * ------------------------------------------------------------------------------------------- */
#define SPNERF_SELFTEST_MAX_KSTEPS 64
typedef struct SpnerfUmmaSelftest {
  const void* a_img;        /* bytes copied verbatim to shared memory (1024-B aligned there) */
  const void* b_img;
  float* d_out;             /* [128][n] */
  uint32_t a_bytes, b_bytes;
  uint32_t n;               /* 16..256, multiple of 16 */
  uint32_t ksteps;          /* number of K=16 MMA instructions */
  uint32_t idesc;           /* instruction descriptor */
  uint32_t _pad;
  uint64_t a_desc_template; /* matrix descriptor without the start-address field */
  uint64_t b_desc_template;
  uint32_t a_off[SPNERF_SELFTEST_MAX_KSTEPS]; /* byte offset of step k inside the image */
  uint32_t b_off[SPNERF_SELFTEST_MAX_KSTEPS];
} SpnerfUmmaSelftest;
int spnerf_selftest_umma(const SpnerfUmmaSelftest* args, void* stream);
/* CTA-pair form (cluster of 2, cta_group::2, M = 256): a_img holds the two CTAs' A images (a_bytes
 * each), b_img the two halves of the B rows (b_bytes each), idesc says M = 256, d_out is [256][n]. */
int spnerf_selftest_umma2(const SpnerfUmmaSelftest* args, void* stream);
/* Profiling aid: when non-NULL, the next spnerf_mlp_fwd / spnerf_mlp_bwd_data launches log clock64()
 * stamps of CTA 0 into dev_buf2048 (2048 int64, device memory): [0,256) epilogue phase stamps,
 * [256,262) issuer wait sums, and for the third tile pair per step i < 256: [512+i] producer saw the
 * ring stage empty, [768+i] issuer started waiting for the stage, [1024+i] issuer saw the stage full,
 * [1280+i] commit issued.  NULL disables. */
void spnerf_debug_phase_clocks_fwd(long long* dev_buf2048);
void spnerf_debug_phase_clocks_bwd(long long* dev_buf2048);

/* ---------------------------------------------------------------------------------------------
 * Point network (replaces models/spnerf.py:162-369 SPNeRF.__init__/forward as executed through
 * models/spnerf.py:83-113, i.e. the chunked per-point MLP call of inference()).
 * ------------------------------------------------------------------------------------------- */
typedef struct SpnerfNetConfig {
  int32_t feat;            /* fc_units: 512 (modules/opt.py:43) or 256 (the class default)    */
  int32_t layers;          /* fc_layers; only 8 (modules/opt.py:45)                           */
  int32_t skip_layer;      /* 4 (models/spnerf.py:164 skips=[4])                              */
  int32_t mapping;         /* 1: 10-frequency positional encoding (models/spnerf.py:5-37)     */
  int32_t sem;             /* 1: label embedding input + semantic head                        */
  int32_t num_sem_classes; /* C <= 8                                                          */
  int32_t emb_dim;         /* C * s_embedding_factor; encoded input beyond 64 columns rides in free aux columns */
  int32_t beta;            /* 1: uncertainty head                                             */
  int32_t t_dim;           /* t_embbeding_tau <= 8                                            */
  int32_t relu;            /* 0: SIREN activations (siren=True, what load_model builds); 1: ReLU (models/spnerf.py:178) */
} SpnerfNetConfig;

/* Parameter slots, in the reference's state_dict order. */
enum {
  SPNERF_P_SEM_EMB = 0,          /* semantic_embedding.weight (C+1, emb_dim)            */
  SPNERF_P_FC_W0 = 1,            /* fc_net.{2i}.weight at 1+2i, .bias at 2+2i, i=0..7   */
  SPNERF_P_SIGMA_W = 17, SPNERF_P_SIGMA_B = 18,
  SPNERF_P_FEATS_W = 19, SPNERF_P_FEATS_B = 20,
  SPNERF_P_SEM0_W = 21, SPNERF_P_SEM0_B = 22, SPNERF_P_SEM2_W = 23, SPNERF_P_SEM2_B = 24,
  SPNERF_P_RGB0_W = 25, SPNERF_P_RGB0_B = 26, SPNERF_P_RGB2_W = 27, SPNERF_P_RGB2_B = 28,
  SPNERF_P_SUN0_W = 29,          /* sun_v_net.{0,2,4,6}: weight at 29+2j, bias at 30+2j */
  SPNERF_P_SKY0_W = 37, SPNERF_P_SKY0_B = 38, SPNERF_P_SKY2_W = 39, SPNERF_P_SKY2_B = 40,
  SPNERF_P_BETA0_W = 41, SPNERF_P_BETA0_B = 42, SPNERF_P_BETA2_W = 43, SPNERF_P_BETA2_B = 44,
  SPNERF_NUM_PARAMS = 45
};

typedef struct SpnerfNetSizes {
  int64_t fwd_blob_bytes;      /* packed fp16 weight stream for the forward kernel          */
  int64_t bwd_blob_bytes;      /* packed (transposed) stream for the backward-data kernel   */
  int64_t small_floats;        /* fp32 block: biases, tiny last layers, embedding, sky net  */
  int64_t steps_bytes;         /* per step table (forward and backward each)                */
  int32_t fwd_steps, bwd_steps;
  int32_t save_slabs_per_tile; /* x 16384 bytes x ceil(points/128) = activation save area   */
  int32_t grad_slabs_per_tile; /* same unit: gradient save area (backward-data -> weight GEMMs) */
  int32_t n_out;               /* columns of the network output row                         */
  int32_t in_dim;
  int32_t tile_points;         /* 128 */
} SpnerfNetSizes;

/* host only; no device work */
int spnerf_net_sizes(const SpnerfNetConfig* cfg, SpnerfNetSizes* sizes_host);
/* Debug aid (host only): the MMA step list of the forward (backward = 0) or backward-data (1) kernel as
 * 8 int32 per step [n, tmem_col, a_slab, ksteps, first, last, issuer lane, 0]; returns the step count. */
int spnerf_debug_step_table(const SpnerfNetConfig* cfg, int backward, int32_t* out, int max_steps);

/* Packing the fp32 parameters into tensor-core operands, in two steps:
 *  spnerf_net_prepare  once per (configuration, parameter pointers, buffers): uploads the pack
 *                      tables into `pack_ws` and the step tables, zeroes operand padding;
 *                      params_host[k] is a device pointer (NULL for absent heads); synchronises.
 *  spnerf_net_pack     after every parameter update: three small kernels, fully asynchronous. */
int64_t spnerf_net_pack_workspace_bytes(const SpnerfNetConfig* cfg);
int spnerf_net_prepare(const SpnerfNetConfig* cfg, const float* const* params_host, void* pack_ws,
                       int64_t pack_ws_bytes, void* fwd_blob, void* bwd_blob, float* small, void* fwd_steps,
                       void* bwd_steps, void* stream);
int spnerf_net_pack(const SpnerfNetConfig* cfg, const void* pack_ws, void* fwd_blob, void* bwd_blob,
                    float* small, void* stream);

/* sky_color(sun_dir) per ray (models/spnerf.py:355, 244-249; constant along a ray).
 * sky (n_rays,3); hidden (n_rays,256) post-ReLU or NULL. */
int spnerf_sky_fwd(const float* small, const SpnerfNetConfig* cfg, const float* rays, int64_t n_rays,
                   float* sky, float* hidden, void* stream);

/* Backward of the sky network over rays: += into the four sky_color gradients (pre-zeroed).
 * sky, hidden: outputs of spnerf_sky_fwd; g_sky: (n_rays,3) dL/d sky summed over samples. */
int spnerf_sky_bwd(const float* small, const SpnerfNetConfig* cfg, const float* rays, const float* sky,
                   const float* hidden, const float* g_sky, int64_t n_rays, float* g_w0, float* g_b0,
                   float* g_w2, float* g_b2, void* stream);

typedef struct SpnerfMlpFwd {
  SpnerfNetConfig cfg;
  const float* rays;      /* (n_rays, 11) */
  const float* z;         /* (n_rays, n_samples) sample depths; points = origin + dir * z      */
  const float* xyz;       /* optional (n_rays*n_samples, 3): used instead of rays/z if non-NULL */
  const float* dir_override; /* optional (n_rays,3): march along this instead of rays[:,3:6]
                                (solar-correction pass, modules/rendering.py:172)             */
  const int64_t* labels;  /* (n_rays) or NULL; -100 = ignore (models/spnerf.py:310-315)        */
  const float* t_emb;     /* (n_rays, t_dim) or NULL                                           */
  const float* sky;       /* (n_rays, 3) from spnerf_sky_fwd                                   */
  int64_t n_rays;
  int32_t n_samples;
  int32_t n_steps;
  const void* blob;       /* fwd_blob */
  const void* steps;      /* fwd_steps (kept for layout; the step list now travels as a launch parameter) */
  const float* small;
  float* out;             /* (n_rays*n_samples, n_out) fp32, reference column order            */
  void* saves;            /* activation save area or NULL (inference)                          */
  int32_t debug_flags;    /* must be 0.  Timing experiments only: 1 = skip the weight copies, 2 = skip
                             the epilogue arithmetic, 4 = skip the MMAs (results invalid); 128 = copy half
                             of every saved tile out of shared memory during the next MMA phase instead
                             of storing all of it from registers (same results, slower)             */
  int32_t _pad;
} SpnerfMlpFwd;
int spnerf_mlp_fwd(const SpnerfMlpFwd* args, void* stream);

/* Backward of the point network, in two kernels (replaces autograd through SPNeRF.forward):
 *  1. spnerf_mlp_bwd_data: dL/d(out) -> pre-activation gradients of every layer (fp16 tiles in
 *     `grad_saves`, scaled by the power of two written to `scale_out`), plus the gradients that
 *     are cheap row reductions: label embedding, last-layer biases, transient embedding input;
 *  2. spnerf_mlp_bwd_weights: every weight / bias gradient as tensor-core GEMMs over the point
 *     dimension between `grad_saves` and the forward's `saves`.                                 */
typedef struct SpnerfMlpBwd {
  SpnerfNetConfig cfg;
  const float* g_out;       /* (points, n_out) from spnerf_composite_bwd                         */
  const float* out;         /* (points, n_out) forward output                                    */
  const float* rays;
  const int64_t* labels;    /* (n_rays) or NULL                                                  */
  const float* t_emb;       /* (n_rays, t_dim) or NULL                                           */
  int64_t n_rays;
  int32_t n_samples;
  int32_t n_steps;          /* sizes.bwd_steps */
  const void* blob;         /* bwd_blob  */
  const void* steps;        /* bwd_steps */
  const float* small;
  const void* saves;        /* written by spnerf_mlp_fwd */
  void* grad_saves;         /* grad_slabs_per_tile * 16384 * ceil(points/128) bytes              */
  const float* g_absmax;    /* max|g_out| (device scalar, from spnerf_composite_bwd)             */
  float* scale_out;         /* device scalar: the gradient scale used                            */
  float* g_emb;             /* (C+1, emb_dim) += ; pre-zeroed; padding row stays zero            */
  float* g_small_bias;      /* 14 floats += : d rgb.2.bias(3), sun.6.bias, sigma.bias, beta.2.bias,
                               logit.2.bias(8); pre-zeroed                                       */
  float* g_t_emb;           /* (n_rays, t_dim) += or NULL; pre-zeroed                            */
  int32_t debug_flags;      /* must be 0.  Timing experiments only: 1 / 2 / 4 as in SpnerfMlpFwd (results
                               invalid); 256 / 512 = two / one (instead of three) of the four column groups
                               store their gradient tile part from registers (same results)          */
  int32_t _pad;
} SpnerfMlpBwd;
int spnerf_mlp_bwd_data(const SpnerfMlpBwd* args, void* stream);

typedef struct SpnerfMlpWgrad {
  SpnerfNetConfig cfg;
  int64_t n_points;
  const void* saves;
  const void* grad_saves;
  const float* scale;        /* scale_out of spnerf_mlp_bwd_data                                 */
  float* const* grads_host;  /* SPNERF_NUM_PARAMS device pointers (same slots / shapes as the
                                parameters); weight and hidden-layer bias gradients are OVERWRITTEN;
                                slots fed by g_small_bias / g_emb are not touched                 */
  void* workspace;           /* >= spnerf_mlp_wgrad_workspace_bytes(cfg) bytes                   */
  int64_t workspace_bytes;
  /* Optional self-cleaning accumulator block (SPNERF_ACCUM_FLOATS floats, zero before the first use): the slots that
   * are summed with atomics (spnerf_mlp_bwd_data's g_small_bias / g_emb, spnerf_sky_bwd's four outputs) can point
   * into it at the SPNERF_ACC_* offsets; the reduce kernel of spnerf_mlp_bwd_weights then copies them to their
   * grads_host slots (rgb_from_xyzdir.2.bias, sun_v_net.6.bias, sigma_from_xyz.0.bias, beta_from_xyz.2.bias,
   * logit_from_label.2.bias, semantic_embedding.weight, sky_color.*) and clears the block, so a step needs no
   * memset of the gradient buffer.  `absmax_reset` (optional) is set to 0 by the same kernel (spnerf_composite_bwd
   * accumulates it with atomicMax).  Both are baked into the tables by spnerf_mlp_wgrad_prepare. */
  float* accum;
  float* absmax_reset;
} SpnerfMlpWgrad;
enum { SPNERF_ACC_SMALL_BIAS = 0, SPNERF_ACC_EMB = 16, SPNERF_ACC_SKY_W0 = 96, SPNERF_ACC_SKY_B0 = 864,
       SPNERF_ACC_SKY_W2 = 1120, SPNERF_ACC_SKY_B2 = 1888, SPNERF_ACCUM_FLOATS = 1892 };
int64_t spnerf_mlp_wgrad_workspace_bytes(const SpnerfNetConfig* cfg);
/* uploads the GEMM tables into the workspace tail: once per (cfg, grads_host, workspace); syncs the stream */
int spnerf_mlp_wgrad_prepare(const SpnerfMlpWgrad* args, void* stream);
int spnerf_mlp_bwd_weights(const SpnerfMlpWgrad* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Volume integration (replaces models/spnerf.py:109-157 and its autograd graph).
 * `out` is the network output (n_rays*n_samples, n_out) with the reference's columns
 * [albedo(3), sigma, sun, sky(3), (beta), (logits)] (models/spnerf.py:110-113,148-156).
 * ------------------------------------------------------------------------------------------- */
typedef struct SpnerfCompositeFwd {
  const float* out;
  const float* z;            /* (n_rays, n_samples) sorted depths                                */
  const float* noise;        /* (n_rays, n_samples) standard normals or NULL (spnerf.py:121-122)  */
  int64_t n_rays;
  int32_t n_samples;         /* <= 256 */
  int32_t n_out;
  int32_t col_sem, n_sem;    /* first logit column / number of classes (0: no semantic head)      */
  float noise_std;
  int32_t _pad;
  float* weights;            /* (n_rays, n_samples)  alpha_i * T_i, or NULL (image export: not needed) */
  float* transparency;       /* (n_rays, n_samples)  T_i, or NULL                                 */
  float* rgb;                /* (n_rays, 3) clamped to [0,1]                                      */
  float* rgb_raw;            /* (n_rays, 3) before the clamp, or NULL (needed by the backward)    */
  float* depth;              /* (n_rays)                                                          */
  float* sem_logits;         /* (n_rays, n_sem): plain mean over samples (spnerf.py:156)          */
  /* Per-ray composited auxiliaries for image export (replaces the (B,N,k) tensors main.py:75-76 ships to
   * the host and the sums eval.py:75-101 forms there), both optional:
   *   ray_aux (n_rays, 8) = [sum_i w_i albedo_i (3), sum_i w_i sun_i, sum_i w_i sky_i (3), sum_i w_i beta_i]
   *   sem_argmax (n_rays) = argmax over classes of sem_logits, first maximum (eval.py:63, main.py:228) */
  float* ray_aux;
  int32_t* sem_argmax;
  int32_t col_beta;          /* beta column of `out`, or -1                                        */
  int32_t _pad2;
} SpnerfCompositeFwd;
int spnerf_composite_fwd(const SpnerfCompositeFwd* args, void* stream);

typedef struct SpnerfCompositeBwd {
  const float* out; const float* z; const float* noise;
  const float* weights; const float* transparency; const float* rgb_raw;   /* saved by the forward */
  const float* g_rgb;          /* (n_rays,3) or NULL          upstream gradients ...              */
  const float* g_depth;        /* (n_rays) or NULL                                                */
  const float* g_sem_logits;   /* (n_rays,n_sem) or NULL                                          */
  const float* g_weights;      /* (n_rays,n_samples) or NULL                                      */
  const float* g_transparency; /* (n_rays,n_samples) or NULL                                      */
  const float* g_out_ext;      /* (points,n_out) or NULL: gradients arriving directly on `out`    */
  int64_t n_rays;
  int32_t n_samples, n_out, col_sem, n_sem;
  float noise_std;
  int32_t _pad;
  float* g_out;                /* (points, n_out) gradient w.r.t. the network output              */
  float* g_sky_ray;            /* (n_rays,3) or NULL: sum over samples of the sky-colour gradient */
  float* g_absmax;             /* one float, atomically max-updated with max|g_out| (pre-zeroed)  */
} SpnerfCompositeBwd;
int spnerf_composite_bwd(const SpnerfCompositeBwd* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Loss reductions with their gradients in one pass (replaces modules/metrics.py:27-45 colour MSE,
 * :68-159 DepthLoss subset / all-depth MSE variants, :162-183 SemanticLoss).
 * Scalars land in `losses` (device): [0] colour, [1] depth, [2] semantic, [3] rays the depth term
 * applied to, [4] labelled rays.  Gradients are d(loss_k)/d(input), NOT summed over k.
 * ------------------------------------------------------------------------------------------- */
typedef struct SpnerfLosses {
  int64_t n_rays;
  int32_t n_samples, n_sem;
  /* colour (NULL rgb: skip) */
  const float* rgb; const float* rgb_target; float* g_rgb;
  /* depth supervision (NULL depth: skip) */
  const float* depth; const float* z; const float* weights;
  const float* target_depth; const float* target_weight; const float* target_std;
  const int64_t* valid_depth;  /* NULL: every ray is valid (metrics.py:83-87) */
  float lambda_ds;             /* the constructor's lambda_ds; the kernel applies the /3 of metrics.py:71 */
  int32_t use_all_depth;       /* metrics.py:140,154-156 */
  float* g_depth;
  /* semantic (NULL sem_logits: skip) */
  const float* sem_logits; const int64_t* labels; float lambda_ss; int32_t _pad;
  float* g_sem_logits;
  float* losses;               /* 8 floats */
  void* workspace;             /* >= spnerf_losses_workspace_bytes() bytes, ZEROED once by the caller; the kernels
                                  leave it zeroed (no memsets on the stream) */
  /* --GNLL subset variant (metrics.py:76,129-130: GaussianNLLLoss with the predicted STD passed as the variance,
   * eps 1e-6): needs use_all_depth = 0; the gradient reaches the weights through the predicted STD */
  int32_t gnll, _pad2;
  float* g_weights;            /* (n_rays, n_samples), written when gnll != 0 */
  int64_t target_stride;       /* element stride of target_depth / target_weight (0 = 1); 2 reads the two columns of the
                                  reference's (n_rays, 2) `depths` tensor in place */
} SpnerfLosses;
int64_t spnerf_losses_workspace_bytes(void);
int spnerf_losses(const SpnerfLosses* args, void* stream);

/* Solar-correction terms of Shadow-NeRF (replaces modules/metrics.py:17-24 solar_correction):
 *   sc_term2 = lambda_sc/3 * mean_r sum_i (transparency_sc - sun_sc)^2
 *   sc_term3 = lambda_sc/3 * mean_r (1 - sum_i weights_sc * sun_sc)
 * transparency_sc / weights_sc are constants (the reference detaches them); the gradient goes to sun_sc only.
 * backward = 0: writes losses[0..1].  backward = 1: writes g_sun = upstream[0] d term2 + upstream[1] d term3
 * (upstream: 2 floats on the device, NULL = ones). */
typedef struct SpnerfLossSolar {
  int64_t n_rays;
  int32_t n_samples, _pad;
  const float* transparency_sc; const float* weights_sc;   /* (n_rays, n_samples) */
  const float* sun_sc; int64_t sun_stride;                 /* element (r,i) at sun_sc[(r*n_samples+i)*sun_stride] */
  float lambda_sc; int32_t _pad2;
  const float* upstream;
  float* g_sun;                                            /* (n_rays, n_samples) */
  float* losses;                                           /* 2 floats */
  void* workspace;                                         /* >= spnerf_losses_workspace_bytes() */
} SpnerfLossSolar;
int spnerf_loss_solar(const SpnerfLossSolar* args, int backward, void* stream);

/* Uncertainty-aware colour loss of Sat-NeRF (replaces modules/metrics.py:10-14 uncertainty_aware_loss):
 *   beta_ray = sum_i weights * beta + beta_min;  color = mean((rgb - target)^2 / (2 beta_ray^2));
 *   logbeta = (3 + mean(log beta_ray)) / 2.
 * backward = 0: writes losses[0..1] and beta_ray.  backward = 1: reads beta_ray, writes g_rgb, g_weights, g_beta
 * scaled by upstream[0] (color) and upstream[1] (logbeta). */
typedef struct SpnerfLossUncertainty {
  int64_t n_rays;
  int32_t n_samples, _pad;
  const float* rgb; const float* rgb_target;               /* (n_rays, 3) */
  const float* weights;                                    /* (n_rays, n_samples) */
  const float* beta; int64_t beta_stride;                  /* element (r,i) at beta[(r*n_samples+i)*beta_stride] */
  float beta_min; int32_t _pad2;
  const float* upstream;
  float* beta_ray;                                         /* (n_rays) */
  float* g_rgb; float* g_weights; float* g_beta;           /* (n_rays,3), (n_rays,n_samples) x 2 */
  float* losses;                                           /* 2 floats */
  void* workspace;
} SpnerfLossUncertainty;
int spnerf_loss_uncertainty(const SpnerfLossUncertainty* args, int backward, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Ray samplers (replace modules/rendering.py:128-144 and :14-116,165-167).  Uniform draws are
 * inputs; `t_table` = torch.linspace(0,1,n_samples) and `gauss_table` =
 * 1/sqrt(2 pi) * exp(-0.5 * linspace(-3,3,n_samples-1)^2) as computed by torch on the host
 * (rendering.py:60,68-70), so the device arithmetic reproduces the reference bit for bit.
 * ------------------------------------------------------------------------------------------- */
int spnerf_sample_coarse(const float* rays, const float* t_table, const float* uniforms, int64_t n_rays,
                         int32_t n_samples, float* z, void* stream);
/* Same sampler with the uniforms drawn on the device (replaces the torch.rand of rendering.py:143): Philox4x32-10
 * keyed by rng_state[0] (seed), counter (element, rng_state[1] = step).  rng_state: 3 x uint64 {seed, step, 0};
 * the launch advances the step itself, so a captured CUDA graph draws fresh numbers on every replay. */
int spnerf_sample_coarse_rng(const float* rays, const float* t_table, uint64_t* rng_state, int64_t n_rays,
                             int32_t n_samples, float* z, void* stream);

typedef struct SpnerfGuided {
  const float* rays;          /* (n_rays,11); rays[0,6:8] clamp every ray (rendering.py:95) */
  const float* z;             /* (n_rays,n) coarse depths (sorted)                                  */
  const float* weights;       /* (n_rays,n) and ...                                                 */
  const float* depth;         /* (n_rays)   ... of the first pass (rendering.py:77-81)               */
  const int64_t* valid_depth; /* (n_rays) or NULL (test mode): >0 -> sample around the depth prior   */
  const float* target_depth;  /* element r at target_depth[r*target_depth_stride] (depths[:,0])      */
  int64_t target_depth_stride;
  const float* target_std;    /* (n_rays) */
  const float* u_pred;        /* (n_rays,n) uniforms for rays sampled around the predicted depth     */
  const float* u_gt;          /* (n_rays,n) uniforms for rays sampled around the prior (row = ray)   */
  const float* t_table;       /* (n) */
  const float* gauss_table;   /* (n-1) */
  int64_t n_rays;
  int32_t n_samples;
  int32_t _pad;
  float* z_unsort;            /* (n_rays,2n) = [z, sort(z2)]   (rendering.py:165-166)                */
  float* z_sorted;            /* (n_rays,2n) = sort([z, z2])   (rendering.py:167)                    */
  int32_t* searchsorted_out;  /* (n_rays,n) indices of rendering.py:38, or NULL (parity tests)       */
} SpnerfGuided;
int spnerf_sample_guided(const SpnerfGuided* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Geometry either side of the renderer.  The RPC localisation (rpcm) before the first and
 * the UTM projection / rasterisation (pyproj, plyflatten) after the second stay on the host.
 * ------------------------------------------------------------------------------------------- */
/* Rays from localised pixels (replaces datasets/satellite_scene.py:38-68 get_rays after rpc.localization, the
 * float32 cast, normalize_rays :415-425 and the sun-direction columns :463-473; geodetic_to_ecef = modules/utils.py:80-100
 * in fp64).  Row i of `rays`: [origin(3), direction(3), near, far, (sun_dir(3))] float32, row_stride floats apart. */
typedef struct SpnerfRaysFromGeodetic {
  const double* lon_near; const double* lat_near;   /* (n) degrees: pixels localised at alt_near = the MAXIMUM altitude */
  const double* lon_far; const double* lat_far;     /* (n) degrees: the same pixels at alt_far = the minimum altitude   */
  double alt_near, alt_far;
  float center[3]; float range;                     /* scene.loc offsets / max scale (satellite_scene.py:122-124)        */
  int32_t normalize;                                /* 1: apply normalize_rays                                           */
  int32_t has_sun;                                  /* 1: write sun_dir into columns 8..10                               */
  float sun_dir[3]; int32_t row_stride;             /* >= 8 (11 with has_sun)                                            */
  int64_t n_rays;
  float* rays;
} SpnerfRaysFromGeodetic;
int spnerf_rays_from_geodetic(const SpnerfRaysFromGeodetic* args, void* stream);

/* DSM point cloud (replaces datasets/satellite_scene.py:475-505 get_latlonalt_from_nerf_prediction and
 * modules/utils.py:103-120 ecef_to_latlon_custom): lat / lon in degrees and altitude in metres, fp64. */
typedef struct SpnerfPointsToGeodetic {
  const float* rays; int32_t row_stride; int32_t _pad;   /* normalised rays, >= 6 columns                    */
  const float* depth;                                    /* (n) predicted depth (render_rays' depth_coarse)  */
  float center[3]; float range;
  int64_t n_rays;
  double* lat; double* lon; double* alt;                 /* (n) each                                         */
} SpnerfPointsToGeodetic;
int spnerf_points_to_geodetic(const SpnerfPointsToGeodetic* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Optimiser step over the flat fp32 parameter buffer (replaces torch.optim.Adam(parameters, lr=args.lr,
 * weight_decay=0) of main.py:96-97; torch's update order; bias corrections from `step` >= 1 and the
 * 1 - beta factors are formed in double like torch's).  All four
 * buffers hold n floats and are 16-byte aligned.
 * ------------------------------------------------------------------------------------------- */
int spnerf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                     int64_t step, double lr, double beta1, double beta2, double eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPNERF_B200_H */
