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

/* ---------------------------------------------------------------------------------------------
 * Point network (replaces models/spnerf.py:162-369 SPNeRF.__init__/forward as executed through
 * models/spnerf.py:83-113, i.e. the chunked per-point MLP call of inference()).
 * ------------------------------------------------------------------------------------------- */
typedef struct SpnerfNetConfig {
  int32_t feat;            /* fc_units; only 512 is built (modules/opt.py:43)                 */
  int32_t layers;          /* fc_layers; only 8 (modules/opt.py:45)                           */
  int32_t skip_layer;      /* 4 (models/spnerf.py:164 skips=[4])                              */
  int32_t mapping;         /* 1: 10-frequency positional encoding (models/spnerf.py:5-37)     */
  int32_t sem;             /* 1: label embedding input + semantic head                        */
  int32_t num_sem_classes; /* C <= 8                                                          */
  int32_t emb_dim;         /* C * s_embedding_factor; encoded input width must stay <= 64     */
  int32_t beta;            /* 1: uncertainty head                                             */
  int32_t t_dim;           /* t_embbeding_tau <= 8                                            */
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
  int32_t n_out;               /* columns of the network output row                         */
  int32_t in_dim;
  int32_t tile_points;         /* 128 */
} SpnerfNetSizes;

/* host only; no device work */
int spnerf_net_sizes(const SpnerfNetConfig* cfg, SpnerfNetSizes* sizes_host);

/* fp32 parameters -> packed operands.  params_host[k] is a device pointer (or NULL for absent
 * heads).  Must be re-run after every parameter update. */
int spnerf_net_pack(const SpnerfNetConfig* cfg, const float* const* params_host, void* fwd_blob,
                    void* bwd_blob, float* small, void* fwd_steps, void* bwd_steps, void* stream);

/* sky_color(sun_dir) per ray (models/spnerf.py:355, 244-249; constant along a ray).
 * sky (n_rays,3); hidden (n_rays,256) post-ReLU or NULL. */
int spnerf_sky_fwd(const float* small, const SpnerfNetConfig* cfg, const float* rays, int64_t n_rays,
                   float* sky, float* hidden, void* stream);

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
  const void* steps;      /* fwd_steps */
  const float* small;
  float* out;             /* (n_rays*n_samples, n_out) fp32, reference column order            */
  void* saves;            /* activation save area or NULL (inference)                          */
  int32_t debug_flags;    /* must be 0.  Timing experiments only (results invalid): 1 = skip the
                             weight copies, 2 = skip the epilogue arithmetic, 4 = skip the MMAs  */
  int32_t _pad;
} SpnerfMlpFwd;
int spnerf_mlp_fwd(const SpnerfMlpFwd* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPNERF_B200_H */
