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

#ifdef __cplusplus
}
#endif
#endif /* SPNERF_B200_H */
