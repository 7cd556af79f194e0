// Library-level entry points that are not tied to one kernel file.
#include <cuda_runtime.h>
#include "../../include/spnerf_b200.h"

extern "C" unsigned int spnerf_watchdog_code_selftest(void);
extern "C" unsigned int spnerf_watchdog_code_fwd(void);
extern "C" unsigned int spnerf_watchdog_code_bwd(void);
extern "C" unsigned int spnerf_watchdog_code_wgrad(void);

extern "C" int spnerf_abi_version(void) { return SPNERF_ABI_VERSION; }

extern "C" unsigned int spnerf_watchdog_code(void) {
  unsigned int v = spnerf_watchdog_code_selftest();
  if (!v) v = spnerf_watchdog_code_fwd();
  if (!v) v = spnerf_watchdog_code_bwd();
  if (!v) v = spnerf_watchdog_code_wgrad();
  return v;
}

// sizeof of every argument struct, in the order of _cabi.STRUCTS (binding self-check)
extern "C" void spnerf_struct_sizes(int32_t* out) {
  int i = 0;
  out[i++] = (int32_t)sizeof(SpnerfUmmaSelftest);
  out[i++] = (int32_t)sizeof(SpnerfNetConfig);
  out[i++] = (int32_t)sizeof(SpnerfNetSizes);
  out[i++] = (int32_t)sizeof(SpnerfMlpFwd);
  out[i++] = (int32_t)sizeof(SpnerfCompositeFwd);
  out[i++] = (int32_t)sizeof(SpnerfCompositeBwd);
  out[i++] = (int32_t)sizeof(SpnerfLosses);
  out[i++] = (int32_t)sizeof(SpnerfGuided);
  out[i++] = (int32_t)sizeof(SpnerfMlpBwd);
  out[i++] = (int32_t)sizeof(SpnerfMlpWgrad);
  out[i++] = (int32_t)sizeof(SpnerfLossSolar);
  out[i++] = (int32_t)sizeof(SpnerfLossUncertainty);
  out[i++] = (int32_t)sizeof(SpnerfRaysFromGeodetic);
  out[i++] = (int32_t)sizeof(SpnerfPointsToGeodetic);
}
