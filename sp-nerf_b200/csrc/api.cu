// Library-level entry points that are not tied to one kernel file.
#include <cuda_runtime.h>
#include "../../include/spnerf_b200.h"

extern "C" unsigned int spnerf_watchdog_code_selftest(void);
extern "C" unsigned int spnerf_watchdog_code_fwd(void);

extern "C" int spnerf_abi_version(void) { return SPNERF_ABI_VERSION; }

extern "C" unsigned int spnerf_watchdog_code(void) {
  unsigned int v = spnerf_watchdog_code_selftest();
  if (!v) v = spnerf_watchdog_code_fwd();
  return v;
}
