// Bring-up check for the tensor-core plumbing: one CTA copies two operand images into shared
// memory with the bulk-copy engine, issues a chain of tcgen05.mma instructions whose descriptors
// are supplied by the caller, and dumps the 128 x N fp32 accumulator from tensor memory.
// tests/test_umma_selftest.py builds the images with the slab layout of sm100.cuh and compares
// the dump with A * B^T, so a wrong descriptor field shows up as a numeric mismatch, not a hang.
#include "sm100.cuh"
#include "../../include/spnerf_b200.h"

using namespace sm100;

namespace {

struct SelftestDev {
  const uint8_t* a_img;
  const uint8_t* b_img;
  float* d_out;
  uint32_t a_bytes, b_bytes, n, ksteps, idesc;
  uint64_t a_tmpl, b_tmpl;
  uint32_t a_off[SPNERF_SELFTEST_MAX_KSTEPS];
  uint32_t b_off[SPNERF_SELFTEST_MAX_KSTEPS];
};

__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const __grid_constant__ SelftestDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((p.a_bytes + 1023u) & ~1023u);
  __shared__ __align__(8) uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_slot;

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_load, p.a_bytes + p.b_bytes);
    bulk_g2s(sA, p.a_img, p.a_bytes, &bar_load);
    bulk_g2s(sB, p.b_img, p.b_bytes, &bar_load);
    mbar_wait(&bar_load, 0, 101);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    for (uint32_t k = 0; k < p.ksteps; ++k) {
      umma_f16(tmem_base, smem_desc(p.a_tmpl, a0 + p.a_off[k]), smem_desc(p.b_tmpl, b0 + p.b_off[k]), p.idesc,
               k > 0 ? 1u : 0u);
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0, 102);
  tc_fence_after();

  // lane quarter `warp` of the accumulator -> rows 32*warp .. 32*warp+31
  const uint32_t row = warp * 32 + lane;
  for (uint32_t c0 = 0; c0 < p.n; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem_base + ((warp * 32u) << 16) + c0, v);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) p.d_out[row * p.n + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// Pair version (cluster of 2, cta_group::2): M = 256.  CTA r holds rows 128r.. of A and B rows
// (n/2)r.. ; the peer forwards "my operands landed" to the leader with a remote mbarrier arrive;
// the leader issues the MMAs and commits to both CTAs; each CTA dumps its own accumulator rows.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma2_selftest_kernel(const __grid_constant__ SelftestDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;     // dynamic shared memory starts at the same offset in both CTAs
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((p.a_bytes + 1023u) & ~1023u);
  __shared__ __align__(8) uint64_t bar_load, bar_peer, bar_mma;
  __shared__ uint32_t tmem_base_slot;
  const uint32_t rank = cluster_ctarank();
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) { atomicCAS(&g_watchdog_code, 0u, 903u); __trap(); }
    mbar_init(&bar_load, 1);
    mbar_init(&bar_peer, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc2(&tmem_base_slot, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_load, p.a_bytes + p.b_bytes);
    bulk_g2s(sA, p.a_img + (size_t)rank * p.a_bytes, p.a_bytes, &bar_load);
    bulk_g2s(sB, p.b_img + (size_t)rank * p.b_bytes, p.b_bytes, &bar_load);
    mbar_wait(&bar_load, 0, 111);
    if (rank == 1) {
      mbar_arrive_remote(mapa_shared(smem_u32(&bar_peer), 0));
    } else {
      mbar_wait_cluster(&bar_peer, 0, 112);
      tc_fence_after();
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
      for (uint32_t k = 0; k < p.ksteps; ++k)
        umma2_f16(tmem_base, smem_desc(p.a_tmpl, a0 + p.a_off[k]), smem_desc(p.b_tmpl, b0 + p.b_off[k]), p.idesc,
                  k > 0 ? 1u : 0u);
      umma2_commit(&bar_mma, 3);
    }
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0, 113);
  tc_fence_after();
  const uint32_t row = warp * 32 + lane;
  for (uint32_t c0 = 0; c0 < p.n; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem_base + ((warp * 32u) << 16) + c0, v);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) p.d_out[(size_t)(rank * 128 + row) * p.n + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tmem_base, 512);
}

}  // namespace

// a_img: two A images (rows 0..127, rows 128..255), a_bytes each; b_img: two B halves, b_bytes each;
// idesc must say M = 256; d_out is [256][n].
extern "C" int spnerf_selftest_umma2(const SpnerfUmmaSelftest* a, void* stream) {
  if (!a || a->ksteps == 0 || a->ksteps > SPNERF_SELFTEST_MAX_KSTEPS || a->n < 16 || a->n > 256 || (a->n % 16))
    return SPNERF_ERR_BAD_ARG;
  if ((a->a_bytes % 16) || (a->b_bytes % 16)) return SPNERF_ERR_BAD_ARG;
  SelftestDev p;
  p.a_img = static_cast<const uint8_t*>(a->a_img);
  p.b_img = static_cast<const uint8_t*>(a->b_img);
  p.d_out = a->d_out;
  p.a_bytes = a->a_bytes; p.b_bytes = a->b_bytes; p.n = a->n; p.ksteps = a->ksteps; p.idesc = a->idesc;
  p.a_tmpl = a->a_desc_template; p.b_tmpl = a->b_desc_template;
  for (uint32_t i = 0; i < SPNERF_SELFTEST_MAX_KSTEPS; ++i) { p.a_off[i] = a->a_off[i]; p.b_off[i] = a->b_off[i]; }
  const size_t smem = ((a->a_bytes + 1023u) & ~1023u) + ((a->b_bytes + 1023u) & ~1023u);
  if (smem > 220 * 1024) return SPNERF_ERR_BAD_ARG;
  cudaError_t e = cudaFuncSetAttribute(umma2_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -(int)e;
  umma2_selftest_kernel<<<2, 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

extern "C" int spnerf_selftest_umma(const SpnerfUmmaSelftest* a, void* stream) {
  if (!a || a->ksteps == 0 || a->ksteps > SPNERF_SELFTEST_MAX_KSTEPS || a->n < 16 || a->n > 256 || (a->n % 16))
    return SPNERF_ERR_BAD_ARG;
  if ((a->a_bytes % 16) || (a->b_bytes % 16)) return SPNERF_ERR_BAD_ARG;
  SelftestDev p;
  p.a_img = static_cast<const uint8_t*>(a->a_img);
  p.b_img = static_cast<const uint8_t*>(a->b_img);
  p.d_out = a->d_out;
  p.a_bytes = a->a_bytes;
  p.b_bytes = a->b_bytes;
  p.n = a->n;
  p.ksteps = a->ksteps;
  p.idesc = a->idesc;
  p.a_tmpl = a->a_desc_template;
  p.b_tmpl = a->b_desc_template;
  for (uint32_t i = 0; i < SPNERF_SELFTEST_MAX_KSTEPS; ++i) {
    p.a_off[i] = a->a_off[i];
    p.b_off[i] = a->b_off[i];
  }
  const size_t smem = ((a->a_bytes + 1023u) & ~1023u) + ((a->b_bytes + 1023u) & ~1023u) + 1024;
  if (smem > 220 * 1024) return SPNERF_ERR_BAD_ARG;
  cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -(int)e;
  umma_selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

SPNERF_DEFINE_WATCHDOG_GETTER(spnerf_watchdog_code_selftest)
