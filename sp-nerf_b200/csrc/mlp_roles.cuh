// Warp roles shared by the forward and backward-data point-network kernels: the weight producer
// and the MMA issuer, both driven by the step list of mlp_pack.cu, and the epilogue-side handshake.
#pragma once
#include "sm100.cuh"
#include "net_plan.h"

namespace roles {
using namespace sm100;
using namespace net;

constexpr int kThreads = 384;
constexpr int kEpiThreads = 256;
constexpr int kEpiWarp0 = 4;

struct Smem {
  uint8_t* act;
  uint8_t* wst;
  uint64_t* bar_full;    // [2] weight stage landed
  uint64_t* bar_empty;   // [2] weight stage consumed
  uint64_t* bar_mma;     // MMA phase retired -> epilogue
  uint64_t* bar_epi;     // epilogue phase done -> MMA
  uint32_t* tmem_slot;
};

__device__ __forceinline__ Smem carve(uint8_t* smem) {
  Smem s;
  s.act = smem;
  s.wst = smem + kNumSlabs * kSlabBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBars);
  s.bar_full = bars; s.bar_empty = bars + 2; s.bar_mma = bars + 4; s.bar_epi = bars + 5;
  s.tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  return s;
}

// all threads; returns the TMEM base address
__device__ __forceinline__ uint32_t setup(const Smem& s, uint8_t* smem) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) { atomicCAS(&g_watchdog_code, 0u, 900u); __trap(); }
    mbar_init(&s.bar_full[0], 1); mbar_init(&s.bar_full[1], 1);
    mbar_init(&s.bar_empty[0], 1); mbar_init(&s.bar_empty[1], 1);
    mbar_init(s.bar_mma, 1); mbar_init(s.bar_epi, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(s.tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *s.tmem_slot;
}

__device__ __forceinline__ void teardown(uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 1) tmem_dealloc(tmem_base, 512);
}

// one lane of warp 0
__device__ __forceinline__ void producer_loop(const Smem& s, const uint8_t* blob, const MmaStep* steps, int n_steps,
                                              int64_t n_tiles, int debug) {
  uint32_t stage = 0, phase = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    for (int i = 0; i < n_steps; ++i) {
      const MmaStep st = steps[i];
      mbar_wait(&s.bar_empty[stage], phase ^ 1, 10);
      const uint32_t bytes = (uint32_t)st.n * 128u;
      if (debug & 1) { mbar_arrive(&s.bar_full[stage]); }
      else {
        mbar_expect_tx(&s.bar_full[stage], bytes);
        bulk_g2s(s.wst + stage * kWStageBytes, blob + (size_t)st.w_off16 * 16, bytes, &s.bar_full[stage]);
      }
      stage ^= 1; if (stage == 0) phase ^= 1;
    }
  }
}

// one lane of warp 1
__device__ __forceinline__ void mma_loop(const Smem& s, uint32_t tmem_base, const MmaStep* steps, int n_steps,
                                         int64_t n_tiles, int debug) {
  constexpr uint64_t tmpl = make_smem_desc_template(16, 1024, kSwizzle128B);
  const uint32_t act_addr = smem_u32(s.act), wst_addr = smem_u32(s.wst);
  uint32_t stage = 0, phase = 0, epi_par = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    int i = 0;
    while (i < n_steps) {
      mbar_wait(s.bar_epi, epi_par, 20); epi_par ^= 1;
      tc_fence_after();
      bool last;
      do {
        const MmaStep st = steps[i++];
        last = st.last;
        mbar_wait(&s.bar_full[stage], phase, 21);
        tc_fence_after();
        const uint32_t a0 = act_addr + (uint32_t)st.a_slab * kSlabBytes, b0 = wst_addr + stage * kWStageBytes;
        const uint32_t idesc = make_idesc_f16(128, st.n, 0, 0);
        for (uint32_t k = 0; k < ((debug & 4) ? 0u : st.ksteps); ++k)
          umma_f16(tmem_base + st.tmem_col, smem_desc(tmpl, a0 + k * 32), smem_desc(tmpl, b0 + k * 32), idesc,
                   (st.first && k == 0) ? 0u : 1u);
        umma_commit(&s.bar_empty[stage]);
        stage ^= 1; if (stage == 0) phase ^= 1;
      } while (!last);
      umma_commit(s.bar_mma);
    }
  }
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// Epilogue-side view of the phase handshake (all 256 epilogue threads call every method).
struct EpiSync {
  const Smem& s;
  bool issuer;
  uint32_t mma_par = 0;
  bool stores_pending = false;
  __device__ EpiSync(const Smem& s_) : s(s_), issuer(threadIdx.x == kEpiWarp0 * 32) {}

  __device__ __forceinline__ void drain_stores() {   // earlier bulk stores must have read their slabs
    if (stores_pending) {
      if (issuer) bulk_wait_read<0>();
      epi_bar_sync();
      stores_pending = false;
    }
  }
  __device__ __forceinline__ void begin() {          // wait for the MMA phase that feeds this epilogue
    mbar_wait(s.bar_mma, mma_par, 30); mma_par ^= 1;
    tc_fence_after();
    drain_stores();
  }
  // publish shared-memory writes to the async proxy, order TMEM reads, release the MMA warp and
  // optionally stream `nslabs` activation slabs to global memory
  __device__ __forceinline__ void end(bool signal, uint8_t* save_dst, int slab0, int nslabs) {
    fence_proxy_async_smem();
    tc_fence_before();
    epi_bar_sync();
    if (issuer) {
      if (signal) mbar_arrive(s.bar_epi);
      if (save_dst) {
        bulk_s2g(save_dst, s.act + slab0 * kSlabBytes, (uint32_t)nslabs * kSlabBytes);
        bulk_commit();
      }
    }
    if (save_dst) stores_pending = true;
  }
  __device__ __forceinline__ void finish() { if (issuer) bulk_wait_all<0>(); }
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace roles
