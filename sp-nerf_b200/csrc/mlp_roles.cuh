// Warp roles shared by the forward and backward-data point-network kernels: the weight producer
// and the MMA issuer, both driven by the step list of mlp_pack.cu, and the epilogue-side handshake.
//
// 640 threads per CTA, one CTA per SM:
//   warp 0        weight producer (one lane): 1-D bulk copies of pre-packed fp16 B tiles, 2-stage ring
//   warp 1        MMA issuer (one lane): tcgen05.mma M=128, N<=256, K=16, fp16 x fp16 -> fp32 in TMEM
//   warps 2-3     idle (keep the epilogue warps aligned to TMEM lane quarters)
//   warps 4-19    epilogue: warp w reads TMEM lanes 32*(w%4).., i.e. point row 32*(w%4)+lane, and
//                 column group (w-4)/4 of the phase's accumulator
#pragma once
#include "sm100.cuh"
#include "net_plan.h"

namespace roles {
using namespace sm100;
using namespace net;

constexpr int kThreads = 640;
constexpr int kEpiThreads = 512;
constexpr int kEpiWarp0 = 4;
constexpr int kColGroups = 4;

struct Smem {
  uint8_t* act;
  uint8_t* aux;
  uint8_t* wst;
  uint64_t* bar_full;    // [2] weight stage landed
  uint64_t* bar_empty;   // [2] weight stage consumed
  uint64_t* bar_mma;     // MMA phase retired -> epilogue
  uint64_t* bar_epi;     // epilogue phase done -> MMA
  uint64_t* bar_par;     // parameter region landed (once)
  uint32_t* tmem_slot;
};

__device__ __forceinline__ Smem carve(uint8_t* smem) {
  Smem s;
  s.act = smem;
  s.aux = smem + kOffAux;
  s.wst = smem + kOffWst;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBars);
  s.bar_full = bars; s.bar_empty = bars + 2; s.bar_mma = bars + 4; s.bar_epi = bars + 5; s.bar_par = bars + 6;
  s.tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  return s;
}

// all threads; returns the TMEM base address.  `smallw` (global image of the parameter region) is
// copied into shared memory once; consumers wait on bar_par (parity 0).
__device__ __forceinline__ uint32_t setup(const Smem& s, uint8_t* smem, const float* smallw) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) { atomicCAS(&g_watchdog_code, 0u, 900u); __trap(); }
    mbar_init(&s.bar_full[0], 1); mbar_init(&s.bar_full[1], 1);
    mbar_init(&s.bar_empty[0], 1); mbar_init(&s.bar_empty[1], 1);
    mbar_init(s.bar_mma, 1); mbar_init(s.bar_epi, 1); mbar_init(s.bar_par, 1);
    fence_mbar_init();
    mbar_expect_tx(s.bar_par, kSmallWFloats * 4);
    bulk_g2s(smem + kOffRgb2, smallw, 3072 * 4, s.bar_par);              // rgb2 | sem2
    bulk_g2s(smem + kOffSun6, smallw + 3072, 512 * 4, s.bar_par);        // sun6 | beta2
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
    MmaStep nxt = steps[0];
    for (int i = 0; i < n_steps; ++i) {
      const MmaStep st = nxt;
      if (i + 1 < n_steps) nxt = steps[i + 1];        // in flight while this lane waits below
      mbar_wait(&s.bar_empty[stage], phase ^ 1, 10);
      const uint32_t bytes = (uint32_t)st.bytes16 * 16u;
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
  constexpr uint64_t tmpl_aux = make_smem_desc_template(128, 256, kSwizzleNone);   // 16-column no-swizzle operand
  const uint32_t act_addr = smem_u32(s.act), wst_addr = smem_u32(s.wst), aux_addr = smem_u32(s.aux);
  uint32_t stage = 0, phase = 0, epi_par = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    int i = 0;
    MmaStep nxt = steps[0];
    while (i < n_steps) {
      mbar_wait(s.bar_epi, epi_par, 20); epi_par ^= 1;
      tc_fence_after();
      bool last;
      do {
        const MmaStep st = nxt;
        ++i;
        if (i < n_steps) nxt = steps[i];
        last = st.last;
        mbar_wait(&s.bar_full[stage], phase, 21);
        tc_fence_after();
        const uint32_t b0 = wst_addr + stage * kWStageBytes;
        const uint32_t idesc = make_idesc_f16(128, st.n, 0, 0);
        if (!(debug & 4)) {
          if (st.a_slab == kAuxSlab) {
            umma_f16(tmem_base + st.tmem_col, smem_desc(tmpl_aux, aux_addr), smem_desc(tmpl_aux, b0), idesc,
                     st.first ? 0u : 1u);
          } else {
            const uint32_t a0 = act_addr + (uint32_t)st.a_slab * kSlabBytes;
            for (uint32_t k = 0; k < st.ksteps; ++k)
              umma_f16(tmem_base + st.tmem_col, smem_desc(tmpl, a0 + k * 32), smem_desc(tmpl, b0 + k * 32), idesc,
                       (st.first && k == 0) ? 0u : 1u);
          }
        }
        umma_commit(&s.bar_empty[stage]);
        stage ^= 1; if (stage == 0) phase ^= 1;
      } while (!last);
      umma_commit(s.bar_mma);
    }
  }
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// Epilogue-side view of the phase handshake (all epilogue threads call every method).
struct EpiSync {
  const Smem& s;
  bool issuer;
  uint32_t mma_par = 0;
  bool stores_pending = false;
  long long* prof;        // optional phase clock log (block 0 only)
  int prof_i = 0;
  __device__ EpiSync(const Smem& s_, long long* prof_)
      : s(s_), issuer(threadIdx.x == kEpiWarp0 * 32), prof((blockIdx.x == 0 && threadIdx.x == kEpiWarp0 * 32) ? prof_ : nullptr) {}

  __device__ __forceinline__ void stamp() {
    if (prof && prof_i < 256) prof[prof_i++] = clock64();
  }
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
    stamp();
    drain_stores();
  }
  // publish shared-memory writes to the async proxy, order TMEM reads, release the MMA warp
  __device__ __forceinline__ void end(bool signal) {
    fence_proxy_async_smem();
    tc_fence_before();
    epi_bar_sync();
    if (issuer && signal) mbar_arrive(s.bar_epi);
    stamp();
  }
  // after end(): stream `nslabs` activation slabs (starting at slab0) to global memory
  __device__ __forceinline__ void store_slabs(uint8_t* dst, int slab0, int nslabs) {
    if (!dst) return;
    if (issuer) {
      bulk_s2g(dst, s.act + slab0 * kSlabBytes, (uint32_t)nslabs * kSlabBytes);
      bulk_commit();
    }
    stores_pending = true;
  }
  __device__ __forceinline__ void finish() { if (issuer) bulk_wait_all<0>(); }
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stg16(uint8_t* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

// byte offset of 8 consecutive columns starting at `col` (multiple of 8) of `row` inside a run of slabs
__device__ __forceinline__ uint32_t slab_off(int col, int row) {
  return (uint32_t)(col >> 6) * kSlabBytes + slab_chunk_offset(row, (col & 63) >> 3);
}
// "row-interleaved" save layout of a sine argument tile (read back only by the backward epilogue,
// never by the tensor core): 16-byte chunk c (8 columns) of all 128 rows is contiguous, so a warp's
// store covers 512 consecutive bytes.  Same size as the slab layout.
__device__ __forceinline__ uint32_t xsave_off(int col, int row) { return ((uint32_t)(col >> 3) * 128u + row) * 16u; }

// Sum per-row partial results over the 4 column groups through shared scratch ([i][group][row] floats).
// After the call the group-0 thread of each row holds the totals.  The caller provides the barrier
// that protects the scratch before its next use.
template <int NV>
__device__ __forceinline__ void reduce_groups(float* scratch, float (&v)[NV], int cg, int row) {
#pragma unroll
  for (int i = 0; i < NV; ++i) scratch[(i * kColGroups + cg) * kTileM + row] = v[i];
  epi_bar_sync();
  if (cg == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      v[i] = (scratch[(i * kColGroups + 0) * kTileM + row] + scratch[(i * kColGroups + 1) * kTileM + row]) +
             (scratch[(i * kColGroups + 2) * kTileM + row] + scratch[(i * kColGroups + 3) * kTileM + row]);
  }
}

}  // namespace roles
