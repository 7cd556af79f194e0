// Warp roles shared by the forward and backward-data point-network kernels: the weight producer
// and the MMA issuer, both driven by the step list of mlp_pack.cu, and the epilogue-side handshake.
//
// The kernels run as CTA pairs (cluster of 2, tcgen05 cta_group::2): each CTA owns a tile of 128
// points (its activations are the A rows 128*rank.. of an M = 256 MMA) and streams only half of
// every weight tile (B rows (n/2)*rank..), so a weight byte fetched from L2 feeds 256 points.
// 640 threads per CTA, one CTA per SM:
//   warps 0-15    epilogue: warp w reads TMEM lanes 32*(w%4).., i.e. point row 32*(w%4)+lane, and
//                 column group w/4 of the phase's accumulator
//   warp 16       weight producer (one lane): 1-D bulk copies of this CTA's half of the pre-packed
//                 fp16 B tiles into a 4-stage ring
//   warps 17, 18  rank 0: the two MMA issuers (tcgen05.mma M=256, N<=256, K=16, fp16 -> fp32 in TMEM).
//                 Every phase is split into accumulator chunks owned by one issuer each (MmaStep::lane),
//                 interleaved in ring order: one thread cannot issue 4 MMAs + a commit + a barrier
//                 wait in the 512 cycles the tensor pipe needs for them (measured ~710).
//                 rank 1, warp 17: relay, forwards "my ring stage landed" to the issuers
//   warp 19       idle
// The control warps sit ABOVE the epilogue warps: with the same code and the control warps at ids 0-2 the
// kernels were 1-2 % slower (the issuers' few instructions queue behind the epilogue warps of their
// scheduler when the phases overlap at their edges: copy-out, window loads).
#pragma once
#include "sm100.cuh"
#include "net_plan.h"
#include <cstdio>
#include <cstdlib>

// L2 eviction hint of the weight-ring copies: 1 = evict_last.  The blob (5 MB) is re-read by every tile pair while
// gigabytes of saved activations stream through the same L2; pinned like this the backward runs 3.1 % faster
// (alternating A/B, 4.19 -> 4.06 ms), the forward -- whose saves are streaming stores already -- the same.
#ifndef SPNERF_W_POLICY
#define SPNERF_W_POLICY 1
#endif

namespace roles {
using namespace sm100;
using namespace net;

constexpr int kThreads = 640;
constexpr int kEpiThreads = 512;
constexpr int kEpiWarp0 = 0;       // epilogue warps 0..15
constexpr int kProducerWarp = 16, kIssuerWarp0 = 17, kIssuerWarp1 = 18;
constexpr int kColGroups = 4;

struct Smem {
  uint8_t* act;
  uint8_t* aux;
  uint8_t* wst;
  uint64_t* bar_full;    // [2][4] weight stage landed, one set per issuer lane at [lane * 8 + stage] (rank 0: both halves, see setup)
  uint64_t* bar_empty;   // [4] weight stage consumed (arrives from the issuer's commit, both CTAs)
  uint64_t* bar_mma;     // MMA phase retired -> epilogue (both CTAs)
  uint64_t* bar_epi;     // rank 0 only: both epilogues done -> MMA
  uint64_t* bar_par;     // parameter region landed (once)
  uint64_t* bar_free;    // split phases: the A slabs of the first half's columns are no longer read (both issuers) -> epilogue
  uint64_t* bar_half;    // split phases: the first half of the accumulator retired (both issuers) -> epilogue (both CTAs)
  uint32_t* tmem_slot;
  uint32_t rank;
};

__device__ __forceinline__ Smem carve(uint8_t* smem) {
  Smem s;
  s.act = smem;
  s.aux = smem + kOffAux;
  s.wst = smem + kOffWst;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBars);
  s.bar_full = bars; s.bar_empty = bars + 4;
  s.bar_mma = bars + 12; s.bar_epi = bars + 13; s.bar_par = bars + 14; s.bar_half = bars + 15; s.bar_free = bars + 17;      // (bars + 16 holds the TMEM base address)
  s.tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  s.rank = cluster_ctarank();
  return s;
}

// all threads of both CTAs; returns the TMEM base address.  `smallw` (global image of the parameter
// region) is copied into shared memory once; consumers wait on bar_par (parity 0).
__device__ __forceinline__ uint32_t setup(const Smem& s, uint8_t* smem, const float* smallw, int debug = 0) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) { atomicCAS(&g_watchdog_code, 0u, 900u); __trap(); }
    for (int i = 0; i < kNumWStages; ++i) {
      // issuer CTA: its own copy (arrive.expect_tx) + the peer's relay; peer CTA: its own copy only
      mbar_init(&s.bar_full[i], (s.rank == 0 && !(debug & 8)) ? 2 : 1); mbar_init(&s.bar_empty[i], 1);
      mbar_init(&s.bar_full[8 + i], (s.rank == 0 && !(debug & 8)) ? 2 : 1);
    }
    mbar_init(s.bar_mma, 2); mbar_init(s.bar_epi, 2); mbar_init(s.bar_par, 1); mbar_init(s.bar_half, 2); mbar_init(s.bar_free, 2);
    fence_mbar_init();
    mbar_expect_tx(s.bar_par, kSmallWFloats * 4);
    bulk_g2s(smem + kOffRgb2, smallw, 3072 * 4, s.bar_par);              // rgb2 | sem2
    bulk_g2s(smem + kOffSun6, smallw + 3072, 512 * 4, s.bar_par);        // sun6 | beta2
  }
  if (warp == kIssuerWarp0) { tmem_alloc2(s.tmem_slot, 512); tmem_relinquish2(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  return *s.tmem_slot;
}

__device__ __forceinline__ void teardown(uint32_t tmem_base) {
  tc_fence_before();
  cluster_sync_all();      // the peer may still be signalling this CTA's barriers / reading its operands
  if ((threadIdx.x >> 5) == kIssuerWarp0) tmem_dealloc2(tmem_base, 512);
}

// number of tile pairs this cluster processes
__device__ __forceinline__ int64_t my_pairs(int64_t n_pairs) {
  const int64_t c = blockIdx.x >> 1, nc = gridDim.x >> 1;
  return c < n_pairs ? (n_pairs - c + nc - 1) / nc : 0;
}

// warp 0 of both CTAs (all lanes run the loop; one elected lane issues the copies)
__device__ __forceinline__ void producer_loop(const Smem& s, const uint8_t* blob, const StepTable& tab,
                                              int64_t n_iters, int debug, long long* prof = nullptr) {
  uint32_t stage = 0, phase = 0;
  const int n_steps = tab.n;
#if SPNERF_W_POLICY
  const uint64_t wpol = l2_policy_evict_last();      // the weight blob is re-read by every tile pair
#endif
  for (int64_t it = 0; it < n_iters; ++it) {
    for (int i = 0; i < n_steps; ++i) {
      const uint32_t w_off16 = tab.s[i].w_off16;
      const uint32_t bytes = (uint32_t)tab.s[i].bytes16 * 8u;     // half of the item
      // "landed" barriers are per issuer lane: a wait tells mbarrier phases apart by one parity bit only, so every
      // barrier must be waited on, phase after phase, by ONE consumer.  With a barrier per stage and the two issuers'
      // items interleaved unevenly (split phases), an issuer that skipped the other one's use of a stage could take the
      // stage's previous completed phase for its own (seen as a rare hang).
      uint64_t* full = &s.bar_full[(uint32_t)tab.s[i].lane * 8u + stage];
      mbar_wait(&s.bar_empty[stage], phase ^ 1, 10);
      if (prof && it == 2 && i < 256 && blockIdx.x == 0 && (threadIdx.x & 31) == 0) prof[512 + i] = clock64();
      if (elect_one()) {
        if (debug & 1) { mbar_arrive(full); }
        else {
          mbar_expect_tx(full, bytes);
#if SPNERF_W_POLICY
          bulk_g2s_hint(s.wst + stage * kWStageBytes, blob + (size_t)w_off16 * 16 + (size_t)s.rank * bytes, bytes, full, wpol);
#else
          bulk_g2s(s.wst + stage * kWStageBytes, blob + (size_t)w_off16 * 16 + (size_t)s.rank * bytes, bytes, full);
#endif
        }
      }
      __syncwarp();
      if (++stage == kNumWStages) { stage = 0; phase ^= 1; }
    }
  }
  (void)prof;
}

// warp 1, rank 1: tell the issuer that this CTA's half of each item has landed (second arrival on
// the issuer's stage barrier)
__device__ __forceinline__ void relay_loop(const Smem& s, const StepTable& tab, int64_t n_iters, int debug = 0) {
  uint32_t stage = 0, par[2] = {0u, 0u};      // one parity bit per (lane, stage) barrier
  if (debug & 8) return;      // timing experiment: the issuer does not wait for this CTA's operands
  const uint32_t remote0 = mapa_shared(smem_u32(&s.bar_full[0]), 0);
  const int n_steps = tab.n;
  for (int64_t it = 0; it < n_iters; ++it)
    for (int i = 0; i < n_steps; ++i) {
      const uint32_t l = tab.s[i].lane, slot = l * 8u + stage;
      mbar_wait(&s.bar_full[slot], (par[l] >> stage) & 1u, 22);
      par[l] ^= 1u << stage;
      if (elect_one()) mbar_arrive_remote(remote0 + slot * 8u);
      __syncwarp();
      if (++stage == kNumWStages) stage = 0;
    }
}

// warps 1 and 2 of rank 0 (all lanes run the loop; one elected lane issues the MMAs and the commits).
// Both issuers walk the whole step list in ring order and act on the items of their own lane.
__device__ __forceinline__ void mma_loop(const Smem& s, uint32_t tmem_base, const StepTable& tab, int my_lane,
                                         int64_t n_iters, int debug, long long* prof = nullptr) {
  const int n_steps = tab.n;
  constexpr uint64_t tmpl = make_smem_desc_template(16, 1024, kSwizzle128B);
  constexpr uint64_t tmpl_aux = make_smem_desc_template(128, 256, kSwizzleNone);   // 16-column no-swizzle operand
  const uint32_t act_addr = smem_u32(s.act), wst_addr = smem_u32(s.wst), aux_addr = smem_u32(s.aux);
  uint32_t stage = 0, phase = 0, epi_par = 0, full_par = 0;      // full_par: parity bits of this lane's stage barriers
  uint64_t* const my_full = &s.bar_full[my_lane * 8];
  const bool plog = prof && blockIdx.x == 0 && (threadIdx.x & 31) == 0;     // per-step log, iteration 2, both lanes
  long long w_epi = 0, w_full = 0, t_all = prof ? clock64() : 0;
  for (int64_t it = 0; it < n_iters; ++it) {
    int i = 0;
    while (i < n_steps) {
      long long t0 = prof ? clock64() : 0;
      mbar_wait_cluster(s.bar_epi, epi_par, 20); epi_par ^= 1;
      if (prof) w_epi += clock64() - t0;
      tc_fence_after();
      bool last;
      do {
        const uint32_t n = tab.s[i].n, tcol = tab.s[i].tmem_col, a_slab = tab.s[i].a_slab, ksteps = tab.s[i].ksteps;
        const uint32_t first = tab.s[i].first, lane = tab.s[i].lane, half = tab.s[i].half;
        last = tab.s[i].last;
        ++i;
        if ((int)lane == my_lane) {
          t0 = prof ? clock64() : 0;
          mbar_wait_cluster(&my_full[stage], (full_par >> stage) & 1u, 21);      // both halves of the item have landed
          full_par ^= 1u << stage;
          if (prof) {
            const long long t1 = clock64();
            w_full += t1 - t0;
            if (plog && it == 2 && i <= 256) { prof[1024 + i - 1] = t1; prof[768 + i - 1] = t0; }
          }
          tc_fence_after();
          const uint32_t b0 = wst_addr + stage * kWStageBytes;
          const uint32_t idesc = make_idesc_f16(256, n, 0, 0);
          if (elect_one()) {
            const uint32_t acc0 = first ? 0u : 1u;
            if (a_slab == kAuxSlab) {
              const uint64_t ad = smem_desc(tmpl_aux, aux_addr), bd = smem_desc(tmpl_aux, b0);
              if (!(debug & 4)) umma2_f16(tmem_base + tcol, ad, bd, idesc, acc0);
            } else {
              // K = 16 steps: 4 per 64-wide slab; a fused item (mlp_pack.cu merge_last_chunk) runs on over the
              // following A slabs, its B tiles (this CTA's n / 2 rows x 128 bytes per slab) back to back
              const uint32_t a0 = act_addr + a_slab * kSlabBytes, b_slab = (n >> 1) * 128u;
              const uint64_t ad0 = smem_desc(tmpl, a0), bd0 = smem_desc(tmpl, b0);
              if (!(debug & 4)) {
                umma2_f16(tmem_base + tcol, ad0, bd0, idesc, acc0);
                for (uint32_t k = 1; k < ksteps; ++k)
                  umma2_f16(tmem_base + tcol, smem_desc(tmpl, a0 + (k >> 2) * kSlabBytes + (k & 3) * 32),
                            smem_desc(tmpl, b0 + (k >> 2) * b_slab + (k & 3) * 32), idesc, 1u);
              }
            }
            umma2_commit(&s.bar_empty[stage]);
          }
          __syncwarp();
          if (plog && it == 2 && i <= 256) prof[1280 + i - 1] = clock64();       // commit issued
        }
        // split phases: the first half of the accumulator is complete once both issuers' MMAs up to here have retired
        if ((half & 1u) && elect_one()) umma2_commit(s.bar_half);
        __syncwarp();
        // ... and from here on nothing reads the A slabs under the first half's columns
        if ((half & 2u) && elect_one()) umma2_commit(s.bar_free);
        __syncwarp();
        if (++stage == kNumWStages) { stage = 0; phase ^= 1; }
      } while (!last);
      if (elect_one()) umma2_commit(s.bar_mma);      // this issuer's share of the phase (possibly empty) has retired
      __syncwarp();
    }
  }
  if (plog && my_lane == 0) {
    prof[256] = w_epi; prof[257] = w_full; prof[258] = 0; prof[259] = clock64() - t_all;
    prof[260] = n_iters; prof[261] = n_steps;
  }
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// Epilogue-side view of the phase handshake (all epilogue threads call every method).
struct EpiSync {
  const Smem& s;
  bool issuer;
  uint32_t mma_par = 0, half_par = 0, free_par = 0;
  long long* prof;        // optional phase clock log (block 0 only)
  int prof_i = 0;
  uint32_t epi_remote;    // rank 1: the issuer CTA's bar_epi
  __device__ EpiSync(const Smem& s_, long long* prof_)
      : s(s_), issuer(threadIdx.x == kEpiWarp0 * 32),
        prof((blockIdx.x == 0 && threadIdx.x == kEpiWarp0 * 32) ? prof_ : nullptr),
        epi_remote(mapa_shared(smem_u32(s_.bar_epi), 0)) {}

  __device__ __forceinline__ void stamp() {
    if (prof && prof_i < 256) prof[prof_i++] = clock64();
  }
  // Wait for the MMA phase that feeds this epilogue.  One thread polls the mbarrier and the rest
  // block on the named barrier: 512 pollers on one mbarrier starve the producer's and the issuer's
  // own barrier traffic.
  __device__ __forceinline__ void begin() {
    if (issuer) mbar_wait(s.bar_mma, mma_par, 30);
    mma_par ^= 1;
    epi_bar_sync();
    tc_fence_after();
    stamp();
  }
  // split phases: the first half of the accumulator (issuer 0's chunk) is complete
  __device__ __forceinline__ void begin_half() {
    if (issuer) mbar_wait(s.bar_half, half_par, 32);
    half_par ^= 1;
    epi_bar_sync();
    tc_fence_after();
    stamp();
  }
  // split phases of the backward trunk: the remaining MMAs of the phase no longer read the A slabs under the first
  // half's columns (MmaStep::half bit 1), so that half's results may go to shared memory now instead of in the exposed
  // epilogue after the phase.  Waited on exactly once per marked phase (one-bit phase parity).
  __device__ __forceinline__ void wait_first_half_free() {
    if (issuer) mbar_wait(s.bar_free, free_par, 33);
    free_par ^= 1;
    epi_bar_sync();
  }
  // publish shared-memory writes to the async proxy, order TMEM reads, release the MMA warp
  __device__ __forceinline__ void end(bool signal) {
    fence_proxy_async_smem();
    tc_fence_before();
    epi_bar_sync();
    if (issuer && signal) {
      if (s.rank == 0) mbar_arrive(s.bar_epi);
      else mbar_arrive_remote(epi_remote);
    }
    stamp();
  }
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// 16-byte store to the activation / gradient save areas (written once, read milliseconds later by another kernel)
#ifndef SPNERF_STG_MODE
#define SPNERF_STG_MODE 0
#endif

__device__ __forceinline__ void stg16(uint8_t* p, const uint4& v) {
#if SPNERF_STG_MODE == 1
  asm volatile("st.global.cs.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#elif SPNERF_STG_MODE == 2
  asm volatile("st.global.wt.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#elif SPNERF_STG_MODE == 3
  asm volatile("st.global.cg.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#else
  *reinterpret_cast<uint4*>(p) = v;
#endif
}
__device__ __forceinline__ void stg4(uint8_t* p, uint32_t v) {
#if SPNERF_STG_MODE == 1
  asm volatile("st.global.cs.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
  *reinterpret_cast<uint32_t*>(p) = v;
#endif
}

// byte offset of 8 consecutive columns starting at `col` (multiple of 8) of `row` inside a run of slabs
__device__ __forceinline__ uint32_t slab_off(int col, int row) {
  return (uint32_t)(col >> 6) * kSlabBytes + slab_chunk_offset(row, (col & 63) >> 3);
}
// "row-interleaved" save layout of a sine argument tile (read back only by the backward epilogue,
// never by the tensor core): 16-byte chunk c (8 columns) of all 128 rows is contiguous, so a warp's
// store covers 512 consecutive bytes.  Same size as the slab layout.
__device__ __forceinline__ uint32_t xsave_off(int col, int row) { return ((uint32_t)(col >> 3) * 128u + row) * 16u; }

// sign-bit save of a sine layer: one uint32 per (32-column batch, point)
__device__ __forceinline__ uint32_t sbit_off(int col, int row) { return ((uint32_t)(col >> 5) * 128u + row) * 4u; }

// Copy `nslabs` activation slabs (128B-swizzled K-major, as the MMAs read them) to the row-interleaved
// save layout in global memory.  Called by all epilogue threads right after end(): it runs while the
// next MMA phase reads the same slabs, so the stores of a layer overlap the tensor-core work of the
// next one instead of lengthening the epilogue (an SM sustains ~30 B/cycle of global stores).
__device__ __forceinline__ void copy_slabs_out(const uint8_t* act, int slab0, int nslabs, uint8_t* dst) {
  if (!dst) return;
  const int etid = (int)threadIdx.x - kEpiWarp0 * 32;
  for (int idx = etid; idx < nslabs * 1024; idx += kEpiThreads) {
    const int r = idx & 127, c = idx >> 7;          // row fastest: a warp stores 512 contiguous bytes
    const uint4 v = *reinterpret_cast<const uint4*>(act + (size_t)(slab0 + (c >> 3)) * kSlabBytes + slab_chunk_offset(r, c & 7));
    stg16(dst + ((size_t)c * 128 + r) * 16, v);
  }
}

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

// Register re-balancing: 640 threads launch at 96 registers each; the control warpgroup (warps 16-19: producer, two
// issuers / relay, one idle warp) gives most of its share back and the four epilogue warpgroups take it
// (512 * kEpiRegs + 128 * kCtlRegs <= 640 * 96: the pool is what the CTA was given at launch, not the whole register
// file -- 112 / 40 passed ptxas and died with a launch failure).  Every warp of a warpgroup executes the instruction.
#ifndef SPNERF_EPI_REGS
#define SPNERF_EPI_REGS 104
#endif
#ifndef SPNERF_CTL_REGS
#define SPNERF_CTL_REGS 64
#endif
constexpr int kEpiRegs = SPNERF_EPI_REGS, kCtlRegs = SPNERF_CTL_REGS;
static_assert(kEpiThreads * kEpiRegs + (kThreads - kEpiThreads) * kCtlRegs <= kThreads * 96, "register pool of the launch");
// (called inside the role branches: the allocator bounds each region by the setmaxnreg that dominates it)
__device__ __forceinline__ void ctl_registers() {
#if SPNERF_EPI_REGS > 96
  reg_dec<kCtlRegs>();
#endif
}
__device__ __forceinline__ void epi_registers() {
#if SPNERF_EPI_REGS > 96
  reg_inc<kEpiRegs>();
#endif
}

// Phase staggering.  Every cluster runs the same MMA / epilogue sequence on equal tiles, so without a start offset
// all 74 pairs reach their epilogues (the only phases that touch HBM: activation saves, saved-activation reloads,
// gradient tiles) at the same moment and the chip alternates between an idle and a saturated memory system.
// Cluster c therefore starts stagger * (c % groups) / groups cycles late (epilogue warps spin; everything else is
// driven by their first hand-off).
__device__ __forceinline__ void stagger_start(int stagger, int groups) {
  if (stagger <= 0 || groups <= 1) return;
  const long long d = (long long)stagger * (long long)((blockIdx.x >> 1) % groups) / groups;
  const long long t0 = clock64();
  while (clock64() - t0 < d) {}
}
inline void host_stagger(int& stagger, int& groups) {
  stagger = 0; groups = 1;
#ifdef SPNERF_EXPERIMENTS
  if (const char* e = getenv("SPNERF_STAGGER")) { int a = 0, b = 1; if (sscanf(e, "%d,%d", &a, &b) >= 1) { stagger = a; groups = b > 0 ? b : 1; } }
#endif
}

}  // namespace roles
