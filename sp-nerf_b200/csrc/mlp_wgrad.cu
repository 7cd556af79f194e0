// Weight / bias gradients of the point network as tensor-core GEMMs over the point dimension.
//
//   dW[m][i] = sum_points G[point][m] * Y[point][i]
// G (pre-activation gradient tiles, from mlp_bwd.cu) and Y (the forward's saved layer inputs) are
// both stored per 128-point tile as 16-byte chunks (8 features) x 128 points, which is exactly the
// no-swizzle MN-major operand form of tcgen05.mma: no transposition, the tiles are bulk-copied to
// shared memory and multiplied as they are.  Bias gradients and the gradients of the per-ray input
// columns (sun direction, transient embedding, encoded input of the first / skip layer) fall out of
// the same GEMMs through the saved "aux" tile [1, sun(3), t(8)] and the saved encoded input.
//
// Replaces the wgrad half of autograd through models/spnerf.py:305-369.
//
// Shape of the computation (transposed, so that the wide side is the accumulator's columns):
//   D^T[i][m] = sum_p A'[p][i] * B'[p][m]        A' = input side (Y / aux / input), B' = G
// A *job* is one CTA pair (cluster of 2, cta_group::2): M' = 256 rows of A' (128 per CTA) times up to
// 512 columns of B' (two MMAs of N' <= 256; each CTA holds half of the columns of each), i.e. the
// whole 512-column TMEM of both CTAs.  Per 128-point K tile a CTA loads 32 KB of A' and 64 KB of B'
// (2-stage ring) for 2048 cycles of MMA.  Jobs that share B' (the 2-3 pairs of one layer) and a
// point slice are adjacent in the item list, so they run side by side and the second and third
// reader of a G tile hit L2.  Every job is cut into point slices of about equal cost; partial
// tiles go to a workspace and a reduce kernel sums the slices, un-scales and scatters into the
// parameter-shaped gradients.
#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "sm100.cuh"
#include "net_plan.h"

using namespace sm100;
using namespace net;

// L2 eviction hints of the operand stream (0 = none, 1 = everything evict_first, 2 = A' evict_first / B' evict_last,
// 3 = A' normal / B' evict_last).  Alternating A/B on one box: 3.40 / 3.51 / 3.34 ms; 3 measures 1 % behind 2.
#ifndef SPNERF_WGRAD_POLICY
#define SPNERF_WGRAD_POLICY 2
#endif

namespace {

constexpr int kChunkBytes = 2048;              // 8 features x 128 points, fp16
constexpr int kABytes = 16 * kChunkBytes;      // A' part of a stage: 128 rows
constexpr int kBBytes = 32 * kChunkBytes;      // B' part: 2 x 128 columns
constexpr int kStageBytes = kABytes + kBBytes; // 96 KB
constexpr int kStages = 2;
constexpr int kSmemW = kStages * kStageBytes + 256;
constexpr int kWThreads = 256;                 // warp 0 producer, warp 1 issuer / relay, warps 4-7 epilogue
#ifndef SPNERF_WGRAD_ITEMS
#define SPNERF_WGRAD_ITEMS 8
#endif
constexpr int kTargetItemsPerPair = SPNERF_WGRAD_ITEMS;

// `nchunks` consecutive chunks of one save region -> chunk slot `dst_chunk` of the A' or B' part
struct Run { int16_t from_grads, unit, chunk0, nchunks, dst_chunk, _pad; };

struct Job {
  Run a[2][3]; int na[2];       // per CTA rank: its 128 rows of A'
  Run b[2][2];                  // per CTA rank: its half of the columns of MMA j (dst_chunk relative to the B' part)
  int nb;                       // number of MMAs (1 or 2)
  int n[2];                     // N' of MMA j (multiple of 16, <= 256); accumulator columns 256*j ..
  int ncols;                    // n[0] + n[1]: columns of the packed partial tile
  int item0, nitems;            // point slices of this job (job-major copy of the item table, for the reduce)
  int colsum;                   // 1: also sum the columns of B' over the points (bias gradients), row 128 of the tile
};
constexpr int kWsRows = 129;    // 128 accumulator rows + the column-sum row
// One unit of work of a CTA pair: point slice `slice` of `job` (job < 0: padding, nothing to do).
// [peer0, peer0 + npeers) are the items that read the same B' tiles over the same points: they run
// on adjacent pairs at the same time and pace each other (see the gate in the producer).
struct Item { int job, slice, nslices, peer0, npeers, _pad; int64_t ws_off; };      // partial tile: [2][kWsRows][ncols] floats
#ifndef SPNERF_WGRAD_SYNC
#define SPNERF_WGRAD_SYNC 1
#endif
#ifndef SPNERF_WGRAD_GATE_WINDOW
#define SPNERF_WGRAD_GATE_WINDOW 6
#endif
constexpr int kGateWindow = SPNERF_WGRAD_GATE_WINDOW;     // a pair may run at most this many point tiles ahead of its slowest peer

struct Segment {                // scatter rule: rows [row0, row0+nrows) of CTA `rank` x packed columns [col0, col0+ncols)
  int job, rank, row0, nrows, col0, ncols;
  float* dst;                   // dst[r * ld_row + c * ld_col]
  int ld_row, ld_col;
};

// scratch block -> gradient slot copies done (and cleared) by the reduce kernel; see SpnerfMlpWgrad::accum
struct Flush { int src_off, n; float* dst; };
constexpr int kMaxFlush = 12;
struct FlushTable { int n; int n_progress; float* accum; float* absmax_reset; int* progress; Flush f[kMaxFlush]; };

struct WgradParams {
  const uint8_t* saves; const uint8_t* gsaves;
  int64_t save_stride, grad_stride;      // bytes per point tile
  int64_t n_ptiles;
  const Job* jobs; const Item* items; int n_items;
  float* ws;
  int* progress;                         // [n_items] arrival counters of the peer groups, [n_items] pacing words (zeroed before the launch)
  long long* prof;                       // optional counters of pair 0 (debug)
};

__device__ __forceinline__ void item_range(const Item& it, int64_t n_ptiles, int64_t& k0, int64_t& k1) {
  k0 = n_ptiles * it.slice / it.nslices;
  k1 = n_ptiles * (it.slice + 1) / it.nslices;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWThreads, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* bar_full = bars;        // [2]  rank 0: both CTAs' operands landed (own copy + peer relay)
  uint64_t* bar_empty = bars + 2;   // [2]  stage consumed (multicast commit)
  uint64_t* bar_acc = bars + 4;     // accumulator complete -> epilogue (multicast commit)
  uint64_t* bar_drained = bars + 5; // rank 0: both accumulators read out -> issuer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  volatile int* pace_item = reinterpret_cast<volatile int*>(bars + 9);            // item the producer is on (-1: none yet, -2: finished)
  volatile long long* pace_limit = reinterpret_cast<volatile long long*>(bars + 10);   // (item << 32) | tiles the producer may have issued
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) { atomicCAS(&g_watchdog_code, 0u, 901u); __trap(); }
    // a stage is free when its MMAs have retired (commit) and the column-sum warps have read it
    for (int i = 0; i < kStages; ++i) { mbar_init(&bar_full[i], rank == 0 ? 2 : 1); mbar_init(&bar_empty[i], 2); }
    mbar_init(bar_acc, 1); mbar_init(bar_drained, 2);
    *pace_item = -1; *pace_limit = -1;
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc2(tmem_slot, 512); tmem_relinquish2(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- producer: this CTA's rows of A' and its half of B', one stage per 128-point tile ----
#if SPNERF_WGRAD_POLICY
    // L2 hints of the operand stream: the input side (A', read by one pair) leaves first, the gradient side (B', read
    // again by the pair that owns the layer's other rows) stays
    const uint64_t pol_a = SPNERF_WGRAD_POLICY == 3 ? l2_policy_evict_normal() : l2_policy_evict_first();
    const uint64_t pol_b = SPNERF_WGRAD_POLICY >= 2 ? l2_policy_evict_last() : l2_policy_evict_first();
#endif
    uint32_t stage = 0, phase = 0;
    for (int it = pair; it < p.n_items; it += npairs) {
      const Item item = p.items[it];
      if (item.job < 0) continue;
      const Job& job = p.jobs[item.job];
      int64_t k0, k1;
      item_range(item, p.n_ptiles, k0, k1);
      uint32_t bytes = 0;
      for (int r = 0; r < job.na[rank]; ++r) bytes += (uint32_t)job.a[rank][r].nchunks * kChunkBytes;
      for (int j = 0; j < job.nb; ++j) bytes += (uint32_t)job.b[rank][j].nchunks * kChunkBytes;
      // The jobs of one layer and point slice read the same G tiles (each of them once): they start together, so
      // that whichever of them requests a tile second finds it in L2 (and from then on the follower, served from
      // L2, runs faster than the leader and stays with it).  Without the rendezvous the pairs reach their items
      // several tile times apart and every G tile came from HBM twice (18.2 GB per launch against 12.1 GB).
#if SPNERF_WGRAD_SYNC >= 1
      if (item.npeers > 1) {
        if (lane == 0) {
          volatile int* cnt = p.progress + item.peer0;
          if (rank == 0) atomicAdd(p.progress + item.peer0, 1);
          long long t0 = 0; uint32_t spins = 0;
          while (*cnt < item.npeers) {
            if ((++spins & 0xff) == 0) {
              const long long now = clock64();
              if (t0 == 0) t0 = now;
              else if (now - t0 > SPNERF_WATCHDOG_CYCLES) { atomicCAS(&g_watchdog_code, 0u, 45u); __trap(); }
            }
          }
        }
        __syncwarp();
      }
#endif
#if SPNERF_WGRAD_SYNC == 3
      if (rank == 0 && item.npeers > 1 && lane == 0) *pace_item = it;
#endif
      for (int64_t k = k0; k < k1; ++k) {
#if SPNERF_WGRAD_SYNC == 3
        // Pacing: this pair may be at most kGateWindow point tiles ahead of the slowest peer, so that a G tile is still
        // in L2 when the other jobs of the layer ask for it.  The producer only posts its own progress (a store) and
        // reads the allowance from shared memory; the pacing warp below does the global polling off the critical path.
        if (rank == 0 && item.npeers > 1 && lane == 0) {
          volatile int* prog = p.progress + p.n_items;
          prog[it] = (int)(k - k0);
          long long t0 = 0; uint32_t spins = 0;
          for (;;) {
            const long long lim = *pace_limit;
            if ((int)(lim >> 32) == it && (int)(k - k0) <= (int)(lim & 0xffffffffLL)) break;
            if ((++spins & 0xfff) == 0) {
              const long long now = clock64();
              if (t0 == 0) t0 = now;
              else if (now - t0 > SPNERF_WATCHDOG_CYCLES) { atomicCAS(&g_watchdog_code, 0u, 46u); __trap(); }
            }
          }
        }
#elif SPNERF_WGRAD_SYNC == 2
        // pacing: at most kGateWindow point tiles ahead of the slowest peer (progress words behind the arrival counters)
        if (rank == 0 && item.npeers > 1 && lane == 0) {
          volatile int* prog = p.progress + p.n_items;
          prog[it] = (int)(k - k0);
          for (int q = item.peer0; q < item.peer0 + item.npeers; ++q) {
            if (q == it) continue;
            long long t0 = 0; uint32_t spins = 0;
            while ((int)(k - k0) - prog[q] > kGateWindow) {
              if ((++spins & 0xff) == 0) {
                const long long now = clock64();
                if (t0 == 0) t0 = now;
                else if (now - t0 > SPNERF_WATCHDOG_CYCLES) { atomicCAS(&g_watchdog_code, 0u, 46u); __trap(); }
              }
            }
          }
        }
#endif
        __syncwarp();
        mbar_wait(&bar_empty[stage], phase ^ 1, 40);
        if (elect_one()) {
          uint8_t* dst = smem + stage * kStageBytes;
          mbar_expect_tx(&bar_full[stage], bytes);
          for (int r = 0; r < job.na[rank]; ++r) {
            const Run run = job.a[rank][r];
            const uint8_t* src = (run.from_grads ? p.gsaves + k * p.grad_stride : p.saves + k * p.save_stride) +
                                 (size_t)run.unit * kSlabBytes + (size_t)run.chunk0 * kChunkBytes;
#if SPNERF_WGRAD_POLICY
            bulk_g2s_hint(dst + run.dst_chunk * kChunkBytes, src, (uint32_t)run.nchunks * kChunkBytes, &bar_full[stage], pol_a);
#else
            bulk_g2s(dst + run.dst_chunk * kChunkBytes, src, (uint32_t)run.nchunks * kChunkBytes, &bar_full[stage]);
#endif
          }
          for (int j = 0; j < job.nb; ++j) {
            const Run run = job.b[rank][j];
            const uint8_t* src = (run.from_grads ? p.gsaves + k * p.grad_stride : p.saves + k * p.save_stride) +
                                 (size_t)run.unit * kSlabBytes + (size_t)run.chunk0 * kChunkBytes;
#if SPNERF_WGRAD_POLICY
            bulk_g2s_hint(dst + kABytes + run.dst_chunk * kChunkBytes, src, (uint32_t)run.nchunks * kChunkBytes,
                          &bar_full[stage], pol_b);
#else
            bulk_g2s(dst + kABytes + run.dst_chunk * kChunkBytes, src, (uint32_t)run.nchunks * kChunkBytes,
                     &bar_full[stage]);
#endif
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
#if SPNERF_WGRAD_SYNC >= 2
      if (rank == 0 && item.npeers > 1 && lane == 0) { volatile int* prog = p.progress + p.n_items; prog[it] = 1 << 30; }   // done: never holds a peer back
#endif
    }
#if SPNERF_WGRAD_SYNC == 3
    if (rank == 0 && lane == 0) *pace_item = -2;
  } else if (warp == 2 && rank == 0) {
    // ---- pacing warp: polls the peers' progress words and publishes the producer's allowance ----
    if (lane == 0) {
      volatile int* prog = p.progress + p.n_items;
      int cur = -1, peer0 = 0, npeers = 0;
      for (;;) {
        const int it = *pace_item;
        if (it == -2) break;
        if (it < 0) continue;
        if (it != cur) { const Item item = p.items[it]; peer0 = item.peer0; npeers = item.npeers; cur = it; }
        int m = 1 << 29;
        for (int q = peer0; q < peer0 + npeers; ++q)
          if (q != it) { const int v = prog[q]; m = v < m ? v : m; }
        *pace_limit = ((long long)it << 32) | (long long)(unsigned)(m + kGateWindow);
      }
    }
#endif
  } else if (warp == 1 && rank == 1) {
    // ---- relay: second arrival on the issuer's stage barrier ----
    uint32_t stage = 0, phase = 0;
    const uint32_t remote0 = mapa_shared(smem_u32(&bar_full[0]), 0);
    for (int it = pair; it < p.n_items; it += npairs) {
      const Item item = p.items[it];
      if (item.job < 0) continue;
      int64_t k0, k1;
      item_range(item, p.n_ptiles, k0, k1);
      for (int64_t k = k0; k < k1; ++k) {
        mbar_wait(&bar_full[stage], phase, 44);
        if (elect_one()) mbar_arrive_remote(remote0 + stage * 8u);
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ---- issuer: 8 K-steps x (1 or 2) pair MMAs per stage ----
    // saved tiles are row-interleaved: core matrices 128 B apart along K (points), 2048 B along M/N
    constexpr uint64_t tmpl = make_smem_desc_template(128, kChunkBytes, kSwizzleNone);
    uint32_t stage = 0, phase = 0, drained_par = 0;
    bool first_item = true;
    long long* prof = (lane == 0 && p.prof) ? p.prof + 8 * pair : nullptr;
    long long w_full = 0, w_drained = 0, t_issue = 0, n_tiles = 0, t_all = prof ? clock64() : 0;
    for (int it = pair; it < p.n_items; it += npairs) {
      const Item item = p.items[it];
      if (item.job < 0) continue;
      const Job& job = p.jobs[item.job];
      int64_t k0, k1;
      item_range(item, p.n_ptiles, k0, k1);
      const int nb = job.nb;
      const uint32_t idesc0 = make_idesc_f16(256, job.n[0], 1, 1), idesc1 = make_idesc_f16(256, nb > 1 ? job.n[1] : 16, 1, 1);
      const uint32_t b1_off = (uint32_t)job.b[0][nb > 1 ? 1 : 0].dst_chunk * kChunkBytes;
      long long t0 = prof ? clock64() : 0;
      if (!first_item) { mbar_wait_cluster(bar_drained, drained_par, 41); drained_par ^= 1; tc_fence_after(); }
      if (prof) w_drained += clock64() - t0;
      first_item = false;
      for (int64_t k = k0; k < k1; ++k) {
        t0 = prof ? clock64() : 0;
        mbar_wait_cluster(&bar_full[stage], phase, 42);
        const long long t1 = prof ? clock64() : 0;
        if (prof) { w_full += t1 - t0; ++n_tiles; }
        tc_fence_after();
        const uint32_t a0 = smem_u32(smem + stage * kStageBytes), b0 = a0 + kABytes;
        if (elect_one()) {
#pragma unroll
          for (uint32_t s = 0; s < 8; ++s) {                 // 128 points = 8 K-steps of 16 rows
            const uint32_t acc = (k > k0 || s > 0) ? 1u : 0u;
            umma2_f16(tmem_base, smem_desc(tmpl, a0 + s * 256), smem_desc(tmpl, b0 + s * 256), idesc0, acc);
            if (nb > 1)
              umma2_f16(tmem_base + 256, smem_desc(tmpl, a0 + s * 256), smem_desc(tmpl, b0 + b1_off + s * 256), idesc1, acc);
          }
          umma2_commit(&bar_empty[stage]);
        }
        __syncwarp();
        if (prof) t_issue += clock64() - t1;
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma2_commit(bar_acc);
      __syncwarp();
    }
    if (prof) { prof[0] = w_full; prof[1] = w_drained; prof[2] = t_issue; prof[3] = n_tiles; prof[4] = clock64() - t_all; }
  } else if (warp >= 4) {
    // ---- epilogue warps.  While the MMAs run they sum the columns of this CTA's half of B' over the
    // points (= the bias gradients of the layer, fp32) straight from shared memory; afterwards they
    // move the CTA's 128 accumulator rows to the workspace (the reduce kernel sums the slices).
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = (int)threadIdx.x - 128;          // 0..127
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t drained_remote = mapa_shared(smem_u32(bar_drained), 0);
    uint32_t acc_par = 0, stage = 0, phase = 0;
    for (int it = pair; it < p.n_items; it += npairs) {
      const Item item = p.items[it];
      if (item.job < 0) continue;
      const Job& job = p.jobs[item.job];
      int64_t k0, k1;
      item_range(item, p.n_ptiles, k0, k1);
      // thread -> chunk (8 columns) et / 4 of the 32 B' chunks, point quarter et % 4
      float cs[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) cs[e] = 0.f;
      const int ch = et >> 2, pq = et & 3;
      for (int64_t k = k0; k < k1; ++k) {
        // always follow the ring (an early second arrival on a stage's empty barrier would complete the
        // wrong phase); the sums themselves only for the job that owns the layer's bias
        if (lane == 0) {
          if (rank == 0) mbar_wait_cluster(&bar_full[stage], phase, 46);
          else mbar_wait(&bar_full[stage], phase, 46);
        }
        __syncwarp();
        if (job.colsum) {
          const uint8_t* bt = smem + stage * kStageBytes + kABytes + ch * kChunkBytes + pq * 512;
#pragma unroll 8
          for (int i = 0; i < 32; ++i) {
            const int pp = (i + et) & 31;            // staggered start: 8 consecutive threads hit 8 distinct 16-B slots
            const uint4 v = *reinterpret_cast<const uint4*>(bt + pp * 16);
            const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(h[e]);
              cs[2 * e] += f.x; cs[2 * e + 1] += f.y;
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 128) mbar_arrive(&bar_empty[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (lane == 0) mbar_wait(bar_acc, acc_par, 43);
      __syncwarp();
      acc_par ^= 1;
      tc_fence_after();
      float* tile = p.ws + item.ws_off + (size_t)rank * kWsRows * job.ncols;
      float* dst = tile + (size_t)row * job.ncols;
      if (k1 > k0) {
        int out = 0;
        for (int j = 0; j < job.nb; ++j)
          for (int c0 = 0; c0 < job.n[j]; c0 += 16, out += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + 256 * j + c0, v);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 16; e += 4)
              *reinterpret_cast<uint4*>(dst + out + e) = make_uint4(v[e], v[e + 1], v[e + 2], v[e + 3]);
          }
      } else {
        for (int c0 = 0; c0 < job.ncols; c0 += 4) *reinterpret_cast<uint4*>(dst + c0) = make_uint4(0, 0, 0, 0);
      }
      if (job.colsum) {
        // the 4 point quarters of a chunk sit in adjacent lanes
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 1);
          cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 2);
        }
        if (pq == 0) {
          // chunk ch of the B' part: MMA j = ch / 16, column (ch % 16) * 8 of this CTA's half
          const int j = ch >> 4;
          if (j < job.nb && (ch & 15) * 8 < job.n[j] / 2) {
            float* crow = tile + (size_t)128 * job.ncols + (j == 0 ? 0 : job.n[0]) + rank * (job.n[j] / 2) + (ch & 15) * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) crow[e] = cs[e];
          }
        }
      }
      tc_fence_before();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 128) {
        if (rank == 0) mbar_arrive(bar_drained);
        else mbar_arrive_remote(drained_remote);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(tmem_base, 512);
}

// dst[r * ld_row + c * ld_col] = inv_scale * sum over the job's slices of the workspace partials
// One block row per scatter segment, 32 x 128 (16-byte reads) or 32 x 32 element tiles strided over blockIdx.y.  Reads are coalesced along
// the packed columns of the partial tiles; the parameter tensors are mostly the transpose (dW[m][i] from
// D^T[i][m]: ld_row == 1), so the tile goes through shared memory and the writes are coalesced too.  The slice
// offsets are staged once per block and the slice loop is unrolled for memory-level parallelism (the first
// version took 250 us for 312 MB: a dependent table load per slice and element, strided 4-byte writes).
constexpr int kMaxSlices = 256;
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const Segment* __restrict__ segs, const Job* __restrict__ jobs,
                                                           const Item* __restrict__ items, const float* __restrict__ ws,
                                                           const float* __restrict__ scale, int n_segs,
                                                           const FlushTable* __restrict__ flush) {
  __shared__ float tile[32][33];
  __shared__ float tile4[32][129];      // the 16-byte path's tile (32 rows x 128 columns, padded)
  __shared__ long long offs[kMaxSlices];
  if ((int)blockIdx.x >= n_segs) {
    // the atomically accumulated slots: scratch -> gradient tensors, scratch cleared for the next step
    if (blockIdx.y != 0) return;
    for (int i = threadIdx.x; i < flush->n_progress; i += blockDim.x) flush->progress[i] = 0;   // rendezvous counters of the GEMM kernel
    if (!flush->accum) return;
    for (int k = 0; k < flush->n; ++k) {
      const Flush fl = flush->f[k];
      for (int i = threadIdx.x; i < fl.n; i += blockDim.x) {
        fl.dst[i] = flush->accum[fl.src_off + i];
        flush->accum[fl.src_off + i] = 0.f;
      }
    }
    if (threadIdx.x == 0 && flush->absmax_reset) *flush->absmax_reset = 0.f;
    return;
  }
  const Segment sg = segs[blockIdx.x];
  const Job& job = jobs[sg.job];
  const int ns = job.nitems < kMaxSlices ? job.nitems : kMaxSlices;      // host guarantees nitems <= kMaxSlices
  int misaligned = 0;
  for (int s = threadIdx.x; s < ns; s += blockDim.x) {
    offs[s] = items[job.item0 + s].ws_off;
    misaligned |= (int)(offs[s] & 3);
  }
  const bool vec_ok = __syncthreads_or(misaligned) == 0;
  const float inv = 1.f / *scale;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;                // 32 x 8 threads
  const bool transposed = sg.ld_row == 1 && sg.ld_col != 1;              // dst contiguous along r
  if (vec_ok && ((sg.ncols | sg.col0 | job.ncols) & 3) == 0) {
    // 16-byte reads: a tile is 32 rows x 128 packed columns, a warp reads 512 contiguous bytes of one partial row
    const int tiles_c = (sg.ncols + 127) >> 7, tiles_r = (sg.nrows + 31) >> 5;
    for (int t = blockIdx.y; t < tiles_c * tiles_r; t += gridDim.y) {
      const int r0 = (t / tiles_c) << 5, c0 = (t % tiles_c) << 7;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = r0 + ty + 8 * k, c = c0 + 4 * tx;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool in = r < sg.nrows && c < sg.ncols;
        if (in) {
          const float* base = ws + ((size_t)sg.rank * kWsRows + sg.row0 + r) * job.ncols + sg.col0 + c;
          float4 a0 = acc, a1 = acc;
          int s_ = 0;
          for (; s_ + 2 <= ns; s_ += 2) {
            const float4 u = *reinterpret_cast<const float4*>(base + offs[s_]);
            const float4 v = *reinterpret_cast<const float4*>(base + offs[s_ + 1]);
            a0.x += u.x; a0.y += u.y; a0.z += u.z; a0.w += u.w;
            a1.x += v.x; a1.y += v.y; a1.z += v.z; a1.w += v.w;
          }
          if (s_ < ns) {
            const float4 u = *reinterpret_cast<const float4*>(base + offs[s_]);
            a0.x += u.x; a0.y += u.y; a0.z += u.z; a0.w += u.w;
          }
          acc = make_float4((a0.x + a1.x) * inv, (a0.y + a1.y) * inv, (a0.z + a1.z) * inv, (a0.w + a1.w) * inv);
        }
        if (transposed) {
          float* trow = &tile4[ty + 8 * k][4 * tx];
          trow[0] = acc.x; trow[1] = acc.y; trow[2] = acc.z; trow[3] = acc.w;
        } else if (in) {
          float* d = sg.dst + (size_t)r * sg.ld_row + (size_t)c * sg.ld_col;
          d[0] = acc.x; d[sg.ld_col] = acc.y; d[2 * (size_t)sg.ld_col] = acc.z; d[3 * (size_t)sg.ld_col] = acc.w;
        }
      }
      if (transposed) {
        __syncthreads();
        for (int cc = ty; cc < 128; cc += 8) {
          const int r = r0 + tx, c = c0 + cc;
          if (r < sg.nrows && c < sg.ncols) sg.dst[(size_t)r + (size_t)c * sg.ld_col] = tile4[tx][cc];
        }
        __syncthreads();
      }
    }
    return;
  }
  const int tiles_c = (sg.ncols + 31) >> 5, tiles_r = (sg.nrows + 31) >> 5;
  for (int t = blockIdx.y; t < tiles_c * tiles_r; t += gridDim.y) {
    const int r0 = (t / tiles_c) << 5, c0 = (t % tiles_c) << 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + ty + 8 * k, c = c0 + tx;
      float acc = 0.f;
      if (r < sg.nrows && c < sg.ncols) {
        const float* base = ws + ((size_t)sg.rank * kWsRows + sg.row0 + r) * job.ncols + sg.col0 + c;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int s = 0;
        for (; s + 4 <= ns; s += 4) {
          a0 += base[offs[s]]; a1 += base[offs[s + 1]]; a2 += base[offs[s + 2]]; a3 += base[offs[s + 3]];
        }
        for (; s < ns; ++s) a0 += base[offs[s]];
        acc = ((a0 + a1) + (a2 + a3)) * inv;
      }
      if (transposed) tile[ty + 8 * k][tx] = acc;
      else if (r < sg.nrows && c < sg.ncols) sg.dst[(size_t)r * sg.ld_row + (size_t)c * sg.ld_col] = acc;
    }
    if (transposed) {
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = r0 + tx, c = c0 + ty + 8 * k;
        if (r < sg.nrows && c < sg.ncols) sg.dst[(size_t)r + (size_t)c * sg.ld_col] = tile[tx][ty + 8 * k];
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host: the list of jobs for a configuration
// ---------------------------------------------------------------------------------------------
struct Plan {
  FlushTable flush{};
  std::vector<Job> jobs;
  std::vector<Item> items;      // launch order, followed by a job-major copy
  int n_launch = 0;
  std::vector<Segment> segs;
  int64_t ws_floats = 0;
};

// one block of 128 A' rows: up to 3 runs + what each row range means
struct RowRange { int row0, nrows; int kind; int i0; };   // kind 0: layer-input features i0.., 1: bias, 2: sun dir, 3: t_emb, 4: encoded input
struct Block { std::vector<Run> runs; std::vector<RowRange> rows; };
// one range of B' columns: columns [c0, c0+ncols) of MMA `mma` are output units m0.. of a layer
struct ColRange {
  int mma, c0, ncols;
  float* W; int ld; int m0;      // weight gradient (rows = output units), row stride
  float* bias;                   // bias gradient or nullptr
  int in_cols;                   // width of the layer-input part of W (columns beyond it: sun / t / encoded input)
  bool take_sun, take_t, take_inp;
};
struct BSpec { int from_grads, unit, n; };   // MMA j reads n columns starting at the region's first chunk

// colsum: the layer has no extras block; its bias gradients come from the column sums of B' computed
// by the first job of the group (row 128 of that job's partial tiles, each CTA its own half)
void add_group(Plan& pl, const std::vector<BSpec>& bspec, const std::vector<Block>& blocks,
               const std::vector<ColRange>& cols, std::vector<std::vector<int>>& groups, bool colsum = false) {
  std::vector<int> group;
  for (size_t b0 = 0; b0 < blocks.size(); b0 += 2) {
    Job job{};
    job.nb = (int)bspec.size();
    int dst_chunk = 0;
    for (int j = 0; j < job.nb; ++j) {
      job.n[j] = bspec[j].n;
      const int half_chunks = bspec[j].n / 16;          // (n / 2) columns / 8 per chunk
      for (int r = 0; r < 2; ++r)
        job.b[r][j] = Run{(int16_t)bspec[j].from_grads, (int16_t)bspec[j].unit, (int16_t)(r * half_chunks),
                          (int16_t)half_chunks, (int16_t)dst_chunk, 0};
      dst_chunk += 16;                                   // the second MMA's half starts at chunk 16
    }
    job.ncols = job.n[0] + (job.nb > 1 ? job.n[1] : 0);
    const int jid = (int)pl.jobs.size();
    for (int r = 0; r < 2; ++r) {
      job.na[r] = 0;
      if (b0 + r >= blocks.size()) continue;
      const Block& blk = blocks[b0 + r];
      for (const Run& run : blk.runs) job.a[r][job.na[r]++] = run;
      for (const RowRange& rr : blk.rows)
        for (const ColRange& cr : cols) {
          const int packed = (cr.mma == 0 ? 0 : job.n[0]) + cr.c0;
          Segment s{};
          s.job = jid; s.rank = r; s.row0 = rr.row0; s.nrows = rr.nrows; s.col0 = packed; s.ncols = cr.ncols;
          s.ld_row = 1; s.ld_col = cr.ld;
          if (rr.kind == 0 && cr.W) s.dst = cr.W + (size_t)cr.m0 * cr.ld + rr.i0;
          else if (rr.kind == 1 && cr.bias) { s.dst = cr.bias + cr.m0; s.ld_col = 1; }
          else if (rr.kind == 2 && cr.take_sun && cr.W) s.dst = cr.W + (size_t)cr.m0 * cr.ld + cr.in_cols;
          else if (rr.kind == 3 && cr.take_t && cr.W) s.dst = cr.W + (size_t)cr.m0 * cr.ld + cr.in_cols;
          else if (rr.kind == 4 && cr.take_inp && cr.W) s.dst = cr.W + (size_t)cr.m0 * cr.ld + cr.in_cols + rr.i0;
          else continue;
          pl.segs.push_back(s);
        }
    }
    if (colsum && b0 == 0) {
      job.colsum = 1;
      for (const ColRange& cr : cols) {
        if (!cr.bias) continue;
        const int nj = job.n[cr.mma], base = (cr.mma == 0 ? 0 : job.n[0]);
        for (int r = 0; r < 2; ++r) {          // CTA r summed columns [r * nj/2, (r+1) * nj/2) of the MMA
          const int lo = std::max(cr.c0, r * nj / 2), hi = std::min(cr.c0 + cr.ncols, (r + 1) * nj / 2);
          if (hi <= lo) continue;
          Segment s{};
          s.job = jid; s.rank = r; s.row0 = 128; s.nrows = 1; s.col0 = base + lo; s.ncols = hi - lo;
          s.dst = cr.bias + cr.m0 + (lo - cr.c0); s.ld_row = 0; s.ld_col = 1;
          pl.segs.push_back(s);
        }
      }
    }
    pl.jobs.push_back(job);
    group.push_back(jid);
  }
  groups.push_back(group);
}

Plan make_plan(const SpnerfNetConfig& c, float* const* G, int n_pairs) {
  Plan pl;
  std::vector<std::vector<int>> groups;
  const SaveMap sm = make_save_map(c);
  const GradMap gm = make_grad_map(c);
  const NetDims d = make_dims(c);
  const int F = c.feat, H = F / 2;                  // trunk width (512 or 256), head width
  // B' of a trunk-wide gradient tile: 512 columns are two MMAs of 256 (the second half starts 4 slabs further), 256 one
  auto wide_b = [&](int unit) {
    std::vector<BSpec> v{{1, unit, 256}};
    if (F == 512) v.push_back({1, unit + 4, 256});
    return v;
  };
  auto wide_cols = [&](float* w, int ld, float* bias, int in_cols, bool inp) {
    std::vector<ColRange> v{ColRange{0, 0, 256, w, ld, 0, bias, in_cols, false, false, inp}};
    if (F == 512) v.push_back(ColRange{1, 0, 256, w, ld, 256, bias, in_cols, false, false, inp});
    return v;
  };
  auto W = [&](int slot) { return G[slot]; };
  // 128-row blocks of a saved activation of `width` features starting at unit `unit`
  auto act_blocks = [&](int unit, int width, std::vector<Block>& out) {
    for (int b = 0; b < width / 128; ++b) {
      Block blk;
      blk.runs.push_back(Run{0, (int16_t)unit, (int16_t)(16 * b), 16, 0, 0});
      blk.rows.push_back(RowRange{0, 128, 0, 128 * b});
      out.push_back(blk);
    }
  };
  // extras block: aux tile [1, sun(3), t(8)] (2 chunks) and optionally the encoded input (8 chunks)
  auto extras_block = [&](bool with_inp) {
    Block blk;
    blk.runs.push_back(Run{0, (int16_t)sm.aux, 0, 2, 0, 0});
    blk.rows.push_back(RowRange{kAuxColOne, 1, 1, 0});
    blk.rows.push_back(RowRange{kAuxColSun, 3, 2, 0});
    if (c.beta) blk.rows.push_back(RowRange{kAuxColT, c.t_dim, 3, 0});
    if (with_inp) {
      blk.runs.push_back(Run{0, (int16_t)sm.inp, 0, 8, 2, 0});
      blk.rows.push_back(RowRange{16, d.in_dim < 64 ? d.in_dim : 64, 4, 0});
      const AuxExtra ax = make_aux_extra(c);            // input columns 64..: their high parts are aux columns
      for (int q = 0; q < ax.n; ++q) blk.rows.push_back(RowRange{ax.col_hi[q], 1, 4, 64 + q});
    }
    return blk;
  };
  auto col = [&](int mma, int c0, int ncols, float* w, int ld, int m0, float* bias, int in_cols, bool sun = false,
                 bool t = false, bool inp = false) {
    return ColRange{mma, c0, ncols, w, ld, m0, bias, in_cols, sun, t, inp};
  };

  // trunk
  for (int L = 0; L < 8; ++L) {
    float* w = W(SPNERF_P_FC_W0 + 2 * L);
    float* bptr = W(SPNERF_P_FC_W0 + 2 * L + 1);
    const bool skip = (L == c.skip_layer);
    const int ld = (L == 0) ? d.in_dim : F + (skip ? d.in_dim : 0);
    std::vector<Block> blocks;
    if (L > 0) act_blocks(sm.y[L - 1], F, blocks);
    const bool extras = (L == 0 || skip);           // encoded-input columns need the extras rows
    if (extras) blocks.push_back(extras_block(true));
    add_group(pl, wide_b(gm.G[L]), blocks, wide_cols(w, ld, bptr, L == 0 ? 0 : F, true), groups, !extras);
  }
  {  // feats_from_xyz
    std::vector<Block> blocks;
    act_blocks(sm.y[7], F, blocks);
    add_group(pl, wide_b(gm.g_f), blocks, wide_cols(W(SPNERF_P_FEATS_W), F, W(SPNERF_P_FEATS_B), F, false), groups, true);
  }
  {  // heads reading h: logit_from_label.0 (if any) and sigma_from_xyz.0 (column 4 of the small-gradient tile)
    std::vector<Block> blocks;
    act_blocks(sm.y[7], F, blocks);
    std::vector<ColRange> cols;
    std::vector<BSpec> bs;
    if (c.sem) {
      bs.push_back({1, gm.G_sem, H});
      cols.push_back(col(0, 0, H, W(SPNERF_P_SEM0_W), F, 0, W(SPNERF_P_SEM0_B), F));
    }
    const int mma = (int)bs.size();
    bs.push_back({1, gm.gsmall, 16});
    cols.push_back(col(mma, 4, 1, W(SPNERF_P_SIGMA_W), F, 0, nullptr, F));
    add_group(pl, bs, blocks, cols, groups, true);
  }
  {  // heads reading feats: rgb_from_xyzdir.0 and sun_v_net.0 (input [feats, sun_dir])
    std::vector<Block> blocks;
    act_blocks(sm.f, F, blocks);
    blocks.push_back(extras_block(false));
    std::vector<ColRange> cols = {col(0, 0, H, W(SPNERF_P_RGB0_W), F, 0, W(SPNERF_P_RGB0_B), F),
                                  col(1, 0, H, W(SPNERF_P_SUN0_W), F + 3, 0, W(SPNERF_P_SUN0_W + 1), F, true)};
    add_group(pl, {{1, gm.G_rgb, H}, {1, gm.G_sun[0], H}}, blocks, cols, groups);
  }
  if (c.beta) {  // beta_from_xyz.0: input [feats, t_emb]
    std::vector<Block> blocks;
    act_blocks(sm.f, F, blocks);
    blocks.push_back(extras_block(false));
    std::vector<ColRange> cols = {col(0, 0, H, W(SPNERF_P_BETA0_W), F + c.t_dim, 0, W(SPNERF_P_BETA0_B), F,
                                      false, true)};
    add_group(pl, {{1, gm.G_beta, H}}, blocks, cols, groups);
  }
  for (int j = 1; j < 3; ++j) {  // sun_v_net.2 / .4
    std::vector<Block> blocks;
    act_blocks(sm.sun_y[j - 1], H, blocks);
    std::vector<ColRange> cols = {col(0, 0, H, W(SPNERF_P_SUN0_W + 2 * j), H, 0, W(SPNERF_P_SUN0_W + 2 * j + 1), H)};
    add_group(pl, {{1, gm.G_sun[j], H}}, blocks, cols, groups, true);
  }
  // tiny last layers: rows = hidden activations, columns = the small-gradient tile
  // [g_u(3), g_v, g_sigma_pre, g_beta_pre, 0, 0, g_logit(8)]
  auto small_head = [&](int unit, int c0, int ncols, float* w2) {
    std::vector<Block> blocks;
    act_blocks(unit, H, blocks);
    std::vector<ColRange> cols;
    for (int k = 0; k < ncols; ++k) cols.push_back(col(0, c0 + k, 1, w2 ? w2 + (size_t)k * H : nullptr, 0, 0, nullptr, H));
    add_group(pl, {{1, gm.gsmall, 16}}, blocks, cols, groups);
  };
  small_head(sm.rgb_y, 0, 3, W(SPNERF_P_RGB2_W));
  small_head(sm.sun_y[2], 3, 1, W(SPNERF_P_SUN0_W + 6));
  if (c.beta) small_head(sm.beta_y, 5, 1, W(SPNERF_P_BETA2_W));
  if (c.sem) small_head(sm.sem_y, 8, c.num_sem_classes, W(SPNERF_P_SEM2_W));

  // ---- point slices: about kTargetItemsPerPair items of equal cost per pair ----
  // cost of one 128-point tile of a job in cycles: the larger of its MMA time (8 K-steps of
  // 128 * n / 256 cycles per MMA) and its operand load time (~40 B/cycle/SM from L2)
  std::vector<double> gcost(groups.size());
  double total = 0;
  for (size_t g = 0; g < groups.size(); ++g) {
    double cst = 0;
    for (int jid : groups[g]) {
      const Job& job = pl.jobs[jid];
      int chunks = 0;
      for (int r = 0; r < 2; ++r) {
        int ch = 0;
        for (int k = 0; k < job.na[r]; ++k) ch += job.a[r][k].nchunks;
        for (int j = 0; j < job.nb; ++j) ch += job.b[r][j].nchunks;
        chunks = std::max(chunks, ch);
      }
      double mma = 0;
      for (int j = 0; j < job.nb; ++j) mma += 4.0 * std::max(job.n[j], 32);
      cst += std::max(mma, chunks * (double)kChunkBytes / 40.0) + 100.0;
    }
    gcost[g] = cst;
    total += cst;
  }
  const double target_items = (double)kTargetItemsPerPair * n_pairs;
  for (size_t g = 0; g < groups.size(); ++g) {
    const int nj = (int)groups[g].size();
    const int slices = std::min(kMaxSlices, std::max(1, (int)std::lround(target_items * gcost[g] / total / nj)));
    for (int jid : groups[g]) pl.jobs[jid].nitems = slices;
    // slice-major: the jobs of a group that share a slice are adjacent items, all in the same wave
    // (item i runs on pair i % n_pairs in wave i / n_pairs), padded with empty items where needed
    for (int s = 0; s < slices; ++s) {
      while ((int)(pl.items.size() % n_pairs) + nj > n_pairs) {
        Item pad{};
        pad.job = -1; pad.nslices = 1;
        pl.items.push_back(pad);
      }
      const int peer0 = (int)pl.items.size();
      for (int jid : groups[g]) {
        Item it{};
        it.job = jid; it.slice = s; it.nslices = slices; it.peer0 = peer0;
        it.npeers = nj;
        it.ws_off = pl.ws_floats;
        pl.ws_floats += 2 * kWsRows * (int64_t)pl.jobs[jid].ncols;
        pl.items.push_back(it);
      }
    }
  }
  pl.n_launch = (int)pl.items.size();
  // the reduce kernel walks a job's items through item0 + s: append a job-major copy
  std::vector<Item> by_job;
  for (size_t jid = 0; jid < pl.jobs.size(); ++jid) {
    pl.jobs[jid].item0 = pl.n_launch + (int)by_job.size();
    for (int i = 0; i < pl.n_launch; ++i)
      if (pl.items[i].job == (int)jid) by_job.push_back(pl.items[i]);
  }
  pl.items.insert(pl.items.end(), by_job.begin(), by_job.end());
  return pl;
}

int n_sms() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 1 ? sms : 148;
}

size_t align16(size_t v) { return (v + 15) & ~size_t(15); }
size_t table_bytes(const Plan& pl) {
  return align16(pl.jobs.size() * sizeof(Job)) + align16(pl.items.size() * sizeof(Item)) + align16(pl.segs.size() * sizeof(Segment)) +
         sizeof(FlushTable) + 64;
}
constexpr int64_t kTableRoom = 262144, kProgressRoom = 16384;

struct Tables { Job* jobs; Item* items; Segment* segs; FlushTable* flush; };
Tables table_ptrs(const Plan& pl, void* workspace) {
  uint8_t* tab = static_cast<uint8_t*>(workspace) + (size_t)pl.ws_floats * 4;
  Tables t;
  t.jobs = reinterpret_cast<Job*>(tab);
  tab += align16(pl.jobs.size() * sizeof(Job));
  t.items = reinterpret_cast<Item*>(tab);
  tab += align16(pl.items.size() * sizeof(Item));
  t.segs = reinterpret_cast<Segment*>(tab);
  tab += align16(pl.segs.size() * sizeof(Segment));
  t.flush = reinterpret_cast<FlushTable*>(tab);
  return t;
}

}  // namespace

// host-side sizes of a plan, cached per (configuration, pair count): the launch path must not rebuild the job lists
namespace {
struct PlanCounts { int64_t ws_floats; int n_launch; int n_segs; size_t jobs_bytes, items_bytes, segs_bytes; };
}  // namespace
#include <map>
#include <mutex>
#include <array>
static PlanCounts plan_counts(const SpnerfNetConfig& c, float* const* G, int pairs) {
  static std::mutex mu;
  static std::map<std::array<int64_t, 11>, PlanCounts> cache;
  int64_t present = 0;                       // scatter segments exist only for the gradient slots the caller provides
  for (int i = 0; i < SPNERF_NUM_PARAMS; ++i) present |= (int64_t)(G[i] != nullptr) << i;
  const std::array<int64_t, 11> key{c.feat, c.layers, c.skip_layer, c.mapping, c.sem, c.num_sem_classes, c.emb_dim, c.beta,
                                    c.t_dim, pairs, present};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  const Plan pl = make_plan(c, G, pairs);
  PlanCounts pc{pl.ws_floats, pl.n_launch, (int)pl.segs.size(), align16(pl.jobs.size() * sizeof(Job)),
                align16(pl.items.size() * sizeof(Item)), align16(pl.segs.size() * sizeof(Segment))};
  cache[key] = pc;
  return pc;
}

static long long* g_prof_wgrad = nullptr;
extern "C" void spnerf_debug_counters_wgrad(long long* dev_buf8_per_pair) { g_prof_wgrad = dev_buf8_per_pair; }

extern "C" int64_t spnerf_mlp_wgrad_workspace_bytes(const SpnerfNetConfig* cfg) {
  if (!cfg) return -1;
  float* G[SPNERF_NUM_PARAMS] = {};
  // the plan depends on the SM count only through the slice counts; query-time device = run-time device
  Plan pl = make_plan(*cfg, G, n_sms() / 2);
  return (int64_t)pl.ws_floats * 4 + kTableRoom + kProgressRoom;   // partial tiles + device tables + progress words
}

// Uploads the job / item / scatter tables to the tail of the workspace.  Call once per (configuration,
// gradient pointers, workspace); synchronises the stream (the tables come from pageable memory).
extern "C" int spnerf_mlp_wgrad_prepare(const SpnerfMlpWgrad* a, void* stream_) {
  if (!a || !a->grads_host || !a->workspace) return SPNERF_ERR_BAD_ARG;
  if (!feat_supported(a->cfg.feat) || a->cfg.layers != 8 || a->cfg.skip_layer != 4) return SPNERF_ERR_UNSUPPORTED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Plan pl = make_plan(a->cfg, a->grads_host, n_sms() / 2);
  const size_t tb = table_bytes(pl);
  if ((int64_t)pl.ws_floats * 4 + kTableRoom + kProgressRoom > a->workspace_bytes || (int64_t)tb > kTableRoom ||
      (int64_t)pl.n_launch * 8 > kProgressRoom)
    return SPNERF_ERR_WORKSPACE;
  const Tables t = table_ptrs(pl, a->workspace);
  cudaMemcpyAsync(t.jobs, pl.jobs.data(), pl.jobs.size() * sizeof(Job), cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(t.items, pl.items.data(), pl.items.size() * sizeof(Item), cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(t.segs, pl.segs.data(), pl.segs.size() * sizeof(Segment), cudaMemcpyHostToDevice, stream);
  {
    // slots summed with atomics: staged in the caller's scratch block, flushed by the reduce kernel
    FlushTable& ft = pl.flush;
    ft.accum = a->accum; ft.absmax_reset = a->absmax_reset;
    float* const* G = a->grads_host;
    auto add = [&](int slot, int off, int n) { if (a->accum && G[slot] && n > 0 && ft.n < kMaxFlush) ft.f[ft.n++] = Flush{off, n, G[slot]}; };
    add(SPNERF_P_RGB2_B, SPNERF_ACC_SMALL_BIAS + 0, 3);
    add(SPNERF_P_SUN0_W + 7, SPNERF_ACC_SMALL_BIAS + 3, 1);
    add(SPNERF_P_SIGMA_B, SPNERF_ACC_SMALL_BIAS + 4, 1);
    if (a->cfg.beta) add(SPNERF_P_BETA2_B, SPNERF_ACC_SMALL_BIAS + 5, 1);
    if (a->cfg.sem) {
      add(SPNERF_P_SEM2_B, SPNERF_ACC_SMALL_BIAS + 6, a->cfg.num_sem_classes);
      add(SPNERF_P_SEM_EMB, SPNERF_ACC_EMB, (a->cfg.num_sem_classes + 1) * a->cfg.emb_dim);
    }
    const int nh = a->cfg.feat / 2;
    add(SPNERF_P_SKY0_W, SPNERF_ACC_SKY_W0, 3 * nh);
    add(SPNERF_P_SKY0_B, SPNERF_ACC_SKY_B0, nh);
    add(SPNERF_P_SKY2_W, SPNERF_ACC_SKY_W2, 3 * nh);
    add(SPNERF_P_SKY2_B, SPNERF_ACC_SKY_B2, 3);
    ft.progress = reinterpret_cast<int*>(static_cast<uint8_t*>(a->workspace) + (size_t)pl.ws_floats * 4 + kTableRoom);
    ft.n_progress = pl.n_launch * 2;
    cudaMemsetAsync(ft.progress, 0, (size_t)ft.n_progress * sizeof(int), stream);     // from then on the reduce kernel re-zeroes them
    cudaMemcpyAsync(t.flush, &ft, sizeof(FlushTable), cudaMemcpyHostToDevice, stream);
  }
  cudaError_t e = cudaStreamSynchronize(stream);
  return e == cudaSuccess ? 0 : -(int)e;
}

extern "C" int spnerf_mlp_bwd_weights(const SpnerfMlpWgrad* a, void* stream_) {
  if (!a || !a->saves || !a->grad_saves || !a->scale || !a->grads_host || !a->workspace) return SPNERF_ERR_BAD_ARG;
  if (!feat_supported(a->cfg.feat) || a->cfg.layers != 8 || a->cfg.skip_layer != 4) return SPNERF_ERR_UNSUPPORTED;
  if (a->n_points <= 0) return a->n_points == 0 ? 0 : SPNERF_ERR_BAD_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int pairs = n_sms() / 2;
  const PlanCounts pl = plan_counts(a->cfg, a->grads_host, pairs);     // the tables themselves were uploaded by _prepare
  if ((int64_t)pl.ws_floats * 4 + kTableRoom + kProgressRoom > a->workspace_bytes ||
      (int64_t)pl.n_launch * 8 > kProgressRoom)
    return SPNERF_ERR_WORKSPACE;
  Tables t;
  {
    uint8_t* tab = static_cast<uint8_t*>(a->workspace) + (size_t)pl.ws_floats * 4;
    t.jobs = reinterpret_cast<Job*>(tab); tab += pl.jobs_bytes;
    t.items = reinterpret_cast<Item*>(tab); tab += pl.items_bytes;
    t.segs = reinterpret_cast<Segment*>(tab); tab += pl.segs_bytes;
    t.flush = reinterpret_cast<FlushTable*>(tab);
  }
  WgradParams p;
  p.saves = static_cast<const uint8_t*>(a->saves); p.gsaves = static_cast<const uint8_t*>(a->grad_saves);
  p.save_stride = (int64_t)make_save_map(a->cfg).total * kSlabBytes;
  p.grad_stride = (int64_t)make_grad_map(a->cfg).total * kSlabBytes;
  p.n_ptiles = (a->n_points + kTileM - 1) / kTileM;
  p.jobs = t.jobs; p.items = t.items;
  p.n_items = pl.n_launch;                                // the job-major copy behind them is for the reduce kernel
  p.ws = static_cast<float*>(a->workspace);
  p.progress = reinterpret_cast<int*>(static_cast<uint8_t*>(a->workspace) + (size_t)pl.ws_floats * 4 + kTableRoom);
  p.prof = g_prof_wgrad;
  if (cudaError_t e = sm100::set_max_dynamic_smem(reinterpret_cast<const void*>(wgrad_kernel), kSmemW); e != cudaSuccess) return -(int)e;
  wgrad_kernel<<<2 * std::min(pairs, pl.n_launch), kWThreads, kSmemW, stream>>>(p);
  wgrad_reduce_kernel<<<dim3((unsigned)pl.n_segs + 1, 16), 256, 0, stream>>>(t.segs, t.jobs, t.items, p.ws, a->scale, pl.n_segs,
                                                                             t.flush);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

SPNERF_DEFINE_WATCHDOG_GETTER(spnerf_watchdog_code_wgrad)
