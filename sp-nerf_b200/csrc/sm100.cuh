// sm_100a primitives used by the SP-NeRF point-network kernels: mbarrier, bulk async copy (TMA
// engine, 1-D form), tcgen05 tensor-memory allocation / MMA / load, UMMA shared-memory and
// instruction descriptors.  Everything is inline PTX; there is no library dependency.
//
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor" / "instruction
// descriptor" tables for .kind::f16.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace sm100 {

// ---------------------------------------------------------------------------------------------
// error flag + bounded waits.  A wait that never completes would wedge the GPU box, so every
// spin is bounded by a wall-clock budget; on expiry the CTA records a code and traps.
// ---------------------------------------------------------------------------------------------
static __device__ unsigned int g_watchdog_code = 0;  // one copy per translation unit

// Each .cu that can trap defines a host getter for its copy; api.cu ORs them together.
#define SPNERF_DEFINE_WATCHDOG_GETTER(fn)                                   \
  extern "C" unsigned int fn(void) {                                      \
    unsigned int v = 0;                                                   \
    cudaMemcpyFromSymbol(&v, sm100::g_watchdog_code, sizeof(v));          \
    return v;                                                             \
  }

#ifndef SPNERF_WATCHDOG_NS
#define SPNERF_WATCHDOG_NS 4000000000ULL  // 4 s
#endif

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Opt a kernel into a large dynamic shared-memory carve-out.  The attribute is per device AND per kernel function
// (not per function type: kernels with the same parameter list share a pointer type), so the cache is keyed by both;
// it only remembers the largest size configured so far.
}  // namespace sm100
#include <map>
#include <mutex>
#include <utility>
namespace sm100 {
inline cudaError_t set_max_dynamic_smem(const void* kernel, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> configured;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = -1;
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = configured[{kernel, dev}];
  if (bytes <= have) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) have = bytes;
  return e;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: `code` identifies the wait site in g_watchdog_code if it expires.  The clock is only
// consulted every 4096 failed polls (a timer read on every wait would sit on the critical path of
// every producer / issuer / epilogue hand-off).
#ifndef SPNERF_WATCHDOG_CYCLES
#define SPNERF_WATCHDOG_CYCLES 8000000000LL   // ~4 s
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t code = 1) {
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xfff) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > SPNERF_WATCHDOG_CYCLES) {
        atomicCAS(&g_watchdog_code, 0u, code);
#ifdef SPNERF_WATCHDOG_NOTRAP
        return;                      // debugging: let the kernel run on (with garbage) so the code can be read back
#else
        __trap();
#endif
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// CTA pairs (cluster of 2, tcgen05 cta_group::2)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
// arrive on an mbarrier that lives in another CTA of the cluster.  Default (CTA-scope release)
// semantics on purpose: `.release.cluster` compiles to MEMBAR.ALL.GPU, which waits for every
// outstanding memory operation of the SM (the in-flight weight copies, the activation saves) on
// each hand-off.  What the consumer reads after this arrive was written by the async proxy (bulk
// copies, completed on the local mbarrier) or published with fence.proxy.async + bar.sync before.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (the result can be consumed later: the issuer overlaps it with MMA issue)
__device__ __forceinline__ bool mbar_test_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait with cluster-scope acquire (the arrivals come from the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, uint32_t code = 2) {
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 0xfff) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > SPNERF_WATCHDOG_CYCLES) {
        atomicCAS(&g_watchdog_code, 0u, code);
#ifdef SPNERF_WATCHDOG_NOTRAP
        return;
#else
        __trap();
#endif
      }
    }
  }
}

// Register re-balancing between warpgroups (4 consecutive warps): the producer / issuer warps need
// few registers, the epilogue warps are pressed against the 65536 / 640 = 96 launch limit.
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// One lane of a fully converged warp.  The single-thread instructions (tcgen05.mma / commit, bulk
// copies) are issued under this predicate while the surrounding loop stays warp-uniform, so their
// operands live in uniform registers instead of being broadcast lane by lane.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// generic-proxy writes (st.shared) -> visible to the async proxy (UMMA / bulk copy reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// bulk async copies (TMA engine, linear form): SASS UBLKCP
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// the same with an L2 eviction-priority hint: policy from l2_policy_evict_last() / _first()
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
// pull a global range into L2 ahead of its (latency-bound) consumer
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------
// tensor memory
// ---------------------------------------------------------------------------------------------
// whole-warp, .sync.aligned
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// pair-wide allocation: the same warp of both CTAs executes these
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues on behalf of the CTA.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread -> one arrive on `bar` when they retire
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Pair MMA: M = 256 (128 rows per CTA, each CTA supplies its own A rows and half of the B rows),
// issued by one thread of the leader CTA (rank 0); the descriptors name the same shared-memory
// offsets in both CTAs.
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued pair MMAs -> one arrive on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma2_commit(uint64_t* bar, uint16_t mask = 3) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// TMEM -> registers, 32 lanes x 32-bit, N consecutive columns per thread (thread t of the warp
// reads lane 32*(warp%4)+t).  taddr = (lane << 16) | column.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// descriptors
// ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64 bit):
//   [0,14)  start address >> 4        [16,30) leading-dim byte offset >> 4
//   [32,46) stride-dim byte offset>>4 [46,48) version = 1 (Blackwell)
//   [49,52) base offset (0: tiles are 1024-B aligned)      [61,64) swizzle: 0 none, 2 128B, 4 64B, 6 32B
constexpr uint32_t kSwizzleNone = 0, kSwizzle128B = 2;

__host__ __device__ constexpr uint64_t make_smem_desc_template(uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                               uint32_t swizzle) {
  return (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16) |
         (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32) | (static_cast<uint64_t>(1) << 46) |
         (static_cast<uint64_t>(swizzle & 7) << 61);
}
__device__ __forceinline__ uint64_t smem_desc(uint64_t tmpl, uint32_t smem_addr) {
  return tmpl | static_cast<uint64_t>((smem_addr >> 4) & 0x3fff);
}

// Instruction descriptor for .kind::f16 (32 bit):
//   [4,6) D format (1 = f32)  [7,10) A format (0 f16, 1 bf16)  [10,13) B format
//   [15] A major (0 = K-major, 1 = MN-major)  [16] B major  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t M, uint32_t N, uint32_t a_mn_major,
                                                      uint32_t b_mn_major, uint32_t a_bf16 = 0,
                                                      uint32_t b_bf16 = 0) {
  return (1u << 4) | (a_bf16 << 7) | (b_bf16 << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// The "slab" layout shared by every operand in this library.
// A slab holds R rows x 64 half-precision columns (128 B per row) in the canonical 128-byte-swizzle
// form: row r lives at byte (r/8)*1024 + (r%8)*128, and within it the 16-byte chunk c (0..7) is
// stored at chunk position c ^ (r%8).  Slab bases are 1024-byte aligned.
//  * as a K-major operand  (rows = M or N index, columns = K):  SBO = 1024, K advances 32 B / 16 cols
//  * as an MN-major operand (rows = K index, columns = M or N): SBO = 1024 (8 K-rows), LBO = slab
//    stride (next 64 M/N columns), K advances 2048 B / 16 rows
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ constexpr uint32_t slab_chunk_offset(uint32_t row, uint32_t chunk) {
  return (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk ^ (row & 7u)) << 4);
}

}  // namespace sm100
