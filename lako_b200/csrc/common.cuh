// Shared device/host helpers for the lako_b200 sm_100a kernels: PTX wrappers for mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and UMMA descriptor builders.
// Everything here is hand-written for sm_100a; there is no other code path.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lako_b200.h"  // error codes shared with the C ABI

namespace lako {

#define LAKO_CHECK_LAUNCH()                                  \
  do {                                                       \
    lako::count_launch();                                    \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) {                                \
      lako::set_last_error(cudaGetErrorString(e__), __FILE__, __LINE__); \
      return LAKO_ERR_LAUNCH;                          \
    }                                                        \
  } while (0)

#define LAKO_REQUIRE(cond, code)                             \
  do {                                                       \
    if (!(cond)) {                                           \
      lako::set_last_error("requirement failed: " #cond, __FILE__, __LINE__); \
      return (code);                                         \
    }                                                        \
  } while (0)

void set_last_error(const char* msg, const char* file, int line);
void count_launch();
void count_launches(long long n);  // kernels replayed from a captured graph / captured but not executed
unsigned long long launch_count_value();
bool pdl_enabled();

// Optional per-category device timing (CUDA events on the launching stream), used by bench.py to report
// the dominant kernel's achieved FLOP/s / GB/s.  Disabled by default: zero overhead beyond one branch.
enum ProfCat : int {
  PROF_GEMM = 0,  // large forward GEMMs (M > 128)
  PROF_ENC_ATTN_FWD,
  PROF_ENC_ATTN_BWD,
  PROF_NORM,
  PROF_XATTN,
  PROF_OTHER,
  PROF_GEMM_DGRAD,  // large dgrad (B operand MN-major)
  PROF_GEMM_WGRAD,  // large wgrad (both operands MN-major, fp32 atomic epilogue)
  PROF_GEMM_SMALL,  // decoder-side GEMMs (<= 128 rows of tokens on one side)
  PROF_NCAT
};
struct ProfScope {
  int cat;
  void* rec;
  ProfScope(int cat, double work, cudaStream_t s);
  ~ProfScope();
};

#ifdef __CUDACC__

// Programmatic dependent launch: every kernel releases its successor at entry (the successor's CTAs become
// resident and run their prologue -- barrier init, TMEM allocation, descriptor prefetch -- while this grid is
// still executing) and blocks in pdl_wait() before its first global-memory access until the predecessor grid
// has completed and flushed.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA store / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a pipeline bug must surface as a trapped launch (an error code at the ABI), never as
// a hung GPU.  ~4 s at 2 GHz.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && clock64() - t0 > 8000000000LL) {
      printf("lako: mbarrier wait timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x,
             smem_u32(bar), parity);
      __trap();
    }
  }
}

// Wait of a whole (converged) warp: ONE lane polls, the rest sleep at the warp barrier -- 32 polling lanes are 32
// shared-memory transactions per poll, taken from the MIO bandwidth of the warps that are doing the work being waited for.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
  __syncwarp();
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2D tile load global -> shared, completion on an mbarrier (coords: c0 = inner/contiguous dim)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// 2D tile global -> L2 only (no shared-memory destination, no barrier): warms the L2 for a later tma_load_2d
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void prefetch_l2_line(const void* ptr) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}
// the same with an L2 eviction-priority hint (createpolicy): evict_first for streams that are read once
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// TMA reduce-add store: global[box] += smem[box] (fp32), performed by the memory system on whole lines
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of the (converged) warp: the issuer of TMA / MMA instructions in a warp whose control flow stays uniform
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------ CTA-pair (cta_group::2) variants
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer of THIS CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr),
               "r"(bytes)
               : "memory");
}
// TMA tile load into THIS CTA's smem, completion bytes signalled on an mbarrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_slot, uint32_t ncols) {  // whole warp, both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs of this thread -> arrive on the mbarrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

// instruction descriptor, kind::f16, bf16 x bf16 -> fp32 (bit layout: cute/arch/mma_sm100_desc.hpp InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                                  // c_format = F32
         | (1u << 7)                                // a_format = BF16
         | (1u << 10)                               // b_format = BF16
         | ((a_mn_major ? 1u : 0u) << 15)           // a_major
         | ((b_mn_major ? 1u : 0u) << 16)           // b_major
         | (static_cast<uint32_t>(N >> 3) << 17)    // n_dim
         | (static_cast<uint32_t>(M >> 4) << 24);   // m_dim
}

// shared-memory matrix descriptor, 128B swizzle (bit layout: SmemDescriptor in the same header).
// K-major operand : rows of 128 B (64 bf16 of K), 8-row groups SBO apart (1024 B dense); LBO unused (=16 B).
// MN-major operand: rows of 128 B (64 bf16 of M/N) indexed by k; 8-k groups SBO apart (1024 B);
//                   successive 64-element M/N chunks LBO apart.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread i = lane i of the
// warp's quarter; taddr = (lane_base << 16) | column)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 16-column variant (this warp's 32 lanes x 16 consecutive fp32 columns)
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM, same shape as tmem_ld_32x32
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// 16-column variant
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M = 128 rows = lanes, K-major, two bf16 per 32-bit column) is read from
// tensor memory -- how softmax probabilities written by tcgen05.st feed the P V product without a shared-memory round
// trip.  Issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------------ explicit shared-memory accesses
// The kernels carve one dynamic shared array after rounding its base up to 1024 B through an integer cast; the compiler
// then no longer knows the pointers are shared and emits GENERIC loads / stores (LD.E / ST.E: longer latency, counted
// against the global load queue).  Hot-loop accesses therefore go through these wrappers with 32-bit shared addresses.
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float lds_f(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f(uint32_t saddr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_u4(uint32_t saddr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void red_shared_add_f(uint32_t saddr, float v) {
  asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}

// ------------------------------------------------------------------ small math helpers
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------ dropout (Philox4x32-7, counter based)
// 7 rounds is the smallest Crush-resistant Philox4x32 variant (Salmon et al., SC'11, table 2); dropout
// needs decorrelated bits, not cryptographic strength, and the rounds are pure issue-slot cost in kernels
// that are issue-bound (attention softmax).
// One Philox call decides 8 consecutive elements: element e of a site belongs to group e >> 3 and uses the
// (e & 7)-th 16-bit lane of the 128 random bits; keep iff lane >= thresh (thresh = round(p * 65536)).
// The same (seed, step, site, group) is recomputed in backward, so no mask is ever stored.
struct DropSpec {
  unsigned long long seed;
  unsigned int step;
  unsigned int site;    // unique per (layer, dropout call site)
  unsigned int thresh;  // 0 = dropout disabled
  float scale;          // 1 / (1 - p)
};
__host__ __device__ inline DropSpec make_drop(unsigned long long seed, unsigned step, unsigned site, float p) {
  DropSpec d;
  d.seed = seed;
  d.step = step;
  d.site = site;
  d.thresh = (p > 0.f) ? static_cast<unsigned>(p * 65536.f + 0.5f) : 0u;
  d.scale = (p > 0.f) ? 1.f / (1.f - static_cast<float>(d.thresh) / 65536.f) : 1.f;
  return d;
}
constexpr int PHILOX_ROUNDS = 7;
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < PHILOX_ROUNDS; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// 8-bit keep mask (bit j = keep element group*8 + j)
__device__ __forceinline__ uint32_t drop_keep8(const DropSpec& d, unsigned long long group) {
  const uint4 r = philox4x32(make_uint4(static_cast<uint32_t>(group), static_cast<uint32_t>(group >> 32), d.site, d.step),
                                make_uint2(static_cast<uint32_t>(d.seed), static_cast<uint32_t>(d.seed >> 32)));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    m |= ((w[j] & 0xffffu) >= d.thresh ? 1u : 0u) << (2 * j);
    m |= ((w[j] >> 16) >= d.thresh ? 1u : 0u) << (2 * j + 1);
  }
  return m;
}

// Attention-probability dropout (HF:332), 32 keys at a time: the keep decisions of one (query row, 32-key chunk) as ONE
// bit-sliced threshold comparison.  The 16 words of 4 consecutive Philox counters are the bit planes of 32 independent
// 16-bit uniforms U_b (bit k of U_b = bit b of word k); keep_b = (U_b >= thresh), evaluated from the least significant
// plane up with one 3-input logic op per plane:  ge = thresh_k ? (w_k & ge) : (w_k | ge).  Same Philox work and the same
// keep probability 1 - thresh / 65536 as drop_keep8, but 16 logic ops per 32 keys instead of ~4 per key, and the result
// is already the packed word the backward reads.  Bit b <-> key: key j of the chunk is bit (j >> 1) + 16 (j & 1), so the
// bf16 pair (2i, 2i+1) owns bits i and i + 16: `word << (15 - i)` puts them on the sign bits of the two halfwords and one
// PRMT (sign replicate) turns that into the pair's AND mask.  lako_b200/philox.py (attn_keep_words) restates it.
__device__ __forceinline__ uint32_t drop_keep32_sliced(const DropSpec& d, unsigned long long group4) {
  uint32_t ge = 0xffffffffu;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const unsigned long long g = group4 + u;
    const uint4 r = philox4x32(make_uint4(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), d.site, d.step),
                                  make_uint2(static_cast<uint32_t>(d.seed), static_cast<uint32_t>(d.seed >> 32)));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t tm = 0u - ((d.thresh >> (u * 4 + j)) & 1u);  // warp-uniform: all ones where the threshold bit is set
      ge = (tm & w[j] & ge) | (~tm & (w[j] | ge));
    }
  }
  return ge;
}
// bit of key j (0..31) in a drop_keep32_sliced word
__host__ __device__ constexpr int sliced_bit(int j) { return (j >> 1) + 16 * (j & 1); }
// AND mask of the bf16 pair (keys 2i, 2i+1) of the chunk
__device__ __forceinline__ uint32_t sliced_pair_mask(uint32_t word, int i) {
  uint32_t m;  // prmt's sign-replicate mode (selector nibble bit 3); __byte_perm only honours 3 bits per nibble
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m) : "r"(word << (15 - i)), "r"(0u), "r"(0xbb99u));
  return m;
}

#endif  // __CUDACC__

#ifdef __CUDACC__
// host: launch with the programmatic-stream-serialization attribute (see pdl_wait above)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// ------------------------------------------------------------------ host: TMA descriptor cache (tma_host.cu)
// 2D bf16 row-major tensor [outer][inner] with row pitch `pitch_elems`; box = {box_inner, box_outer},
// 128B swizzle (box_inner must be 64), zero OOB fill.
int get_tensor_map_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                      uint32_t box_inner, uint32_t box_outer);
// 3D bf16 tensor [d2][d1][inner]; strides in elements for d1 and d2.
int get_tensor_map_2d_f32(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                          uint32_t box_inner, uint32_t box_outer);
int get_tensor_map_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t d1, uint64_t d2,
                      uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box_inner, uint32_t box1,
                      uint32_t box2);
int device_sm_count();
int check_arch();

}  // namespace lako
