// common.cuh -- shared helpers for the sm_100a kernels of libahrag_b200 (error plumbing + PTX wrappers).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

#include <nvtx3/nvToolsExt.h>

#include "../../include/ahrag_b200.h"

namespace ahr {

// NVTX range over a phase of an API call (host-side enqueue; tools map it onto the kernels launched inside).  Header-only
// NVTX v3: a no-op unless a profiler is attached.
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
#define AHR_NVTX_CAT2(a, b) a##b
#define AHR_NVTX_CAT(a, b) AHR_NVTX_CAT2(a, b)
#define AHR_NVTX(name) ahr::NvtxRange AHR_NVTX_CAT(ahr_nvtx_range_, __LINE__)(name)

// ---------------------------------------------------------------------------------------------
// error plumbing (thread-local message, C return codes)
// ---------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define AHR_CUDA_CHECK(expr)                                                                     \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      ahr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return AHR_E_CUDA;                                                                         \
    }                                                                                            \
  } while (0)

#define AHR_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      ahr::set_error(__VA_ARGS__);  \
      return AHR_E_INVALID;         \
    }                               \
  } while (0)

#define AHR_KERNEL_CHECK()                  \
  do {                                      \
    ahr::count_launch();                    \
    AHR_CUDA_CHECK(cudaGetLastError());     \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    target = dev;
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != target) cudaSetDevice(prev);
  }
  int target = -1;
};

// Device-side invariant checks of the debug build (python -m ah_rag_b200._build --debug -> libahrag_b200_dbg.so, selected
// with AHR_LIB_DEBUG=1).  compute-sanitizer is closed on the GPU pool this library is developed on, so out-of-range
// indices are caught by these assertions on small cases whose results are compared with the CPU oracle
// (tools/sanitize_cases.py): a violated bound prints its location and traps.  Compiled out of the release build.
#ifdef AHR_DEBUG_BOUNDS
#define AHR_DEV_ASSERT(cond)                                                                         \
  do {                                                                                               \
    if (!(cond)) {                                                                                   \
      printf("AHR_DEV_ASSERT failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, \
             (int)blockIdx.x, (int)threadIdx.x);                                                     \
      __trap();                                                                                      \
    }                                                                                                \
  } while (0)
#else
#define AHR_DEV_ASSERT(cond) ((void)0)
#endif

// ranking comparator shared by every selection kernel: (score desc, id asc)
__host__ __device__ __forceinline__ bool ranks_before(float s1, int64_t i1, float s2, int64_t i2) {
  return (s1 > s2) || (s1 == s2 && i1 < i2);
}

__device__ __forceinline__ float bf16_bits_to_float(uint16_t h) { return __uint_as_float(((uint32_t)h) << 16); }

// ---------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA, tcgen05 (sm_100a)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
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
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void prefetch_tensormap(const void *desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}

// 2D tiled TMA load global -> shared::cta, completion on an mbarrier (bytes), with an L2 cache hint
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const void *desc, uint64_t *bar, int32_t c0,
                                            int32_t c1, uint64_t cache_hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "l"(cache_hint)
      : "memory");
}
// L2 cache-hint encodings (same constants CUTLASS's TMA::CacheHintSm90 uses)
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

// ---- tcgen05 ----
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16 inputs, fp32 accumulate), single CTA
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread t of the warp = TMEM lane base+t)
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
        "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
        "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- cluster / CTA-pair (cta_group::2) variants ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` as seen in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
// Remote arrive WITHOUT release semantics: used by the epilogue to hand a TMEM accumulator back to the MMA issuer.
// Nothing the epilogue wrote is read by the waiter; the only hazard is the MMA overwriting TMEM columns still being
// read, and tcgen05.wait::ld (+ tcgen05.fence::before_thread_sync) has retired those reads before this instruction.
// The .release.cluster form compiles to ERRBAR + a cluster-scope membar: ~1.6 k cycles per tile on the accumulator's
// critical path (17 % of the epilogue warps' samples in profiles/r1_c4_pairs_v6_ncu.md).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
// TMA load issued by either CTA of a pair: data lands in the issuing CTA's smem, completion bytes are
// credited to the mbarrier at `cluster_bar_addr` (the leader CTA's full barrier)
__device__ __forceinline__ void tma_load_2d_2sm(void *smem_dst, const void *desc, uint32_t cluster_bar_addr, int32_t c0,
                                                int32_t c1, uint64_t cache_hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(cluster_bar_addr), "r"(c0), "r"(c1),
        "l"(cache_hint)
      : "memory");
}
// L2 prefetch of one TMA box (no shared memory, no barrier)
__device__ __forceinline__ void tma_prefetch_2d(const void *desc, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(desc)),
               "r"(c0), "r"(c1)
               : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t *smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// pair MMA: D (256 x N, 128 rows in each CTA's TMEM) (+)= A (each CTA's 128 rows) * B (each CTA's N/2 rows)
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior MMAs of this thread completed) on the barrier at the same smem offset in every CTA of mask
__device__ __forceinline__ void umma_commit_2sm(uint64_t *bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// UMMA descriptors (sm_100): K-major operand tile in the canonical 128-byte-swizzle layout TMA writes
// (rows of 64 bf16 = 128 B; 8-row swizzle atoms of 1024 B, consecutive atoms 1024 B apart)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address, 16-byte units (14 bits)
  d |= (uint64_t)1 << 16;                       // leading byte offset: unused for swizzled K-major, canonical 1
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset: 8 rows * 128 B between swizzle atoms
  d |= (uint64_t)1 << 46;                       // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                       // layout type: SWIZZLE_128B
  return d;
}
// instruction descriptor for kind::f16: bf16 x bf16 -> fp32, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t m, uint32_t n) {
  return (1u << 4)            // c_format  = F32
         | (1u << 7)          // a_format  = BF16
         | (1u << 10)         // b_format  = BF16
         | ((n >> 3) << 17)   // n_dim
         | ((m >> 4) << 24);  // m_dim
}

}  // namespace ahr
