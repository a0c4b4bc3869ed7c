// Shared helpers for the kfb (kronfluence-B200) CUDA library: error plumbing, PTX wrappers for
// mbarrier / TMA / tcgen05 / TMEM on sm_100a.  Everything here is device- or host-inline; the
// C ABI lives in kfb_api.cu and is declared in include/kfb.h.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/kfb.h"

namespace kfb {

// ---------------------------------------------------------------------------------------------
// Host-side error plumbing.  Every C-ABI entry point returns KFB_OK or a negative code and leaves
// a human-readable message retrievable through kfb_last_error().
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define KFB_CUDA_TRY(expr)                                                                    \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::kfb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,      \
                       __LINE__);                                                             \
      return (_e == cudaErrorMemoryAllocation) ? KFB_ERR_OOM : KFB_ERR_CUDA;                  \
    }                                                                                         \
  } while (0)

#define KFB_REQUIRE(cond, ...)                                                                \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      ::kfb::set_error(__VA_ARGS__);                                                          \
      return KFB_ERR_INVALID;                                                                 \
    }                                                                                         \
  } while (0)

#define KFB_TRY(expr)                                                                         \
  do {                                                                                        \
    int _rc = (expr);                                                                         \
    if (_rc != KFB_OK) return _rc;                                                            \
  } while (0)

static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }
static inline long long round_up_ll(long long a, long long b) { return ceil_div_ll(a, b) * b; }

// Number of SMs of the current device (cached per device).
int sm_count();

#if defined(__CUDACC__)

// ---------------------------------------------------------------------------------------------
// Device-side PTX wrappers (sm_100a only).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}

// Bounded wait: a protocol bug must surface as a trapped kernel (CUDA error), never as a hung GPU.
#ifndef KFB_WATCHDOG_CYCLES
#define KFB_WATCHDOG_CYCLES 6000000000ll
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > KFB_WATCHDOG_CYCLES) __trap();
  }
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}

// TMA: 3-D tiled bulk tensor load global -> shared, completion signalled on an mbarrier.
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* tmap, uint32_t bar, uint32_t dst,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// TMA: 3-D tiled bulk tensor store shared -> global (bulk async-group completion).
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tmap, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// Register re-allocation between warpgroups (all four warps of a warpgroup execute it).
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// tcgen05 / TMEM ------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem]; bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}


// ---- 2-CTA (cta_group::2) variants: a CTA pair on one TPC shares one 256-row MMA ---------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (+expect_tx) on the same-offset mbarrier of CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t bar, uint32_t cta, uint32_t bytes) {
  asm volatile(
      "{\n\t"
      ".reg .b32 rem;\n\t"
      "mapa.shared::cluster.u32 rem, %0, %1;\n\t"
      "mbarrier.arrive.expect_tx.shared::cluster.b64 _, [rem], %2;\n\t"
      "}"
      :
      : "r"(bar), "r"(cta), "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 rem;\n\t"
      "mapa.shared::cluster.u32 rem, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [rem];\n\t"
      "}"
      :
      : "r"(bar), "r"(cta)
      : "memory");
}
// TMA load issued by either CTA of a pair; completion bytes are credited to the LEADER's mbarrier
// (peer bit of the shared::cluster address cleared, as CUTLASS's SM100_TMA_2SM_LOAD does).
__device__ __forceinline__ void tma_load_3d_2sm(const CUtensorMap* tmap, uint32_t bar, uint32_t dst,
                                                int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// Same, multicast: the box lands at the same CTA-relative offset in every CTA of `mask`, and each copy credits the
// mbarrier at the same offset in the LEADER of the destination's pair (CUTLASS SM100_TMA_2SM_LOAD_MULTICAST).
__device__ __forceinline__ void tma_load_3d_2sm_mc(const CUtensorMap* tmap, uint32_t bar, uint32_t dst, int c0, int c1,
                                                   int c2, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
// L2 eviction policies for TMA loads: a streamed operand (used once) must not push a reused one out of L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_3d_2sm_hint(const CUtensorMap* tmap, uint32_t bar, uint32_t dst, int c0, int c1,
                                                     int c2, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm_mc_hint(const CUtensorMap* tmap, uint32_t bar, uint32_t dst, int c0, int c1,
                                                        int c2, uint16_t mask, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      ".L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6, %7;"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "h"(mask),
        "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the same-offset mbarrier of every CTA in `mask` once the pair's MMAs have retired
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}

// TMEM -> registers: this thread's lane, 32 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major shared-memory matrix descriptor (sm_100 "version 1"), hardware swizzle of
// kSwizzleBytes (128 or 64) whose span equals one BLOCK_K row; rows are dense, 8-row groups are
// kSwizzleBytes*8 apart (SBO); LBO is unused for swizzled K-major operands.
template <int kSwizzleBytes>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
  static_assert(kSwizzleBytes == 128 || kSwizzleBytes == 64 || kSwizzleBytes == 32, "swizzle");
  constexpr uint64_t layout = kSwizzleBytes == 128 ? 2 : (kSwizzleBytes == 64 ? 4 : 6);
  constexpr uint64_t sbo = (8 * kSwizzleBytes) >> 4;
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);  // start address, bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                    // LBO (ignored), bits [16,30)
  d |= sbo << 32;                                         // SBO, bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                    // descriptor version = 1 (Blackwell)
  d |= layout << 61;                                      // swizzle mode, bits [61,64)
  return d;
}

// Instruction descriptor for kind::f16: BF16 x BF16 -> FP32 (or FP16 x FP16 -> FP32 for the strict operands), both
// operands K-major, M x N tile.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool f16 = false) {
  return (1u << 4)                                  // D format  = F32
         | ((f16 ? 0u : 1u) << 7)                   // A format  = BF16 (1) / F16 (0)
         | ((f16 ? 0u : 1u) << 10)                  // B format  = BF16 (1) / F16 (0)
         | (static_cast<uint32_t>(N >> 3) << 17)    // N >> 3
         | (static_cast<uint32_t>(M >> 4) << 24);   // M >> 4
}

// Split an fp32 value into bf16 hi + bf16 lo with hi + lo == x to ~2^-17 relative.
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// Strict operands: FP16 planes of x * strict_scale(absmax), a power of two that puts the operand's largest
// magnitude into [2^13, 2^14) (FP16 overflows at 65504).  hi + lo then holds x to max(2^-23 |x|, 2^-39 absmax):
// 22 mantissa bits for every element within 2^-17 of the largest one, degrading gracefully below (FP16
// subnormals), where fp32 itself would only add ~2 bits.  The products of two such planes are exact in fp32.
__host__ __device__ __forceinline__ float strict_scale(float absmax) {
  if (!(absmax > 0.f) || absmax > 3.0e38f) return 1.f;
  int e;
  frexpf(absmax, &e);          // absmax = m * 2^e, m in [0.5, 1)
  return ldexpf(1.f, 14 - e);  // absmax * scale in [2^13, 2^14)
}
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}

#endif  // __CUDACC__

}  // namespace kfb
