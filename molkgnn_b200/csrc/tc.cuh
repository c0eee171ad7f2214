// sm_100a tensor-core plumbing for the molkgnn_b200 kernels: tcgen05 (UMMA) instruction / shared-memory descriptors,
// TMEM allocation and loads, mbarrier completion, async-proxy fences.  Inline PTX only (no CUTLASS).
//
// Shared-memory operand layout used everywhere in this library ("interleaved", SWIZZLE_NONE): a tile of R rows x C
// 16-bit elements is stored as   [R/8 row groups][C/8 column chunks][8 rows][8 elements = 16 B]
// i.e. every 8x8 core matrix is 128 contiguous bytes.  The same bytes serve
//   * as a K-major operand  (rows = M or N index, columns = K):  LBO = 128 (next K chunk), SBO = (C/8)*128 (next row group)
//   * as an MN-major operand (rows = K index, columns = M or N): LBO = (C/8)*128 (next K group), SBO = 128 (next M/N chunk)
// One fp16 MMA consumes K = 16: two K chunks (K-major, +256 B per step) or two row groups (MN-major, +2*(C/8)*128 B).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace mk {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (r, c) of a [R x C] 16-bit tile in the interleaved layout; C % 8 == 0
__host__ __device__ __forceinline__ uint32_t il_off(int r, int c, int C) {
    return (uint32_t)((r >> 3) * (C >> 3) + (c >> 3)) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(c & 7) * 2u;
}
__host__ __device__ __forceinline__ uint32_t il_tile_bytes(int R, int C) { return (uint32_t)((R + 7) / 8) * (uint32_t)(C / 8) * 128u; }

// ---- mbarrier -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// suspend-time hint of try_wait: a waiting thread sleeps in the hardware until the phase completes or this many nanoseconds have
// passed, instead of coming back every few hundred cycles (the spin loops of the ring / MMA warps were 19 % of the instructions the
// fused forward executed -- issue slots taken from the consumer warps)
constexpr uint32_t MBAR_SUSPEND_NS = 20000u;
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(MBAR_SUSPEND_NS)
        : "memory");
    return ok != 0;
}
// Bounded wait: a lost completion traps the kernel (reported as a launch failure) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 21); ++it)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}

// generic-proxy shared-memory writes -> visible to the tensor core (async proxy)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM -----------------------------------------------------------------------------------------------------
// warp-collective; ncols power of two in 32..512; the base address lands in *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// this warp's 32 TMEM lanes (warp w of the CTA owns lanes 32*(w%4) ..), `n` consecutive 32-bit columns from `col`
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
    return v;
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// store 8 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors -------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_NONE, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor, kind::f16, fp16 A and B, fp32 accumulate; major: 0 = K-major, 1 = MN-major
__host__ __device__ __forceinline__ uint32_t idesc_f16(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M = 128 lanes, K-major: lane = row, 32-bit column c holds the fp16 pair
// (k = 2c, 2c + 1)) is read from tensor memory; one K = 16 step consumes 8 columns
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared memory -> tensor memory: 128 rows x 256 bits (= one K = 16 step of an fp16 A operand: lane = row, 8 columns) described by
// a K-major SWIZZLE_NONE matrix descriptor (two 16-byte chunks per row, LBO apart; 8-row groups SBO apart).  Issued by ONE
// thread; ordered with the tcgen05.mma instructions the same thread issues afterwards (implicit tensor-core pipeline).
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// arrive on `bar` when every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- fp32 -> (hi, lo) fp16 pair:  v ~= hi + lo * 2^-11,  |error| <= 2^-24 |v|  (|v| <= 1 here) ----
constexpr float LO_SCALE = 2048.0f;          // 2^11
constexpr float LO_UNSCALE = 1.0f / 2048.0f;
__device__ __forceinline__ void split_h(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn((v - __half2float(hi)) * LO_SCALE);
}

// unscaled split: v ~= hi + lo (lo usually subnormal; |error| <= 2^-25 absolute for |v| <= 1)
__device__ __forceinline__ void split_u2(float a, float b, __half2& hi, __half2& lo) {
    hi = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(hi);
    lo = __floats2half2_rn(a - hf.x, b - hf.y);
}

}  // namespace tc
}  // namespace mk
