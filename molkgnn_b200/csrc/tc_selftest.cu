// Known-answer self test of the tcgen05 plumbing in tc.cuh: one CTA computes D[128,N] = A * B^T on the tensor cores
// from fp16 operands staged in the interleaved shared-memory layout, for K-major and MN-major operands.
// tests/test_tc_gpu.py compares D with a float64 product; this pins the descriptor encodings the conv kernels rely on.
#include "common.cuh"
#include "tc.cuh"

namespace mk {

struct SelfArgs {
    const __half* A; const __half* B; float* D;
    int N, K;          // M is 128
    int a_mn, b_mn;    // operand majors
    int swap;          // swap LBO/SBO (diagnostic)
};

__global__ void __launch_bounds__(128, 1) k_tc_selftest(SelfArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int M = 128, N = a.N, K = a.K;
    // tile geometry: K-major operand is [rows = M|N][cols = K]; MN-major operand is [rows = K][cols = M|N]
    const int aR = a.a_mn == 1 ? K : M, aC = a.a_mn == 1 ? M : K;      // a_mn 2 / 3: staged K-major like 0
    const int bR = a.b_mn ? K : N, bC = a.b_mn ? N : K;
    unsigned char* As = smem;
    unsigned char* Bs = smem + tc::il_tile_bytes(aR, aC);
    for (int i = tid; i < aR * aC; i += 128) {
        int r = i / aC, c = i % aC;
        *reinterpret_cast<__half*>(As + tc::il_off(r, c, aC)) = a.A[i];
    }
    for (int i = tid; i < bR * bC; i += 128) {
        int r = i / bC, c = i % bC;
        *reinterpret_cast<__half*>(Bs + tc::il_off(r, c, bC)) = a.B[i];
    }
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&tslot, 512);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tslot;
    const bool a_tmem = a.a_mn == 2;             // A operand in tensor memory (columns 256..), K-major, lane = row
    if (a_tmem) {
        for (int c0 = 0; c0 < K / 2; c0 += 8) {
            uint32_t v[8];
            for (int i = 0; i < 8; ++i) {
                const __half2 h2 = __halves2half2(a.A[(size_t)tid * K + 2 * (c0 + i)], a.A[(size_t)tid * K + 2 * (c0 + i) + 1]);
                v[i] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            tc::tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + 256u + (uint32_t)c0, v);
        }
        tc::tmem_st_wait();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }
    const bool a_cp = a.a_mn == 3;               // A operand copied shared -> tensor memory by tcgen05.cp (stack_fwd_fused.cu)
    if (tid == 0 && (a_tmem || a_cp)) {
        const uint32_t idesc = tc::idesc_f16(M, N, 0, a.b_mn);
        if (a_cp)
            for (int ks = 0; ks < K / 16; ++ks)
                tc::tmem_cp_128x256b(tmem + 256u + (uint32_t)(ks * 8), tc::smem_desc(tc::smem_u32(As) + ks * 256u, 128u, (aC / 8) * 128u));
        const uint32_t b_lbo = a.b_mn ? (bC / 8) * 128 : 128, b_sbo = a.b_mn ? 128 : (bC / 8) * 128;
        const uint32_t b_step = a.b_mn ? 2 * (bC / 8) * 128 : 256;
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t bd = tc::smem_desc(tc::smem_u32(Bs) + ks * b_step, b_lbo, b_sbo);
            tc::umma_f16_ts(tmem, tmem + 256u + (uint32_t)(ks * 8), bd, idesc, ks > 0 ? 1u : 0u);
        }
        tc::umma_commit(&bar);
    } else if (tid == 0) {
        const uint32_t idesc = tc::idesc_f16(M, N, a.a_mn, a.b_mn);
        uint32_t a_lbo = a.a_mn ? (aC / 8) * 128 : 128, a_sbo = a.a_mn ? 128 : (aC / 8) * 128;
        uint32_t b_lbo = a.b_mn ? (bC / 8) * 128 : 128, b_sbo = a.b_mn ? 128 : (bC / 8) * 128;
        const uint32_t a_step = a.a_mn ? 2 * (aC / 8) * 128 : 256, b_step = a.b_mn ? 2 * (bC / 8) * 128 : 256;
        if (a.swap) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t ad = tc::smem_desc(tc::smem_u32(As) + ks * a_step, a_lbo, a_sbo);
            const uint64_t bd = tc::smem_desc(tc::smem_u32(Bs) + ks * b_step, b_lbo, b_sbo);
            tc::umma_f16(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
        }
        tc::umma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    const int row = warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tc::tmem_ld_wait();
        for (int i = 0; i < 16; ++i) a.D[(size_t)row * N + c0 + i] = __uint_as_float(v[i]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace mk

using namespace mk;

extern "C" int molkgnn_tc_selftest(const void* A, const void* B, float* D, int32_t N, int32_t K, int32_t a_mn,
                                   int32_t b_mn, int32_t swap, void* stream_) {
    MK_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 256, "tc_selftest: bad N=%d K=%d", N, K);
    SelfArgs a{(const __half*)A, (const __half*)B, D, N, K, a_mn, b_mn, swap};
    const int aR = a_mn == 1 ? K : 128, aC = a_mn == 1 ? 128 : K, bR = b_mn ? K : N, bC = b_mn ? N : K;
    const int smem = tc::il_tile_bytes(aR, aC) + tc::il_tile_bytes(bR, bC);
    MK_CHECK_CUDA(cudaFuncSetAttribute(k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    count_launches(1);
    k_tc_selftest<<<1, 128, smem, (cudaStream_t)stream_>>>(a);
    MK_CHECK_CUDA(cudaGetLastError());
    return 0;
}
