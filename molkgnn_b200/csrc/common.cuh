// Shared helpers for the molkgnn_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/molkgnn_b200.h"

namespace mk {

void set_error(const char* fmt, ...);

#define MK_CHECK_CUDA(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            mk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return -2;                                                                   \
        }                                                                                \
    } while (0)

#define MK_REQUIRE(cond, ...)                                                            \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            mk::set_error(__VA_ARGS__);                                                  \
            return -1;                                                                   \
        }                                                                                \
    } while (0)

constexpr int EP = MOLKGNN_EDGE_PAD;  // padded bond-attribute row

// ---- permutation tables (kernels.py:109-128): lexicographic for d != 4, the 12 even ones for d == 4 ----
template <int D> struct Perm;
template <> struct Perm<1> {
    static constexpr int P = 1;
    __host__ __device__ static constexpr int at(int p, int j) { return 0; }
};
template <> struct Perm<2> {
    static constexpr int P = 2;
    __host__ __device__ static constexpr int at(int p, int j) { return p == 0 ? j : 1 - j; }
};
template <> struct Perm<3> {
    static constexpr int P = 6;
    __host__ __device__ static constexpr int at(int p, int j) {
        constexpr int T[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
        return T[p][j];
    }
};
template <> struct Perm<4> {
    static constexpr int P = 12;
    __host__ __device__ static constexpr int at(int p, int j) {
        constexpr int T[12][4] = {{0, 1, 2, 3}, {0, 2, 3, 1}, {0, 3, 1, 2}, {1, 0, 3, 2}, {1, 2, 0, 3}, {1, 3, 2, 0},
                                  {2, 0, 1, 3}, {2, 1, 3, 0}, {2, 3, 0, 1}, {3, 0, 2, 1}, {3, 1, 0, 2}, {3, 2, 1, 0}};
        return T[p][j];
    }
};
// packed code of permutation p: 2 bits per j -> s = perm[j]
template <int D> __host__ __device__ constexpr uint32_t perm_code(int p) {
    uint32_t c = 0;
    for (int j = 0; j < D; ++j) c |= (uint32_t)Perm<D>::at(p, j) << (2 * j);
    return c;
}
// packed code of the inverse permutation: 2 bits per s -> j with perm[j] == s
template <int D> __host__ __device__ constexpr uint32_t perm_inv_code(int p) {
    uint32_t c = 0;
    for (int j = 0; j < D; ++j) c |= (uint32_t)j << (2 * Perm<D>::at(p, j));
    return c;
}
// The same codes as compile-time tables in constant memory, [degree - 1][permutation]: a kernel prologue copies them to shared
// memory.  (Selecting perm_code<D>(q) with a run-time q materialises Perm<D>'s table on the thread's stack: ~700 local stores in
// the prologue of k_conv_bwd_tile.)
#define MK_PC3(f) f<3>(0), f<3>(1), f<3>(2), f<3>(3), f<3>(4), f<3>(5)
#define MK_PC4(f) f<4>(0), f<4>(1), f<4>(2), f<4>(3), f<4>(4), f<4>(5), f<4>(6), f<4>(7), f<4>(8), f<4>(9), f<4>(10), f<4>(11)
static __constant__ unsigned char c_perm_code[4][12] = {{0}, {perm_code<2>(0), perm_code<2>(1)}, {MK_PC3(perm_code)}, {MK_PC4(perm_code)}};
static __constant__ unsigned char c_perm_inv_code[4][12] = {{0}, {perm_inv_code<2>(0), perm_inv_code<2>(1)}, {MK_PC3(perm_inv_code)},
                                                            {MK_PC4(perm_inv_code)}};
__host__ __device__ inline int num_perms(int d) { return d == 1 ? 1 : d == 2 ? 2 : d == 3 ? 6 : 12; }

// ---- packed (normalised) kernel-set layout of one degree; all offsets in floats ----
struct PackedLayout {
    int d, L, Fp;
    int rows_x;   // (d+1)*L : support rows s*L+k (s<d), centre rows d*L+k
    int64_t sup, es, norm, enorm, w, sign, total;
    __host__ __device__ PackedLayout(int d_, int L_, int Fp_) : d(d_), L(L_), Fp(Fp_) {
        rows_x = (d + 1) * L;
        sup = 0;
        es = sup + (int64_t)rows_x * Fp;
        norm = es + (int64_t)d * L * EP;
        enorm = norm + rows_x;
        w = (enorm + d * L + 3) / 4 * 4;
        sign = w + 8;
        const int64_t base_total = (sign + (L * 12 + 3) / 4 + 3) / 4 * 4;
        // tensor-core operand images (fp16 hi/lo pairs in the interleaved UMMA layout of tc.cuh), 128-byte aligned
        Fk = (Fp + 15) / 16 * 16;
        DS = d == 3 ? 4 : d;
        KSpad = (L * DS + 15) / 16 * 16 + 16;   // + one spare 16-row block: kernel ranges may over-read up to 15 rows
        Lpad = (L + 15) / 16 * 16;
        tc = (base_total + 31) / 32 * 32;
        tc_sup_bytes = (int64_t)(KSpad / 8) * (Fk / 8) * 128;
        tc_cen_bytes = (int64_t)(Lpad / 8) * (Fk / 8) * 128;
        total = tc + (2 * tc_sup_bytes + 2 * tc_cen_bytes) / 4;
    }
    int Fk, DS, KSpad, Lpad;
    int64_t tc, tc_sup_bytes, tc_cen_bytes;
    // byte offsets of the four images from (char*)(packed + tc)
    __host__ __device__ int64_t tc_sup_hi() const { return 0; }
    __host__ __device__ int64_t tc_sup_lo() const { return tc_sup_bytes; }
    __host__ __device__ int64_t tc_cen_hi() const { return 2 * tc_sup_bytes; }
    __host__ __device__ int64_t tc_cen_lo() const { return 2 * tc_sup_bytes + tc_cen_bytes; }
};
// w block: [0]=ws [1]=wc [2]=we [3]=W=ws+wc+we [4..7] reserved

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// x / y for a divisor whose reciprocal ry = RN(1/y) is known: one product, one exact residual, one correction
// (Markstein).  Bit-identical to IEEE division for operands in the normal range, without the MUFU/FCHK sequence and
// its slow path that `x / y` compiles to.
__device__ __forceinline__ float div_by(float x, float y, float ry) {
    const float q = x * ry;
    const float r = fmaf(-y, q, x);
    return fmaf(r, ry, q);
}
// mean over d items: true division by d (torch.mean), exact scalings for d = 1, 2, 4
template <int D> __device__ __forceinline__ float div_deg(float x) {
    if (D == 1) return x;
    if (D == 2) return x * 0.5f;
    if (D == 4) return x * 0.25f;
    return div_by(x, 3.0f, 0x1.555556p-2f);
}

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

void count_launches(int n);
// named CUDA-event scope on the launching stream (no-op unless molkgnn_profile_enable(1)); params.cu
class ProfScope {
public:
    ProfScope(const char* name, cudaStream_t st);
    ~ProfScope();
private:
    int rec_;
    cudaStream_t st_;
};

// ---- optional in-kernel phase clocks (profiling build only: -DMK_PHASE_CLOCKS, tools/phase_clocks.py) ----
// One designated thread accumulates clock64() deltas per phase; barriers make its view the CTA's critical path.
#ifdef MK_PHASE_CLOCKS
#define MK_PH_DECL(on_)                                                     \
    const bool ph_on_ = (on_);                                              \
    unsigned long long ph_last_ = clock64();                                \
    unsigned long long ph_acc_[16];                                         \
    _Pragma("unroll") for (int i_ = 0; i_ < 16; ++i_) ph_acc_[i_] = 0ull;
#define MK_PH(i_)                                                           \
    do {                                                                    \
        if (ph_on_) {                                                       \
            const unsigned long long t_ = clock64();                        \
            ph_acc_[i_] += t_ - ph_last_;                                   \
            ph_last_ = t_;                                                  \
        }                                                                   \
    } while (0)
#define MK_PH_FLUSH(arr_)                                                   \
    do {                                                                    \
        if (ph_on_) {                                                       \
            _Pragma("unroll") for (int i_ = 0; i_ < 16; ++i_)               \
                if (ph_acc_[i_]) atomicAdd(&(arr_)[i_], ph_acc_[i_]);       \
        }                                                                   \
    } while (0)
#else
#define MK_PH_DECL(on_)
#define MK_PH(i_)
#define MK_PH_FLUSH(arr_)
#endif
extern bool g_fwd_counters_zeroed;   // conv_fwd_tile.cu: the caller (stack.cu) zeroed the tile queues of all layers already
int device_num_sms();
int device_index();          // current CUDA device, clipped to 0..15 (index of the per-device caches)
int device_max_smem_optin();

}  // namespace mk
