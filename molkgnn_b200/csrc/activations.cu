// Row norms / zero padding of the layer input and the fused neighbour sum that ends every MolGCN layer.
//   k_pad_norm       x -> zero-padded copy + ||row||  (norm half of torch's cosine_similarity, kernels.py:189-190)
//   k_propagate_fwd  h[i] = sum_{(j->i)} sim_sc[j]     (MessagePassing aggr='add', KernelLayer.py:119-123), reading
//                    only the L_deg(j) non-zero columns of each source row from the compact score blocks, summing in
//                    edge order (deterministic, no atomics), and emitting ||h_i|| for the next layer's cosines.
// Both are pure HBM-bound streaming kernels: one warp per row, float4 where the layout allows.
#include "common.cuh"

namespace mk {

__global__ void k_pad_norm(const float* __restrict__ x, int N, int F, int ldx, float* out, int ldo, float* norm) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= N) return;
    const float* src = x + (size_t)row * ldx;
    float ss = 0.f;
    for (int f = lane; f < ldo; f += 32) {
        float v = f < F ? src[f] : 0.f;
        ss += v * v;
        if (out != x) out[(size_t)row * ldo + f] = v;
    }
    ss = warp_sum(ss);
    if (lane == 0 && norm) norm[row] = sqrtf(ss);
}

struct PropArgs {
    int N, K, ldh;
    int L[4], koff[4];
    long long scoff[4];
    const int* deg; const int* pos; const int* in_cnt; const int* in_src;
    const float* sc;
    float* h; float* hnorm;
};

// one warp per target node; lane owns columns lane, lane+32, ...  (K <= 32*MAXC)
template <int MAXC>
__global__ void __launch_bounds__(256) k_propagate_fwd(const __grid_constant__ PropArgs a) {
    int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (i >= a.N) return;
    float acc[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) acc[c] = 0.f;
    const int cnt = min(a.in_cnt[i], 4);
    for (int t = 0; t < cnt; ++t) {          // edge order
        const int j = a.in_src[4 * i + t];
        const int d = a.deg[j];
        const int L = a.L[d - 1], ko = a.koff[d - 1];
        const float* src = a.sc + a.scoff[d - 1] + (size_t)a.pos[j] * L - ko;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            const int col = lane + 32 * c;
            if (col >= ko && col < ko + L) acc[c] += src[col];
        }
    }
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int col = lane + 32 * c;
        if (col < a.ldh) {
            const float v = col < a.K ? acc[c] : 0.f;
            a.h[(size_t)i * a.ldh + col] = v;
            ss += v * v;
        }
    }
    ss = warp_sum(ss);
    if (lane == 0 && a.hnorm) a.hnorm[i] = sqrtf(ss);
}

}  // namespace mk

using namespace mk;

extern "C" int molkgnn_pad_norm(const float* x, int32_t N, int32_t F, int32_t ldx, float* out, int32_t ldo, float* norm,
                                void* stream_) {
    MK_REQUIRE(ldo >= F, "pad_norm: ldo=%d < F=%d", ldo, F);
    if (N == 0) return 0;
    count_launches(1);
    k_pad_norm<<<(N + 7) / 8, 256, 0, (cudaStream_t)stream_>>>(x, N, F, ldx, out, ldo, norm);
    MK_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int molkgnn_propagate_fwd(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* sc,
                                     const int64_t scoff[4], float* h, int32_t ldh, float* hnorm, void* stream_) {
    PropArgs a;
    a.N = plan->N; a.K = layer->K; a.ldh = ldh;
    MK_REQUIRE(ldh >= layer->K, "propagate_fwd: ldh=%d < K=%d", ldh, layer->K);
    for (int d = 0; d < 4; ++d) { a.L[d] = layer->L[d]; a.koff[d] = layer->koff[d]; a.scoff[d] = scoff[d]; }
    a.deg = plan->deg; a.pos = plan->pos; a.in_cnt = plan->in_cnt; a.in_src = plan->in_src;
    a.sc = sc; a.h = h; a.hnorm = hnorm;
    if (a.N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream_;
    const int grid = (a.N + 7) / 8;
    count_launches(1);
    if (ldh <= 32 * 1) k_propagate_fwd<1><<<grid, 256, 0, st>>>(a);
    else if (ldh <= 32 * 4) k_propagate_fwd<4><<<grid, 256, 0, st>>>(a);
    else if (ldh <= 32 * 8) k_propagate_fwd<8><<<grid, 256, 0, st>>>(a);
    else if (ldh <= 32 * 16) k_propagate_fwd<16><<<grid, 256, 0, st>>>(a);
    else { MK_REQUIRE(false, "propagate_fwd: K=%d > 512 not supported", layer->K); }
    MK_CHECK_CUDA(cudaGetLastError());
    return 0;
}
