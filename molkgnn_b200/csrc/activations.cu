// Row norms / zero padding of the layer input and the fused neighbour sum that ends every MolGCN layer.
//   k_pad_norm       x -> zero-padded copy + ||row||  (norm half of torch's cosine_similarity, kernels.py:189-190)
//   k_propagate_fwd  h[i] = sum_{(j->i)} sim_sc[j]     (MessagePassing aggr='add', KernelLayer.py:119-123), reading
//                    only the L_deg(j) non-zero columns of each source row from the compact score blocks, summing in
//                    edge order (deterministic, no atomics), and emitting ||h_i|| for the next layer's cosines.
// Both are pure HBM-bound streaming kernels: one warp per row, float4 where the layout allows.
#include "common.cuh"
#include "tc.cuh"
#include "tile.cuh"

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

// Tile-ordered neighbour sum (plans with molecule tiles, K <= 112): one CTA per tile.  The compact score rows of the tile's
// nodes are expanded once into a dense [node][column] block in shared memory (coalesced reads, one thread per
// (node, kernel) pair); every in-neighbour of a tile node is a tile node (TileMetaG::inl), so h[v] is then a sum of <= 4
// shared-memory rows in edge order -- the same fp32 sums, element by element, as k_propagate_fwd (bitwise equal
// results), an eighth of the instructions of a warp-per-node gather from global memory.  Writes h, ||h|| and (optionally) the next layer's fp16 (hi, lo) tile images.
struct PropTileArgs {
    const TileMetaG* meta;
    int K, ldh;
    int L[4], koff[4];
    long long scoff[4];
    const float* sc;
    float* h; float* hnorm;
    unsigned char* ximg; int Fk, x_one;
};

constexpr int PT_THREADS = 512;

__global__ void __launch_bounds__(PT_THREADS) k_propagate_tile(const __grid_constant__ PropTileArgs a) {
    extern __shared__ __align__(16) unsigned char smem_p[];
    TileMetaG& m = *reinterpret_cast<TileMetaG*>(smem_p);
    float* scS = reinterpret_cast<float*>(smem_p + (sizeof(TileMetaG) + 15) / 16 * 16);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.meta + tile);
        uint4* dst = reinterpret_cast<uint4*>(smem_p);
        for (int i = tid; i < (int)(sizeof(TileMetaG) / 16); i += PT_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const int nn = m.nn, t0 = m.t0, ld = a.ldh;
    for (int i = tid; i < nn * (ld >> 2); i += PT_THREADS) reinterpret_cast<float4*>(scS)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
#pragma unroll
    for (int d = 1; d <= 4; ++d) {
        const int L = a.L[d - 1];
        if (L == 0) continue;
        const int np = m.cnt[d - 1] * L;
        const float rL = 1.0f / (float)L;
        const float* base = a.sc + a.scoff[d - 1];
        const int ko = a.koff[d - 1];
#pragma unroll 4
        for (int p = tid; p < np; p += PT_THREADS) {
            const int i = (int)(((float)p + 0.5f) * rL);
            const int k = p - i * L;
            const int nl_ = m.list[d - 1][i];
            scS[nl_ * ld + ko + k] = __ldg(base + (size_t)m.posl[nl_] * L + k);
        }
    }
    __syncthreads();
    const int c0 = 4 * lane;
    unsigned char* Xhi = a.ximg ? a.ximg + (size_t)tile * 2 * a.x_one : nullptr;
    unsigned char* Xlo = a.ximg ? Xhi + a.x_one : nullptr;
    for (int v = warp; v < nn; v += PT_THREADS / 32) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const int cnt = min((int)m.incnt[v], 4);
        const uint32_t w = m.inl[v];
        if (c0 < ld) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {        // edge order
                if (t < cnt) {
                    const float4 s4 = *reinterpret_cast<const float4*>(scS + (int)((w >> (8 * t)) & 0xffu) * ld + c0);
                    acc[0] += s4.x; acc[1] += s4.y; acc[2] += s4.z; acc[3] += s4.w;
                }
            }
        }
        float ss = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) { if (c0 + u >= a.K) acc[u] = 0.f; ss += acc[u] * acc[u]; }
        ss = warp_sum(ss);
        const float nrm = sqrtf(ss);
        const int i = t0 + v;
        if (c0 < ld) st4(a.h + (size_t)i * ld + c0, make_float4(acc[0], acc[1], acc[2], acc[3]));
        if (lane == 0 && a.hnorm) a.hnorm[i] = nrm;
        if (Xhi && c0 < a.Fk) {
            const float rinv = 1.0f / fmaxf(nrm, MOLKGNN_COS_EPS);
            __align__(8) __half2 hi[2];
            __align__(8) __half2 lo[2];
            tc::split_u2(acc[0] * rinv, acc[1] * rinv, hi[0], lo[0]);
            tc::split_u2(acc[2] * rinv, acc[3] * rinv, hi[1], lo[1]);
            const uint32_t off = tc::il_off(v, c0, a.Fk);
            *reinterpret_cast<uint2*>(Xhi + off) = *reinterpret_cast<const uint2*>(hi);
            *reinterpret_cast<uint2*>(Xlo + off) = *reinterpret_cast<const uint2*>(lo);
        }
    }
    // pad rows up to the next multiple of 16: the extent the tile kernels' MMAs read
    if (Xhi && c0 < a.Fk) {
        const int rend = min(TNODES, (nn + 15) & ~15);
        for (int rr = nn + warp; rr < rend; rr += PT_THREADS / 32) {
            const uint32_t off = tc::il_off(rr, c0, a.Fk);
            *reinterpret_cast<uint2*>(Xhi + off) = make_uint2(0u, 0u);
            *reinterpret_cast<uint2*>(Xlo + off) = make_uint2(0u, 0u);
        }
    }
}

}  // namespace mk

using namespace mk;

extern "C" int molkgnn_pad_norm(const float* x, int32_t N, int32_t F, int32_t ldx, float* out, int32_t ldo, float* norm,
                                void* stream_) {
    MK_REQUIRE(ldo >= F, "pad_norm: ldo=%d < F=%d", ldo, F);
    if (N == 0) return 0;
    count_launches(1);
    ProfScope prof("pad_norm", (cudaStream_t)stream_);
    k_pad_norm<<<(N + 7) / 8, 256, 0, (cudaStream_t)stream_>>>(x, N, F, ldx, out, ldo, norm);
    MK_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int molkgnn_propagate_fwd(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* sc,
                                     const int64_t scoff[4], float* h, int32_t ldh, float* hnorm, void* ximg,
                                     void* stream_) {
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
    ProfScope prof("propagate_fwd", st);
    // tile-ordered kernel: tiled plan, K <= 112, 16-byte rows (always the case on the molecule-tile path)
    if (plan->n_tiles > 0 && plan->tile_meta && plan->tile_max_nodes <= TNODES && ldh % 4 == 0 && ldh <= 112 &&
        (reinterpret_cast<uintptr_t>(h) & 15) == 0 && (!ximg || (hnorm && (reinterpret_cast<uintptr_t>(ximg) & 127) == 0))) {
        PropTileArgs t;
        t.meta = reinterpret_cast<const TileMetaG*>(plan->tile_meta);
        t.K = layer->K; t.ldh = ldh;
        for (int d = 0; d < 4; ++d) { t.L[d] = layer->L[d]; t.koff[d] = layer->koff[d]; t.scoff[d] = scoff[d]; }
        t.sc = sc; t.h = h; t.hnorm = hnorm;
        t.ximg = reinterpret_cast<unsigned char*>(ximg);
        t.Fk = tile_fk(ldh); t.x_one = tile_img_one(t.Fk);
        const int smem = (int)((sizeof(TileMetaG) + 15) / 16 * 16) + TNODES * ldh * 4;
        static int s_attr_dev[16] = {0};
        int& s_attr = s_attr_dev[device_index()];          // function attributes are per device
        if (smem > s_attr) {
            MK_CHECK_CUDA(cudaFuncSetAttribute(k_propagate_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            s_attr = smem;
        }
        k_propagate_tile<<<plan->n_tiles, PT_THREADS, smem, st>>>(t);
        MK_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    MK_REQUIRE(!ximg, "propagate_fwd: fused images need a tiled plan, hnorm, 16-byte aligned h, 128-byte aligned ximg and "
                      "ldh %% 4 == 0, ldh <= 112 (got ldh=%d)", ldh);
    if (ldh <= 32 * 1) k_propagate_fwd<1><<<grid, 256, 0, st>>>(a);
    else if (ldh <= 32 * 4) k_propagate_fwd<4><<<grid, 256, 0, st>>>(a);
    else if (ldh <= 32 * 8) k_propagate_fwd<8><<<grid, 256, 0, st>>>(a);
    else if (ldh <= 32 * 16) k_propagate_fwd<16><<<grid, 256, 0, st>>>(a);
    else { MK_REQUIRE(false, "propagate_fwd: K=%d > 512 not supported", layer->K); }
    MK_CHECK_CUDA(cudaGetLastError());
    return 0;
}
