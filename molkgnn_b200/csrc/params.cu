// Kernel-parameter side of the cosine similarities.
//   k_param_pack      normalise kernel rows once per step (the kernel half of torch's cosine_similarity,
//                     reference kernels.py:189-190), softmax mixing weights (kernels.py:402-412), chirality sign of
//                     every permuted support (kernels.py:338-341) -> packed, smem-ready layout
//   k_param_finalize  deterministic reduction of the per-CTA partial sums written by the backward kernel, chain
//                     rule through the normalisation and the softmax weights, results in the reference's layouts
#include <algorithm>
#include <stdarg.h>
#include <string.h>
#include "common.cuh"
#include "tc.cuh"
#include "tile.cuh"

namespace mk {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static long long g_launches = 0;
void count_launches(int n) { __atomic_fetch_add(&g_launches, (long long)n, __ATOMIC_RELAXED); }
long long launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

// ---- optional CUDA-event profiler: named scopes around the launches of the host entry points --------------------------
// Off by default (one relaxed load per scope).  bench.py switches it on around its timed region to get the live
// per-kernel launch durations on the launching stream (molkgnn_profile_enable / molkgnn_profile_read).
struct ProfRec { const char* name; cudaEvent_t e0, e1; };
static int g_prof_on = 0;
static char g_prof_only[64] = "";        // record only scopes of this name ("" = all)
static ProfRec g_prof[4096];
static int g_prof_n = 0;
static cudaEvent_t g_prof_pool[8192];
static int g_prof_pool_n = 0, g_prof_pool_used = 0;
static cudaEvent_t prof_event() {
    if (g_prof_pool_used == g_prof_pool_n) {
        if (g_prof_pool_n == 8192) return nullptr;
        if (cudaEventCreate(&g_prof_pool[g_prof_pool_n]) != cudaSuccess) return nullptr;
        ++g_prof_pool_n;
    }
    return g_prof_pool[g_prof_pool_used++];
}
ProfScope::ProfScope(const char* name, cudaStream_t st) : rec_(-1), st_(st) {
    if (!__atomic_load_n(&g_prof_on, __ATOMIC_RELAXED) || g_prof_n >= 4096) return;
    if (g_prof_only[0] && strcmp(g_prof_only, name)) return;
    cudaEvent_t a = prof_event(), b = prof_event();
    if (!a || !b) return;
    rec_ = g_prof_n++;
    g_prof[rec_] = ProfRec{name, a, b};
    cudaEventRecord(a, st_);
}
ProfScope::~ProfScope() {
    if (rec_ >= 0) cudaEventRecord(g_prof[rec_].e1, st_);
}

int device_num_sms() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}
int device_index() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return 0;
    return dev < 16 ? dev : 15;
}
int device_max_smem_optin() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return -1;
    return n;
}

struct PackArgs {
    int F, Fp, Fe;
    int L[4];
    int row_begin[5];  // block index ranges per degree: (2d+1)*L rows each, then 4 tail blocks
    const float* x_center[4];
    const float* x_support[4];
    const float* edge_attr_support[4];
    const float* p_support[4];
    const float* w_support[4];
    const float* w_center[4];
    const float* w_edge[4];
    float* packed[4];
};

__device__ __forceinline__ float block_sum_128(float v, float* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = red[0] + red[1] + red[2] + red[3];
    __syncthreads();
    return t;
}

template <int D>
__device__ void sup_signs(const float* __restrict__ ps, int L, int8_t* out) {
    // sign of b2 . (b0 x b1) with b_j = p_support[k, perm[j]] (kernels.py:331-341), for every kernel and permutation
    for (int i = threadIdx.x; i < L * Perm<D>::P; i += blockDim.x) {
        int k = i / Perm<D>::P, pi = i % Perm<D>::P;
        float b[3][3];
        for (int j = 0; j < 3; ++j) {
            int s = 0;
#pragma unroll
            for (int q = 0; q < Perm<D>::P; ++q)
                if (q == pi) s = Perm<D>::at(q, j < D ? j : 0);
            for (int c = 0; c < 3; ++c) b[j][c] = ps[((size_t)k * D + s) * 3 + c];
        }
        float cx = __fsub_rn(__fmul_rn(b[0][1], b[1][2]), __fmul_rn(b[0][2], b[1][1]));
        float cy = __fsub_rn(__fmul_rn(b[0][2], b[1][0]), __fmul_rn(b[0][0], b[1][2]));
        float cz = __fsub_rn(__fmul_rn(b[0][0], b[1][1]), __fmul_rn(b[0][1], b[1][0]));
        float dt = __fadd_rn(__fadd_rn(__fmul_rn(b[2][0], cx), __fmul_rn(b[2][1], cy)), __fmul_rn(b[2][2], cz));
        out[i] = dt > 0.f ? 1 : (dt < 0.f ? -1 : 0);
    }
}

// the pack kernels take up to PACK_NL layers per launch (blockIdx.y = layer): a step packs its whole stack in three launches
constexpr int PACK_NL = 4;
struct PackArgsN { PackArgs a[PACK_NL]; };

__global__ void __launch_bounds__(128) k_param_pack(const __grid_constant__ PackArgsN an) {
    __shared__ float red[4];
    const PackArgs& a = an.a[blockIdx.y];
    int b = blockIdx.x;
    if (b >= a.row_begin[4] + 4) return;
    if (b >= a.row_begin[4]) {  // tail blocks: weights + support signs of degree d
        int d = b - a.row_begin[4] + 1;
        int L = a.L[d - 1];
        if (L == 0) return;
        PackedLayout pl(d, L, a.Fp);
        float* pk = a.packed[d - 1];
        if (threadIdx.x == 0) {
            float es = expf(*a.w_support[d - 1]), ec = expf(*a.w_center[d - 1]), ee = expf(*a.w_edge[d - 1]);
            float den = (es + ec) + ee;
            float ws = es / den, wc = ec / den, we = ee / den;
            float W = (ws + wc) + we;
            pk[pl.w + 0] = ws; pk[pl.w + 1] = wc; pk[pl.w + 2] = we; pk[pl.w + 3] = W;
            pk[pl.w + 4] = 0.f; pk[pl.w + 5] = 0.f; pk[pl.w + 6] = 0.f; pk[pl.w + 7] = 0.f;
        }
        if (d == 4) sup_signs<4>(a.p_support[3], L, reinterpret_cast<int8_t*>(pk + pl.sign));
        return;
    }
    int d = 1;
    while (b >= a.row_begin[d]) ++d;
    int L = a.L[d - 1];
    int row = b - a.row_begin[d - 1];  // 0 .. (2d+1)L-1 : support rows (s*L+k), centre rows, edge rows (s*L+k)
    PackedLayout pl(d, L, a.Fp);
    float* pk = a.packed[d - 1];
    const float* src;
    float* dst;
    float* nrm_out;
    int n, npad;
    if (row < pl.rows_x) {
        int s = row / L, k = row % L;
        src = s < d ? a.x_support[d - 1] + ((size_t)k * d + s) * a.F : a.x_center[d - 1] + (size_t)k * a.F;
        dst = pk + pl.sup + (size_t)row * a.Fp;
        nrm_out = pk + pl.norm + row;
        n = a.F; npad = a.Fp;
    } else {
        int er = row - pl.rows_x;
        int s = er / L, k = er % L;
        src = a.edge_attr_support[d - 1] + ((size_t)k * d + s) * a.Fe;
        dst = pk + pl.es + (size_t)er * EP;
        nrm_out = pk + pl.enorm + er;
        n = a.Fe; npad = EP;
    }
    float ss = 0.f;
    for (int i = threadIdx.x; i < n; i += 128) { float v = src[i]; ss += v * v; }
    ss = block_sum_128(ss, red);
    float nrm = sqrtf(ss);
    float den = fmaxf(nrm, MOLKGNN_COS_EPS);
    for (int i = threadIdx.x; i < npad; i += 128) dst[i] = i < n ? src[i] / den : 0.f;
    if (threadIdx.x == 0) *nrm_out = nrm;
}

// fp16 (hi, lo) images of the normalised kernel rows in the interleaved UMMA operand layout (tc.cuh): support image rows
// k*DS + s (DS = 4 for d = 3, else d; unused rows zero), centre image rows k.  One thread per (row, 8-column chunk).
struct PackTcArgs {
    int Fp;
    int L[4];
    int item_begin[5];
    float* packed[4];
};

struct PackTcArgsN { PackTcArgs a[PACK_NL]; };

__global__ void __launch_bounds__(256) k_param_pack_tc(const __grid_constant__ PackTcArgsN an) {
    const PackTcArgs& a = an.a[blockIdx.y];
    const int it = blockIdx.x * 256 + threadIdx.x;
    if (it >= a.item_begin[4]) return;
    int d = 1;
    while (it >= a.item_begin[d]) ++d;
    const int L = a.L[d - 1];
    const PackedLayout pl(d, L, a.Fp);
    const int nch = pl.Fk / 8;
    const int local = it - a.item_begin[d - 1];
    int row = local / nch;
    const int c = local % nch;
    float* pk = a.packed[d - 1];
    unsigned char* img = reinterpret_cast<unsigned char*>(pk + pl.tc);
    const float* src = nullptr;
    int64_t hi_off, lo_off;
    if (row < pl.KSpad) {
        const int k = row / pl.DS, s = row % pl.DS;
        if (k < L && s < d) src = pk + pl.sup + ((size_t)s * L + k) * a.Fp;
        hi_off = pl.tc_sup_hi(); lo_off = pl.tc_sup_lo();
    } else {
        row -= pl.KSpad;
        if (row < L) src = pk + pl.sup + ((size_t)d * L + row) * a.Fp;
        hi_off = pl.tc_cen_hi(); lo_off = pl.tc_cen_lo();
    }
    __align__(16) __half hi[8];
    __align__(16) __half lo[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int col = 8 * c + t;
        const float v = (src && col < a.Fp) ? src[col] : 0.f;
        tc::split_h(v, hi[t], lo[t]);
    }
    const uint32_t off = tc::il_off(row, 8 * c, pl.Fk);
    *reinterpret_cast<uint4*>(img + hi_off + off) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(img + lo_off + off) = *reinterpret_cast<const uint4*>(lo);
}

// fp16 (hi, lo) images of the kernel blocks for the tile kernels (tile.cuh): block b at img + b * 2 * tile_img_one(Fk),
// hi image then lo image, rows in TileBlocks order, unscaled split.  One thread per (block, row, 8-column chunk).
struct PackTileArgs {
    int Fp, Fk;
    int L[4];
    TileBlocks tb;
    const float* packed[4];
    unsigned char* img;
};

struct PackTileArgsN { PackTileArgs a[PACK_NL]; };

__global__ void __launch_bounds__(256) k_param_pack_tile(const __grid_constant__ PackTileArgsN an) {
    const PackTileArgs& a = an.a[blockIdx.y];
    if (!a.img) return;
    const int nch = a.Fk / 8;
    const int it = blockIdx.x * 256 + threadIdx.x;
    if (it >= a.tb.nb * 128 * nch) {
        // bond-support table of the layer-fused forward: block by block, segment by segment, [slot][half][kernel] float4
        int e = it - a.tb.nb * 128 * nch;
        if (e >= tile_es_f4(a.L)) return;
        float4* es = reinterpret_cast<float4*>(a.img + tile_es_off(a.tb.nb, a.Fk));
        const int e_out = e;
        for (int blk = 0; blk < a.tb.nb; ++blk)
            for (int sgi = 0; sgi < a.tb.nseg[blk]; ++sgi) {
                const TileSeg sg = a.tb.seg[blk][sgi];
                const int cnt = sg.d * 2 * sg.nk;
                if (e >= 0 && e < cnt) {
                    const int kl = e % sg.nk, sh = e / sg.nk;               // sh = slot * 2 + half
                    const int L = a.L[sg.d - 1];
                    const PackedLayout pl(sg.d, L, a.Fp);
                    es[e_out] = *reinterpret_cast<const float4*>(a.packed[sg.d - 1] + pl.es +
                                                                 (size_t)((sh >> 1) * L + sg.k0 + kl) * EP + (sh & 1) * 4);
                    return;
                }
                e -= cnt;
            }
        return;
    }
    const int c = it % nch;
    const int row = (it / nch) % 128;
    const int blk = it / (nch * 128);
    const float* src = nullptr;
    for (int sgi = 0; sgi < a.tb.nseg[blk]; ++sgi) {
        const TileSeg sg = a.tb.seg[blk][sgi];
        const int r = row - sg.rowbase;
        if (r >= 0 && r < sg.nk * (sg.d + 1)) {
            const int slot = r / sg.nk, k = sg.k0 + r % sg.nk;
            const int L = a.L[sg.d - 1];
            const PackedLayout pl(sg.d, L, a.Fp);
            src = a.packed[sg.d - 1] + pl.sup + ((size_t)slot * L + k) * a.Fp;   // packed rows: s*L+k supports, d*L+k centre
        }
    }
    __align__(16) __half2 hi[4];
    __align__(16) __half2 lo[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int col = 8 * c + 2 * t;
        const float v0 = (src && col < a.Fp) ? src[col] : 0.f;
        const float v1 = (src && col + 1 < a.Fp) ? src[col + 1] : 0.f;
        tc::split_u2(v0, v1, hi[t], lo[t]);
    }
    const int one = tile_img_one(a.Fk);
    unsigned char* base = a.img + (size_t)blk * 2 * one;
    const uint32_t off = tc::il_off(row, 8 * c, a.Fk);
    *reinterpret_cast<uint4*>(base + off) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(base + one + off) = *reinterpret_cast<const uint4*>(lo);
    // K-step-major copy (tile.cuh): K step ks = c / 2 of block blk is one contiguous 8 KB stage [hi | lo]
    unsigned char* ks = a.img + tile_img_ks_off(a.tb.nb, a.Fk) + ((size_t)blk * (a.Fk >> 4) + (c >> 1)) * TILE_KS_BYTES +
                        (size_t)(row >> 3) * 256 + (size_t)(c & 1) * 128 + (size_t)(row & 7) * 16;
    *reinterpret_cast<uint4*>(ks) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(ks + TILE_KS_BYTES / 2) = *reinterpret_cast<const uint4*>(lo);
}

// ---------------------------------------------------------------------------------------------------------------
struct FinArgs {
    int F, Fp, Fe, FW;       // FW = Fp + EP : row width of the partial accumulators
    int L[4];
    int row_begin[5];        // blocks per degree: (d+1)*L rows
    int ncta[4];             // number of partial copies per degree
    int64_t part_off[4];     // float offset of degree d's partial block; copy c at + c*rows_x*FW
    const float* partials;
    const float* packed[4];
    float* gx_center[4];
    float* gx_support[4];
    float* ge_support[4];
    float* q;                // [sum rows] per-row Q contributions: q[row_begin[d]+row] (x part), and edge part after
    int q_edge_off;          // offset of edge Q values
    int prescaled;           // node-attribute partial sums already carry the mixing factor (tile backward)
};

// one block (128 threads) per packed row; writes final dL/d(raw row) and per-row Q = sum(hat * G)
__global__ void __launch_bounds__(128) k_param_finalize(FinArgs a) {
    __shared__ float red[4];
    int b = blockIdx.x;
    int d = 1;
    while (b >= a.row_begin[d]) ++d;
    int L = a.L[d - 1];
    int row = b - a.row_begin[d - 1];
    PackedLayout pl(d, L, a.Fp);
    const float* pk = a.packed[d - 1];
    const int s = row / L, k = row % L;
    const bool is_center = (s == d);
    const float ws = pk[pl.w + 0], wc = pk[pl.w + 1], we = pk[pl.w + 2], W = pk[pl.w + 3];
    const size_t stride = (size_t)pl.rows_x * a.FW;
    const float* part = a.partials + a.part_off[d - 1] + (size_t)row * a.FW;
    const int nc = a.ncta[d - 1];
    // ---- node-attribute part ----
    {
        const float* hat = pk + pl.sup + (size_t)row * a.Fp;
        float dot = 0.f;
        float g[4];  // up to 512 features per thread-strided loop: keep values in registers for the second pass
        int cnt = 0;
        for (int f = threadIdx.x; f < a.Fp; f += 128, ++cnt) {
            float acc = 0.f;
            for (int c = 0; c < nc; ++c) acc += part[c * stride + f];
            if (cnt < 4) g[cnt] = acc;
            dot += acc * hat[f];
        }
        dot = block_sum_128(dot, red);
        const float nrm = pk[pl.norm + row];
        const float coef = (is_center ? wc : ws / (float)d) / W;
        if (a.prescaled) {   // the tile backward accumulates coef * G: recover G (the chain rule below is written for it)
            const float rc = 1.0f / coef;
#pragma unroll
            for (int c = 0; c < 4; ++c) g[c] *= rc;
            dot *= rc;
        }
        float* out = is_center ? (a.gx_center[d - 1] ? a.gx_center[d - 1] + (size_t)k * a.F : nullptr)
                               : (a.gx_support[d - 1] ? a.gx_support[d - 1] + ((size_t)k * d + s) * a.F : nullptr);
        cnt = 0;
        for (int f = threadIdx.x; f < a.Fp; f += 128, ++cnt) {
            float G;
            if (cnt < 4) G = g[cnt];
            else {
                G = 0.f;
                for (int c = 0; c < nc; ++c) G += part[c * stride + f];
                if (a.prescaled) G *= 1.0f / coef;
            }
            float v = nrm > MOLKGNN_COS_EPS ? (G - dot * hat[f]) / nrm : G / MOLKGNN_COS_EPS;
            if (out && f < a.F) out[f] = coef * v;
        }
        if (threadIdx.x == 0) a.q[b] = is_center ? dot : dot / (float)d;
    }
    // ---- bond-attribute part (support rows only) ----
    if (!is_center) {
        const float* hat = pk + pl.es + (size_t)row * EP;
        float G = 0.f, h = 0.f;
        if (threadIdx.x < EP) {
            for (int c = 0; c < nc; ++c) G += part[c * stride + a.Fp + threadIdx.x];
            h = hat[threadIdx.x];
        }
        float dot = block_sum_128(G * h, red);
        const float nrm = pk[pl.enorm + row];
        if (threadIdx.x < a.Fe && a.ge_support[d - 1]) {
            float v = nrm > MOLKGNN_COS_EPS ? (G - dot * h) / nrm : G / MOLKGNN_COS_EPS;
            a.ge_support[d - 1][((size_t)k * d + s) * a.Fe + threadIdx.x] = (we / (float)d / W) * v;
        }
        if (threadIdx.x == 0) a.q[a.q_edge_off + b] = dot / (float)d;
    }
}

struct ThetaArgs {
    int L[4];
    int row_begin[5];
    int Fp;
    int q_edge_off;
    const float* q;
    const float* packed[4];
    float* gw_support[4];
    float* gw_center[4];
    float* gw_edge[4];
};

// d(theta_i) = (w_i/W) * (Q_i - sum_j (w_j/W) Q_j),  Q_i = sum_{n,k} g*chi*T_i   (T = S*, C, E);  one block per degree
__global__ void __launch_bounds__(128) k_theta(ThetaArgs a) {
    __shared__ float red[4];
    int d = blockIdx.x + 1;
    int L = a.L[d - 1];
    if (L == 0) return;
    PackedLayout pl(d, L, a.Fp);
    const float* pk = a.packed[d - 1];
    float qs = 0.f, qc = 0.f, qe = 0.f;
    int base = a.row_begin[d - 1];
    for (int r = threadIdx.x; r < pl.rows_x; r += 128) {
        if (r < d * L) { qs += a.q[base + r]; qe += a.q[a.q_edge_off + base + r]; }
        else qc += a.q[base + r];
    }
    qs = block_sum_128(qs, red);
    qc = block_sum_128(qc, red);
    qe = block_sum_128(qe, red);
    if (threadIdx.x == 0) {
        const float ws = pk[pl.w + 0], wc = pk[pl.w + 1], we = pk[pl.w + 2], W = pk[pl.w + 3];
        float mean = (ws * qs + wc * qc + we * qe) / W;
        if (a.gw_support[d - 1]) *a.gw_support[d - 1] = ws / W * (qs - mean);
        if (a.gw_center[d - 1]) *a.gw_center[d - 1] = wc / W * (qc - mean);
        if (a.gw_edge[d - 1]) *a.gw_edge[d - 1] = we / W * (qe - mean);
    }
}

// entry used by conv_bwd.cu
int launch_param_finalize(const molkgnn_layer_t* layer, const float* partials, const int64_t part_off[4],
                          const int ncta[4], float* q_scratch, const molkgnn_layer_grads_t* grads, int prescaled,
                          cudaStream_t st) {
    FinArgs fa;
    fa.prescaled = prescaled;
    ThetaArgs ta;
    fa.F = layer->F; fa.Fp = layer->Fp; fa.Fe = layer->Fe; fa.FW = layer->Fp + EP;
    int rb = 0;
    for (int d = 0; d < 4; ++d) {
        fa.L[d] = ta.L[d] = layer->L[d];
        fa.row_begin[d] = ta.row_begin[d] = rb;
        rb += (d + 2) * layer->L[d];
        fa.ncta[d] = ncta[d];
        fa.part_off[d] = part_off[d];
        fa.packed[d] = ta.packed[d] = layer->packed[d];
        fa.gx_center[d] = grads->x_center[d];
        fa.gx_support[d] = grads->x_support[d];
        fa.ge_support[d] = grads->edge_attr_support[d];
        ta.gw_support[d] = grads->w_support[d];
        ta.gw_center[d] = grads->w_center[d];
        ta.gw_edge[d] = grads->w_edge[d];
    }
    fa.row_begin[4] = ta.row_begin[4] = rb;
    fa.partials = partials;
    fa.q = q_scratch;
    fa.q_edge_off = ta.q_edge_off = rb;
    ta.q = q_scratch;
    ta.Fp = layer->Fp;
    if (rb == 0) return 0;
    ProfScope prof("param_finalize", st);
    k_param_finalize<<<rb, 128, 0, st>>>(fa);
    k_theta<<<4, 128, 0, st>>>(ta);
    count_launches(2);
    MK_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mk

using namespace mk;

extern "C" const char* molkgnn_last_error(void) { return mk::g_err; }
extern "C" int molkgnn_version(void) { return 100; }
extern "C" int molkgnn_num_sms(void) { return device_num_sms(); }
namespace mk { long long launches(); }
extern "C" int64_t molkgnn_launch_count(void) { return mk::launches(); }

extern "C" int molkgnn_profile_only(const char* name) {
    if (!name) { mk::g_prof_only[0] = 0; return 0; }
    strncpy(mk::g_prof_only, name, sizeof(mk::g_prof_only) - 1);
    mk::g_prof_only[sizeof(mk::g_prof_only) - 1] = 0;
    return 0;
}
extern "C" int molkgnn_profile_enable(int on) {
    const int old = mk::g_prof_on;
    if (on) { mk::g_prof_n = 0; mk::g_prof_pool_used = 0; }
    __atomic_store_n(&mk::g_prof_on, on ? 1 : 0, __ATOMIC_RELAXED);
    return old;
}
// "name launches total_ms\n" per scope name, in first-seen order; synchronises the device.  Returns the text length.
extern "C" int molkgnn_profile_read(char* buf, int cap) {
    if (cudaDeviceSynchronize() != cudaSuccess) { mk::set_error("profile_read: device error"); return -2; }
    const char* names[64]; int cnt[64]; double ms[64]; int nn = 0;
    for (int i = 0; i < mk::g_prof_n; ++i) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, mk::g_prof[i].e0, mk::g_prof[i].e1) != cudaSuccess) continue;
        int j = 0;
        while (j < nn && strcmp(names[j], mk::g_prof[i].name)) ++j;
        if (j == nn) { if (nn == 64) continue; names[nn] = mk::g_prof[i].name; cnt[nn] = 0; ms[nn] = 0.0; ++nn; }
        ++cnt[j]; ms[j] += t;
    }
    int len = 0;
    for (int j = 0; j < nn && len < cap - 96; ++j) len += snprintf(buf + len, cap - len, "%s %d %.6f\n", names[j], cnt[j], ms[j]);
    if (cap > 0) buf[len < cap ? len : cap - 1] = 0;
    mk::g_prof_n = 0; mk::g_prof_pool_used = 0;
    return len;
}

extern "C" int64_t molkgnn_packed_floats(int32_t d, int32_t L, int32_t Fp) {
    if (d < 1 || d > 4 || L < 0 || Fp < 4 || Fp % 4) return -1;
    return PackedLayout(d, L, Fp).total;
}

namespace mk {
// tile kernels are eligible when the kernel rows fit TILE_MAXB blocks and one block + a node tile fit shared memory
bool tile_layer_ok(const molkgnn_layer_t* layer) {
    TileBlocks tb;
    return layer->K > 0 && tile_fk(layer->Fp) <= 112 && tb.build(layer->L);
}
}  // namespace mk

namespace mk {
bool wide_layer_ok(const molkgnn_layer_t* layer);
int64_t wide_img_bytes(const molkgnn_layer_t* layer);
int launch_param_pack_wide(const molkgnn_layer_t* layer, cudaStream_t st);
}

extern "C" int64_t molkgnn_tile_img_bytes(const molkgnn_layer_t* layer) {
    TileBlocks tb;
    if (wide_layer_ok(layer)) return wide_img_bytes(layer);      // wide layers: stage-major images (conv_fwd_wide.cu)
    if (!tile_layer_ok(layer) || !tb.build(layer->L)) return 0;
    // block-major images + K-step-major copy + bond-support table (tile.cuh)
    return tile_es_off(tb.nb, tile_fk(layer->Fp)) + ((int64_t)tile_es_f4(layer->L) * 16 + 127) / 128 * 128;
}

// what: bit 0 = normalised rows / weights / signs, bit 1 = bucket-order tensor-core images, bit 2 = tile images
extern "C" int molkgnn_param_pack_layers(const molkgnn_layer_t* layers, int32_t nl, int32_t what, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    MK_REQUIRE(nl >= 1 && layers, "param_pack: no layers");
    for (int l0 = 0; l0 < nl; l0 += PACK_NL) {
        const int n = std::min(PACK_NL, nl - l0);
        PackArgsN pa;
        PackTcArgsN pt;
        PackTileArgsN pi;
        memset(&pa, 0, sizeof(pa)); memset(&pt, 0, sizeof(pt)); memset(&pi, 0, sizeof(pi));
        int rb_max = 0, ib_max = 0, items_max = 0;
        for (int i = 0; i < n; ++i) {
            const molkgnn_layer_t* layer = layers + l0 + i;
            MK_REQUIRE(layer->Fp % 4 == 0 && layer->Fp >= layer->F, "param_pack: Fp=%d must be a multiple of 4 >= F=%d",
                       layer->Fp, layer->F);
            MK_REQUIRE(layer->Fe >= 1 && layer->Fe <= EP, "param_pack: edge_attr_dim %d not in 1..%d", layer->Fe, EP);
            PackArgs& a = pa.a[i];
            a.F = layer->F; a.Fp = layer->Fp; a.Fe = layer->Fe;
            int rb = 0;
            for (int d = 0; d < 4; ++d) {
                a.L[d] = layer->L[d];
                a.row_begin[d] = rb;
                rb += (2 * (d + 1) + 1) * layer->L[d];
                a.x_center[d] = layer->x_center[d];
                a.x_support[d] = layer->x_support[d];
                a.edge_attr_support[d] = layer->edge_attr_support[d];
                a.p_support[d] = layer->p_support[d];
                a.w_support[d] = layer->w_support[d];
                a.w_center[d] = layer->w_center[d];
                a.w_edge[d] = layer->w_edge[d];
                a.packed[d] = layer->packed[d];
            }
            a.row_begin[4] = rb;
            rb_max = std::max(rb_max, rb + 4);
            PackTcArgs& t = pt.a[i];
            t.Fp = layer->Fp;
            int ib = 0;
            for (int d = 0; d < 4; ++d) {
                t.L[d] = layer->L[d];
                t.packed[d] = layer->packed[d];
                t.item_begin[d] = ib;
                if (layer->L[d] > 0) {
                    const PackedLayout pl(d + 1, layer->L[d], layer->Fp);
                    ib += (pl.KSpad + pl.Lpad) * (pl.Fk / 8);
                }
            }
            t.item_begin[4] = ib;
            ib_max = std::max(ib_max, ib);
            PackTileArgs& ta = pi.a[i];
            ta.img = nullptr;
            if (layer->tile_img && tile_layer_ok(layer)) {
                MK_REQUIRE((reinterpret_cast<uintptr_t>(layer->tile_img) & 127) == 0, "param_pack: tile_img must be 128-byte aligned");
                ta.Fp = layer->Fp; ta.Fk = tile_fk(layer->Fp);
                for (int d = 0; d < 4; ++d) { ta.L[d] = layer->L[d]; ta.packed[d] = layer->packed[d]; }
                ta.tb.build(layer->L);
                ta.img = reinterpret_cast<unsigned char*>(layer->tile_img);
                items_max = std::max(items_max, ta.tb.nb * 128 * (ta.Fk / 8) + tile_es_f4(layer->L));
            }
        }
        ProfScope prof("param_pack", st);
        if (what & 1) {
            k_param_pack<<<dim3(rb_max, n), 128, 0, st>>>(pa);
            count_launches(1);
            MK_CHECK_CUDA(cudaGetLastError());
        }
        if ((what & 2) && ib_max > 0) {
            k_param_pack_tc<<<dim3((ib_max + 255) / 256, n), 256, 0, st>>>(pt);
            count_launches(1);
            MK_CHECK_CUDA(cudaGetLastError());
        }
        if ((what & 4) && items_max > 0) {
            k_param_pack_tile<<<dim3((items_max + 255) / 256, n), 256, 0, st>>>(pi);
            count_launches(1);
            MK_CHECK_CUDA(cudaGetLastError());
        }
        if (what & 4)
            for (int i = 0; i < n; ++i) {
                const int rc = launch_param_pack_wide(layers + l0 + i, st);
                if (rc) return rc;
            }
    }
    return 0;
}

extern "C" int molkgnn_param_pack(const molkgnn_layer_t* layer, void* stream_) {
    return molkgnn_param_pack_layers(layer, 1, 7, stream_);
}
