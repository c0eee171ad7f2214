// Fused forward of the molecular-kernel convolution for all four degree buckets of one layer.
// Replaces KernelConv.calculate_total_score (reference kernels.py:353-425) and the bucket gathers / output
// assembly of BaseKernelSetConv.forward (kernels.py:519-548, 674-747).
//
// Structure (fp32 SIMT, one persistent CTA per SM, dynamic tile queue ordered by degree 4,3,2,1):
//   * the L2-normalised kernel set of the current degree (support rows, centre rows, bond rows) is staged in
//     shared memory once and stays resident while the CTA keeps pulling node tiles of that degree;
//   * a node tile gathers the d neighbour rows + the focal row of TN nodes (float4, coalesced along features),
//     dividing by max(||x||, eps) on the way in, so every cosine below is a plain dot product;
//   * a warp owns (32*RN nodes) x (RK kernels): lane <-> node, kernel rows are warp-broadcast shared loads,
//     the d x d dot-product tile of each (node, kernel) pair plus the centre dot live in registers;
//   * epilogue in registers: mean over j for every permutation of the compile-time table, first-max arg-max
//     (kernels.py:373), bond-attribute cosine at the arg-max permutation (kernels.py:382-390), chirality sign
//     (kernels.py:279-350) and the softmax-weighted mix (kernels.py:402-425).
// If the kernel set or a node tile does not fit in shared memory (wide layers) the feature dimension is processed
// in chunks with the accumulators kept in registers, and kernels in ranges.
#include <algorithm>
#include <stdlib.h>
#include "common.cuh"

namespace mk {

constexpr int FWD_THREADS = 256;   // 8 warps = 2 per scheduler: the register file allows 255 registers per thread
constexpr int FWD_WARPS = FWD_THREADS / 32;

template <int D> struct FwdTile;   // register tile: RN node slots x RK kernels per lane
// tuned so that the README kernel counts 10/20/30/50 give 8 (node set, kernel group) items per tile = one per warp
template <> struct FwdTile<1> { static constexpr int RN = 2, RK = 3; };
template <> struct FwdTile<2> { static constexpr int RN = 2, RK = 5; };
template <> struct FwdTile<3> { static constexpr int RN = 2, RK = 4; };
template <> struct FwdTile<4> { static constexpr int RN = 1, RK = 7; };

struct FwdCfg {
    int LKc;   // kernels per range
    int nkr;   // kernel ranges
    int sets;  // node sets (32*RN nodes) per tile
    int TN;    // nodes per tile
    int KGc;   // kernel groups per full range
    // byte offsets into dynamic shared memory (the layout is per degree; B rows start at 0)
    int sm_A, sm_E, sm_ES, sm_rownode, sm_rinv, sm_dup;
};

struct FwdArgs {
    const float* x; const float* xnorm; int ldx;
    const int* sel; const int* nei; const float* ehat; const int8_t* tsign;
    int n[4], boff[4], eoff[4], L[4], koff[4];
    const float* packed[4];
    int F, Fp, FC, nfc, fsa;
    FwdCfg cfg[4];
    int tile_begin[5];     // queue position q = 0..3 <-> degree 4-q
    int is_last;
    float* sc; int sc_mode; int ld_sc; long long scoff[4];
    uint8_t* argmax; uint8_t* argmax_free; const uint8_t* argmax_in;
    int* counter;
};

__device__ __forceinline__ void fma4(float& acc, const float4& a, const float4& b) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    acc = fmaf(a.w, b.w, acc);
}

template <int D>
__device__ __forceinline__ void fwd_tile(const FwdArgs& a, unsigned char* smem, int tl, int& staged_key) {
    constexpr int RN = FwdTile<D>::RN, RK = FwdTile<D>::RK, P = Perm<D>::P;
    const FwdCfg c = a.cfg[D - 1];
    const int L = a.L[D - 1], n = a.n[D - 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ntn = (n + c.TN - 1) / c.TN;
    const int kr = tl / ntn, tn = tl % ntn;
    const int node0 = tn * c.TN;
    const int k0r = kr * c.LKc;
    const int LK = min(c.LKc, L - k0r);
    const int KG = (LK + RK - 1) / RK;
    const int TN = c.TN;
    const PackedLayout pl(D, L, a.Fp);
    const float* __restrict__ pk = a.packed[D - 1];

    float* Bs = reinterpret_cast<float*>(smem);
    float* As = reinterpret_cast<float*>(smem + c.sm_A);
    float* Es = reinterpret_cast<float*>(smem + c.sm_E);
    float* ESs = reinterpret_cast<float*>(smem + c.sm_ES);
    int* rowNode = reinterpret_cast<int*>(smem + c.sm_rownode);
    float* rowRinv = reinterpret_cast<float*>(smem + c.sm_rinv);
    float* rowNrm = rowRinv + (D + 1) * c.TN;
    unsigned char* dupf = smem + c.sm_dup;

    const int eoff = a.eoff[D - 1], boff = a.boff[D - 1];
    // ---- tile set-up: row -> node table, neighbour bond rows, support bond rows ----
    for (int i = tid; i < (D + 1) * TN; i += FWD_THREADS) {
        int j = i / TN, nl = i % TN, r = node0 + nl;
        int node = -1;
        if (r < n) node = j < D ? a.nei[(size_t)eoff + (size_t)r * D + j] : a.sel[boff + r];
        rowNode[i] = node;
        // 1 / max(||x||, eps): the node half of the cosine denominators (kernels.py:189-190)
        const float nrm = node >= 0 ? fmaxf(a.xnorm[node], MOLKGNN_COS_EPS) : 1.0f;
        rowNrm[i] = nrm;
        rowRinv[i] = 1.0f / nrm;
    }
    {
        const int nvalid = min(TN, n - node0) * D * EP;
        const float* src = a.ehat + ((size_t)eoff + (size_t)node0 * D) * EP;
        for (int i = tid * 4; i < TN * D * EP; i += FWD_THREADS * 4)
            st4(Es + i, i < nvalid ? ld4(src + i) : make_float4(0.f, 0.f, 0.f, 0.f));
    }
    const int key = D * 65536 + kr;
    const bool restage_B = (a.nfc > 1) || (staged_key != key);
    if (staged_key != key) {
        for (int i = tid; i < D * LK * (EP / 4); i += FWD_THREADS) {
            int row = i / (EP / 4), q = i % (EP / 4);
            int s = row / LK, kk = row % LK;
            st4(ESs + (s * c.LKc + kk) * EP + 4 * q, ld4(pk + pl.es + ((size_t)s * L + k0r + kk) * EP + 4 * q));
        }
    }
    staged_key = key;
    __syncthreads();
    if (D == 4 && a.is_last) {
        // chirality gate: any two of the four neighbour feature rows bit-equal (torch.equal, kernels.py:310-317)
        for (int nl = warp; nl < TN; nl += FWD_WARPS) {
            int u[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) u[j] = rowNode[(j < D ? j : 0) * TN + nl];
            bool dup = false;
            if (u[0] >= 0) {
                unsigned neq = 0;  // bit per pair: rows differ somewhere
                for (int f = lane; f < a.F; f += 32) {
                    float v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = a.x[(size_t)u[j] * a.ldx + f];
                    int b = 0;
#pragma unroll
                    for (int p = 0; p < 4; ++p)
#pragma unroll
                        for (int q = p + 1; q < 4; ++q, ++b) if (!(v[p] == v[q])) neq |= 1u << b;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) neq |= __shfl_xor_sync(0xffffffffu, neq, o);
                dup = (neq != 0x3fu);
            }
            if (lane == 0) dupf[nl] = dup ? 1 : 0;
        }
    }

    const int nitems = c.sets * KG;
    const int FC = a.FC, FCq = FC / 4, fsa = a.fsa;
    const float ws = pk[pl.w + 0], wc = pk[pl.w + 1], we = pk[pl.w + 2], W = pk[pl.w + 3];
    const float rW = 1.0f / W;

    float acc[RN][RK][D][D];
    float accc[RN][RK];

    auto stage = [&](int fc) {
        const int f0 = fc * FC;
        // gather the neighbour / focal rows: a warp covers one row per quad-column pass (coalesced float4), UNR rows
        // in flight per warp so that the dependent global loads overlap instead of serialising
        constexpr int UNR = 8;
        const int nrows = (D + 1) * TN;
        for (int q = lane; q < FCq; q += 32) {
            const bool fin = f0 + 4 * q < a.Fp;
            for (int row0 = warp; row0 < nrows; row0 += FWD_WARPS * UNR) {
                float4 v[UNR];
                float ri[UNR], nr[UNR];
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int row = row0 + u * FWD_WARPS;
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    ri[u] = 0.f;
                    nr[u] = 1.f;
                    if (row < nrows) {
                        const int node = rowNode[row];
                        ri[u] = rowRinv[row];
                        nr[u] = rowNrm[row];
                        if (node >= 0 && fin) v[u] = ld4(a.x + (size_t)node * a.ldx + f0 + 4 * q);
                    }
                }
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int row = row0 + u * FWD_WARPS;
                    if (row < nrows) {
                        // x / max(||x||, eps) with the precomputed reciprocal: product, exact residual, correction
                        float4 o;
                        o.x = div_by(v[u].x, nr[u], ri[u]); o.y = div_by(v[u].y, nr[u], ri[u]);
                        o.z = div_by(v[u].z, nr[u], ri[u]); o.w = div_by(v[u].w, nr[u], ri[u]);
                        st4(As + (size_t)row * fsa + 4 * q, o);
                    }
                }
            }
        }
        if (restage_B) {
            for (int row = warp; row < (D + 1) * LK; row += FWD_WARPS) {
                int s = row / LK, kk = row % LK;
                const float* src = pk + pl.sup + ((size_t)s * L + k0r + kk) * a.Fp + f0;
                float* dst = Bs + (size_t)(s * c.LKc + kk) * FC;
                for (int q = lane; q < FCq; q += 32)
                    st4(dst + 4 * q, f0 + 4 * q < a.Fp ? ld4(src + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f));
            }
        }
    };

    auto zero_acc = [&]() {
#pragma unroll
        for (int r = 0; r < RN; ++r)
#pragma unroll
            for (int k = 0; k < RK; ++k) {
                accc[r][k] = 0.f;
#pragma unroll
                for (int j = 0; j < D; ++j)
#pragma unroll
                    for (int s = 0; s < D; ++s) acc[r][k][j][s] = 0.f;
            }
    };

    auto accumulate = [&](int item) {
        const int set = item / KG, kg = item % KG;
        const float* arow[RN];
#pragma unroll
        for (int r = 0; r < RN; ++r) arow[r] = As + (size_t)(set * 32 * RN + r * 32 + lane) * fsa;
        const float* brow[RK];
#pragma unroll
        for (int k = 0; k < RK; ++k) brow[k] = Bs + (size_t)min(kg * RK + k, LK - 1) * FC;
        const size_t astep = (size_t)TN * fsa, bstep = (size_t)c.LKc * FC;
#pragma unroll 2
        for (int q = 0; q < FCq; ++q) {
            float4 av[RN][D + 1];
#pragma unroll
            for (int r = 0; r < RN; ++r)
#pragma unroll
                for (int j = 0; j <= D; ++j) av[r][j] = ld4(arow[r] + j * astep + 4 * q);
#pragma unroll
            for (int k = 0; k < RK; ++k) {
#pragma unroll
                for (int s = 0; s <= D; ++s) {
                    const float4 b = ld4(brow[k] + s * bstep + 4 * q);
#pragma unroll
                    for (int r = 0; r < RN; ++r) {
                        if (s < D) {
#pragma unroll
                            for (int j = 0; j < D; ++j) fma4(acc[r][k][j][s < D ? s : 0], av[r][j], b);
                        } else {
                            fma4(accc[r][k], av[r][D], b);
                        }
                    }
                }
            }
        }
    };

    auto epilogue = [&](int item) {
        const int set = item / KG, kg = item % KG;
        const int8_t* supsign = reinterpret_cast<const int8_t*>(pk + pl.sign);
#pragma unroll
        for (int r = 0; r < RN; ++r) {
            const int nl = set * 32 * RN + r * 32 + lane;
            const int row = node0 + nl;
            if (row >= n) continue;
#pragma unroll
            for (int k = 0; k < RK; ++k) {
                const int kk = kg * RK + k;
                if (kk >= LK) continue;
                const int kglob = k0r + kk;
                const size_t cidx = (size_t)a.scoff[D - 1] + (size_t)row * L + kglob;
                const int forced = a.argmax_in ? (a.argmax_in[cidx] & 0x7f) : -1;
                // mean over j for every permutation: sequential sum, then true division (kernels.py:194)
                float best = 0.f, used = 0.f;
                int bi = 0;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    float s = acc[r][k][0][Perm<D>::at(p, 0)];
#pragma unroll
                    for (int j = 1; j < D; ++j) s += acc[r][k][j][Perm<D>::at(p, j)];
                    s = div_deg<D>(s);
                    if (p == 0 || s > best) { best = s; bi = p; }   // first maximum wins (torch.max, kernels.py:373)
                    if (p == forced) used = s;
                }
                if (a.argmax_free) a.argmax_free[cidx] = (uint8_t)bi;
                if (forced >= 0 && forced < P) { bi = forced; best = used; }
                // bond-attribute cosine at the chosen permutation (kernels.py:382-390)
                uint32_t code = 0;
#pragma unroll
                for (int p = 0; p < P; ++p) if (p == bi) code = perm_code<D>(p);
                float esum = 0.f;
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const int s = (code >> (2 * j)) & 3;
                    const float* en = Es + (size_t)(nl * D + j) * EP;
                    const float* es = ESs + (size_t)(s * c.LKc + kk) * EP;
                    const float4 e0 = ld4(en), e1 = ld4(en + 4), s0 = ld4(es), s1 = ld4(es + 4);
                    float dd = 0.f;
                    fma4(dd, e0, s0);
                    fma4(dd, e1, s1);
                    esum = j == 0 ? dd : esum + dd;
                }
                const float E = div_deg<D>(esum);
                float sc = div_by((best * ws + accc[r][k] * wc) + E * we, W, rW);
                uint8_t am = (uint8_t)bi;
                if (D == 4 && a.is_last) {
                    // chirality (kernels.py:279-350): +1 if any two neighbours are identical, else sign agreement
                    int chi = 1;
                    if (!dupf[nl]) chi = (a.tsign[row] == supsign[kglob * 12 + bi]) ? 1 : -1;
                    if (chi < 0) { sc = -sc; am |= 0x80; }
                }
                a.argmax[cidx] = am;
                if (a.sc_mode == 0) a.sc[cidx] = sc;
                else a.sc[(size_t)rowNode[D * TN + nl] * a.ld_sc + a.koff[D - 1] + kglob] = sc;
            }
        }
    };

    if (a.nfc == 1) {
        stage(0);
        __syncthreads();
        for (int item = warp; item < nitems; item += FWD_WARPS) {
            zero_acc();
            accumulate(item);
            epilogue(item);
        }
    } else {
        zero_acc();
        for (int fc = 0; fc < a.nfc; ++fc) {
            if (fc) __syncthreads();
            stage(fc);
            __syncthreads();
            if (warp < nitems) accumulate(warp);
        }
        if (warp < nitems) epilogue(warp);
    }
}

__global__ void __launch_bounds__(FWD_THREADS, 1) k_conv_fwd(const __grid_constant__ FwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_tile;
    int staged_key = -1;
    const int total = a.tile_begin[4];
    while (true) {
        if (threadIdx.x == 0) s_tile = atomicAdd(a.counter, 1);
        __syncthreads();
        const int t = s_tile;
        __syncthreads();
        if (t >= total) break;
        if (t < a.tile_begin[1]) fwd_tile<4>(a, smem, t - a.tile_begin[0], staged_key);
        else if (t < a.tile_begin[2]) fwd_tile<3>(a, smem, t - a.tile_begin[1], staged_key);
        else if (t < a.tile_begin[3]) fwd_tile<2>(a, smem, t - a.tile_begin[2], staged_key);
        else fwd_tile<1>(a, smem, t - a.tile_begin[3], staged_key);
    }
}

// ---- host-side configuration -----------------------------------------------------------------------------------
static int fwd_rn(int d) { return d == 1 ? FwdTile<1>::RN : d == 2 ? FwdTile<2>::RN : d == 3 ? FwdTile<3>::RN : FwdTile<4>::RN; }
static int fwd_rk(int d) { return d == 1 ? FwdTile<1>::RK : d == 2 ? FwdTile<2>::RK : d == 3 ? FwdTile<3>::RK : FwdTile<4>::RK; }

static int odd_quads(int fl) { return ((fl / 4) % 2 == 0) ? fl + 4 : fl; }

// Chooses chunking so that everything fits in `budget` bytes of shared memory.  The layout is per degree (a CTA works
// on one degree at a time); returns the maximum over degrees of the bytes needed, or -1.
static int64_t fwd_configure(const molkgnn_layer_t* layer, int budget, FwdArgs* a) {
    const int Fp = layer->Fp;
    for (int attempt = 0; attempt < 2; ++attempt) {
        const bool resident = (attempt == 0);
        const int FC = resident ? Fp : (Fp > 64 ? 64 : Fp);
        const int nfc = (Fp + FC - 1) / FC;
        const int fsa = odd_quads(FC);
        int64_t total = 0;
        bool ok = true;
        for (int d = 1; d <= 4; ++d) {
            const int L = layer->L[d - 1];
            FwdCfg& c = a->cfg[d - 1];
            const int RN = fwd_rn(d), RK = fwd_rk(d);
            if (L == 0) { c = FwdCfg{1, 0, 1, 32 * RN, 1, 0, 0, 0, 0, 0, 0}; continue; }
            const int LKc = resident ? L : std::min(L, FWD_WARPS * RK);
            const int KGc = (LKc + RK - 1) / RK;
            // with several feature chunks the accumulators of an item live across chunks: one item per warp
            int sets = std::max(1, FWD_WARPS / KGc);
            int64_t need = 0;
            while (true) {
                const int TN = sets * 32 * RN;
                int64_t off = (int64_t)(d + 1) * LKc * FC * 4;                 // B rows
                c.sm_A = (int)off;       off += (int64_t)(d + 1) * TN * fsa * 4;
                c.sm_E = (int)off;       off += (int64_t)TN * d * EP * 4;
                c.sm_ES = (int)off;      off += (int64_t)d * LKc * EP * 4;
                c.sm_rownode = (int)off; off += (int64_t)(d + 1) * TN * 4;
                c.sm_rinv = (int)off;    off += (int64_t)(d + 1) * TN * 8;     // reciprocal + clamped norm
                c.sm_dup = (int)off;     off += (TN + 15) / 16 * 16;
                need = off;
                if (need <= budget || sets == 1) break;
                --sets;
            }
            if (need > budget) { ok = false; break; }
            c.LKc = LKc; c.nkr = (L + LKc - 1) / LKc; c.sets = sets; c.TN = sets * 32 * RN; c.KGc = KGc;
            total = std::max(total, need);
        }
        if (!ok) continue;
        a->FC = FC; a->nfc = nfc; a->fsa = fsa;
        return std::max<int64_t>(total, 16);
    }
    return -1;
}

int launch_conv_fwd_tc(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                       const float* xnorm, int32_t is_last_layer, float* sc, int32_t sc_mode, int32_t ld_sc,
                       const int64_t scoff[4], uint8_t* argmax, uint8_t* argmax_free, const uint8_t* argmax_in,
                       int32_t* counter, cudaStream_t st);
int launch_conv_fwd_tile(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                         const void* ximg, int32_t is_last_layer, float* sc, int32_t sc_mode, int32_t ld_sc,
                         const int64_t scoff[4], uint8_t* argmax, uint8_t* argmax_free, const uint8_t* argmax_in,
                         uint8_t* argmax_tile, int32_t* counter, cudaStream_t st);

int launch_conv_fwd_wide(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                         const void* ximg, int32_t is_last_layer, float* sc, int32_t sc_mode, const int64_t scoff[4],
                         uint8_t* argmax, uint8_t* argmax_free, const uint8_t* argmax_in, cudaStream_t st);

extern long long g_path_counts[4];

}  // namespace mk

using namespace mk;

// 1 = tensor-core kernel (default), 0 = fp32 SIMT kernel (also the automatic choice for layers whose kernel set does
// not fit the resident tensor-core design).  MOLKGNN_FWD=simt in the environment selects 0 at load time.
static int g_fwd_path = -1;
extern "C" int molkgnn_set_fwd_path(int path) {
    const int old = g_fwd_path;
    g_fwd_path = path < 0 ? 0 : path > 3 ? 3 : path;
    return old;
}

static void resolve_fwd_path() {
    if (g_fwd_path < 0) {
        const char* e = getenv("MOLKGNN_FWD");
        // simt | bucket-order tc | tile (per layer) | layer-fused tile kernel for the stack, per-layer tile kernel otherwise (default)
        g_fwd_path = (e && e[0] == 's') ? 0 : (e && e[0] == 'b') ? 1 : (e && e[0] == 't') ? 2 : 3;
    }
}
extern "C" int molkgnn_get_fwd_path(void) {
    resolve_fwd_path();
    return g_fwd_path;
}

extern "C" int64_t molkgnn_conv_fwd_smem_bytes(const molkgnn_layer_t* layer) {
    FwdArgs a;
    int budget = device_max_smem_optin();
    if (budget <= 0) budget = 227 * 1024;
    return fwd_configure(layer, budget - 1024, &a);
}

extern "C" int molkgnn_conv_fwd(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                                const float* xnorm, int32_t is_last_layer, float* sc, int32_t sc_mode, int32_t ld_sc,
                                const int64_t scoff[4], uint8_t* argmax, uint8_t* argmax_free,
                                const uint8_t* argmax_in, int32_t* counter, const void* ximg, uint8_t* argmax_tile,
                                void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    MK_REQUIRE(ldx % 4 == 0 && ldx >= layer->Fp, "conv_fwd: ldx=%d must be a multiple of 4 and >= Fp=%d", ldx, layer->Fp);
    MK_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "conv_fwd: x must be 16-byte aligned");
    static int s_budget = 0, s_sms = 0;
    if (!s_budget) {
        s_budget = device_max_smem_optin();
        s_sms = device_num_sms();
        MK_REQUIRE(s_budget > 0 && s_sms > 0, "conv_fwd: no CUDA device");
    }
    ProfScope prof("conv_fwd", st);
    resolve_fwd_path();
    if (g_fwd_path >= 2) {
        const int rc = launch_conv_fwd_tile(plan, layer, x, ldx, ximg, is_last_layer, sc, sc_mode, ld_sc, scoff, argmax,
                                            argmax_free, argmax_in, argmax_tile, counter, st);
        if (rc > 0) ++g_path_counts[0];
        if (rc != 0) return rc < 0 ? rc : 0;
        // wide layers (more blocks / features than the resident-operand kernels take): both operands streamed
        const int rw = launch_conv_fwd_wide(plan, layer, x, ldx, ximg, is_last_layer, sc, sc_mode, scoff, argmax, argmax_free,
                                            argmax_in, st);
        if (rw > 0) ++g_path_counts[0];
        if (rw != 0) return rw < 0 ? rw : 0;
    }
    ++g_path_counts[1];
    if (g_fwd_path >= 1) {
        const int rc = launch_conv_fwd_tc(plan, layer, x, ldx, xnorm, is_last_layer, sc, sc_mode, ld_sc, scoff, argmax,
                                          argmax_free, argmax_in, counter, st);
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    FwdArgs a;
    const int64_t smem = fwd_configure(layer, s_budget - 1024, &a);
    MK_REQUIRE(smem > 0, "conv_fwd: layer does not fit in shared memory (Fp=%d)", layer->Fp);
    a.x = x; a.xnorm = xnorm; a.ldx = ldx;
    a.sel = plan->sel; a.nei = plan->nei; a.ehat = plan->ehat; a.tsign = plan->tsign;
    a.F = layer->F; a.Fp = layer->Fp;
    int tb = 0;
    for (int q = 0; q < 4; ++q) {
        const int d = 4 - q;
        a.tile_begin[q] = tb;
        const int n = plan->n[d - 1], L = layer->L[d - 1];
        if (n > 0 && L > 0) tb += ((n + a.cfg[d - 1].TN - 1) / a.cfg[d - 1].TN) * a.cfg[d - 1].nkr;
    }
    a.tile_begin[4] = tb;
    for (int d = 0; d < 4; ++d) {
        a.n[d] = plan->n[d]; a.boff[d] = plan->boff[d]; a.eoff[d] = plan->eoff[d];
        a.L[d] = layer->L[d]; a.koff[d] = layer->koff[d];
        a.packed[d] = layer->packed[d];
        a.scoff[d] = scoff[d];
    }
    a.is_last = is_last_layer;
    a.sc = sc; a.sc_mode = sc_mode; a.ld_sc = ld_sc;
    a.argmax = argmax; a.argmax_free = argmax_free; a.argmax_in = argmax_in;
    a.counter = counter;
    if (tb == 0) return 0;
    MK_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
    static int64_t s_attr_dev[16] = {0};
    int64_t& s_attr = s_attr_dev[device_index()];      // function attributes are per device
    if (smem > s_attr) {
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_conv_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        s_attr = smem;
    }
    const int grid = std::min(tb, s_sms);
    count_launches(1);
    k_conv_fwd<<<grid, FWD_THREADS, smem, st>>>(a);
    MK_CHECK_CUDA(cudaGetLastError());
    return 0;
}
