// Backward of the molecular-kernel convolution (the reference relies on autograd over kernels.py:353-425 and
// KernelLayer.py:119).  Gradients are routed through the SAVED arg-max permutation (torch.max backward).
//
//   k_bwd_w   bucketed by focal degree.  Per node tile: g[n,k] = incoming gradient (optionally summed over the node's
//             neighbours = transpose of propagate, KernelLayer.py:119), a = chi*g written to `coef`; kernel-parameter
//             gradients  G_sup[k,s] += a * [xhat | ehat](n, pi^-1(s)),  G_cen[k] += a * xhat_n  accumulate in registers
//             over all tiles of the CTA and leave as ONE partial copy per CTA (no global atomics); k_param_finalize
//             (params.cu) reduces the copies in fixed order and applies the chain rule of the normalisation / softmax.
//   k_bwd_x   node ordered "pull": one warp (or sub-warp) owns node v and sums, in fixed order, its focal term and the
//             terms of every neighbourhood it belongs to, using the normalised kernel rows of ALL degrees resident in
//             shared memory; then applies d(x/|x|)/dx and writes the row once.  Deterministic, no atomics.
#include <algorithm>
#include "common.cuh"

namespace mk {

int launch_param_finalize(const molkgnn_layer_t* layer, const float* partials, const int64_t part_off[4],
                          const int ncta[4], float* q_scratch, const molkgnn_layer_grads_t* grads, int prescaled,
                          cudaStream_t st);
int launch_conv_bwd_tile(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                         const float* xnorm, const void* ximg, const float* grad, int32_t ldg, int32_t grad_mode,
                         const uint8_t* argmax, const int64_t scoff[4], float* coef, float* partials, float* scratch,
                         float* grad_x, int32_t ldgx, int64_t part_off[4], int ncta[4], int64_t* part_total, bool do_launch,
                         const uint8_t* argmax_tile, cudaStream_t st);
int launch_conv_bwd_wide(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx, const float* xnorm,
                         const void* ximg, const float* grad, int32_t ldg, int32_t grad_mode, const uint8_t* argmax,
                         const int64_t scoff[4], float* coef, float* partials, float* grad_x, int32_t ldgx, int64_t part_off[4],
                         int ncta[4], int64_t* part_total, bool do_launch, cudaStream_t st);
int64_t wide_bwd_partial_floats(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer);
int64_t wide_bwd_coef_floats(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer);
int tile_bwd_grid(const molkgnn_plan_t* plan);
bool tile_bwd_ok(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer);
long long g_path_counts[4] = {0, 0, 0, 0};   // forward tile / other, backward tile / other

// =============================================================================================================
// k_bwd_w
// =============================================================================================================
constexpr int BW_THREADS = 512;
constexpr int BW_PP = 3;   // (kernel, feature-quad) pairs per thread

struct BwdWCfg {
    int TN;     // nodes per tile
    int LKc;    // kernels per range
    int nkr;
};

struct BwdWArgs {
    const float* x; const float* xnorm; int ldx;
    const int* sel; const int* nei; const float* ehat;
    int n[4], boff[4], eoff[4], L[4], koff[4];
    long long scoff[4];
    const float* grad; int ldg; int grad_mode;
    const uint8_t* argmax;
    float* coef;
    float* partials; long long part_off[4];
    int cta_begin[5];
    BwdWCfg cfg[4];
    int F, Fp, FW;      // FW = Fp + EP
    int FC, nfc, fsw;   // chunk of the FW-wide [xhat | ehat] row
    int sm_A, sm_coef, sm_perm, sm_rownode;
};

template <int D>
__device__ __forceinline__ void bwd_w_body(const BwdWArgs& a, unsigned char* smem, int cta_local, int ncta) {
    constexpr int P = Perm<D>::P;
    const BwdWCfg c = a.cfg[D - 1];
    const int L = a.L[D - 1], n = a.n[D - 1], TN = c.TN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = BW_THREADS / 32;
    float* As = reinterpret_cast<float*>(smem + a.sm_A);
    float* coefS = reinterpret_cast<float*>(smem + a.sm_coef);
    unsigned char* permS = smem + a.sm_perm;
    int* rowNode = reinterpret_cast<int*>(smem + a.sm_rownode);
    float* rowNrm = reinterpret_cast<float*>(rowNode + (D + 1) * TN);
    float* rowRinv = rowNrm + (D + 1) * TN;
    __shared__ unsigned char invcode[16];
    if (tid < P) {
        uint32_t code = 0;
#pragma unroll
        for (int p = 0; p < P; ++p) if (p == tid) code = perm_inv_code<D>(p);
        invcode[tid] = (unsigned char)code;
    }
    const int eoff = a.eoff[D - 1], boff = a.boff[D - 1], koff = a.koff[D - 1];
    const int ntiles = (n + TN - 1) / TN;
    const int FCq = a.FC / 4, fsw = a.fsw;
    const int rows_x = (D + 1) * L;
    float* part = a.partials + a.part_off[D - 1] + (size_t)cta_local * rows_x * a.FW;

    for (int kr = 0; kr < c.nkr; ++kr) {
        const int k0r = kr * c.LKc;
        const int LK = min(c.LKc, L - k0r);
        for (int fc = 0; fc < a.nfc; ++fc) {
            const int f0 = fc * a.FC;
            const int npairs = LK * FCq;
            float4 acc[BW_PP][D + 1];
#pragma unroll
            for (int i = 0; i < BW_PP; ++i)
#pragma unroll
                for (int s = 0; s <= D; ++s) acc[i][s] = make_float4(0.f, 0.f, 0.f, 0.f);

            for (int tile = cta_local; tile < ntiles; tile += ncta) {
                const int node0 = tile * TN;
                const int nv = min(TN, n - node0);
                __syncthreads();   // previous tile fully consumed
                for (int i = tid; i < (D + 1) * TN; i += BW_THREADS) {
                    int j = i / TN, nl = i % TN, r = node0 + nl;
                    int node = -1;
                    if (r < n) node = j < D ? a.nei[(size_t)eoff + (size_t)r * D + j] : a.sel[boff + r];
                    rowNode[i] = node;
                    const float nrm = node >= 0 ? fmaxf(a.xnorm[node], MOLKGNN_COS_EPS) : 1.0f;
                    rowNrm[i] = nrm;
                    rowRinv[i] = 1.0f / nrm;
                }
                __syncthreads();
                // ---- incoming gradient -> a[n,k] = chi * g ----
                for (int i = tid; i < nv * LK; i += BW_THREADS) {
                    const int nl = i / LK, kk = i % LK, k = k0r + kk;
                    const int col = koff + k;
                    float g;
                    if (a.grad_mode == 0) {
                        g = a.grad[(size_t)rowNode[D * TN + nl] * a.ldg + col];
                    } else {
                        g = a.grad[(size_t)rowNode[nl] * a.ldg + col];
#pragma unroll
                        for (int j = 1; j < D; ++j) g += a.grad[(size_t)rowNode[j * TN + nl] * a.ldg + col];
                    }
                    const size_t cidx = (size_t)a.scoff[D - 1] + (size_t)(node0 + nl) * L + k;
                    const uint8_t am = a.argmax[cidx];
                    const float av = (am & 0x80) ? -g : g;
                    coefS[nl * c.LKc + kk] = av;
                    permS[nl * c.LKc + kk] = invcode[am & 0x7f];
                    if (fc == 0) a.coef[cidx] = av;
                }
                // ---- stage [xhat | ehat] rows (chunk f0 .. f0+FC of the FW-wide row); UNR rows in flight per warp ----
                {
                    constexpr int UNR = 8;
                    const int nrows = (D + 1) * TN;
                    for (int q = lane; q < FCq; q += 32) {
                        const int w = f0 + 4 * q;
                        for (int row0 = warp; row0 < nrows; row0 += NW * UNR) {
                            float4 v[UNR];
                            float ri[UNR], nr[UNR];
#pragma unroll
                            for (int u = 0; u < UNR; ++u) {
                                const int row = row0 + u * NW;
                                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                                ri[u] = 1.f; nr[u] = 1.f;
                                if (row < nrows) {
                                    const int node = rowNode[row];
                                    const int j = row / TN, nl = row - j * TN;
                                    if (node >= 0) {
                                        if (w < a.Fp) {
                                            ri[u] = rowRinv[row]; nr[u] = rowNrm[row];
                                            v[u] = ld4(a.x + (size_t)node * a.ldx + w);
                                        } else if (w < a.FW && j < D) {
                                            v[u] = ld4(a.ehat + ((size_t)eoff + (size_t)(node0 + nl) * D + j) * EP + (w - a.Fp));
                                        }
                                    }
                                }
                            }
#pragma unroll
                            for (int u = 0; u < UNR; ++u) {
                                const int row = row0 + u * NW;
                                if (row < nrows) {
                                    float4 o;
                                    o.x = div_by(v[u].x, nr[u], ri[u]); o.y = div_by(v[u].y, nr[u], ri[u]);
                                    o.z = div_by(v[u].z, nr[u], ri[u]); o.w = div_by(v[u].w, nr[u], ri[u]);
                                    st4(As + (size_t)row * fsw + 4 * q, o);
                                }
                            }
                        }
                    }
                }
                __syncthreads();
                // ---- accumulate ----
#pragma unroll
                for (int i = 0; i < BW_PP; ++i) {
                    const int p = tid + i * BW_THREADS;
                    if (p < npairs) {
                        const int kk = p / FCq, fq = p % FCq;
                        const float* abase = As + 4 * fq;
                        for (int nl = 0; nl < nv; ++nl) {
                            const float av = coefS[nl * c.LKc + kk];
                            const uint32_t ic = permS[nl * c.LKc + kk];
#pragma unroll
                            for (int s = 0; s < D; ++s) {
                                const int j = (ic >> (2 * s)) & 3;
                                const float4 xv = ld4(abase + (size_t)(j * TN + nl) * fsw);
                                acc[i][s].x = fmaf(av, xv.x, acc[i][s].x);
                                acc[i][s].y = fmaf(av, xv.y, acc[i][s].y);
                                acc[i][s].z = fmaf(av, xv.z, acc[i][s].z);
                                acc[i][s].w = fmaf(av, xv.w, acc[i][s].w);
                            }
                            const float4 xc = ld4(abase + (size_t)(D * TN + nl) * fsw);
                            acc[i][D].x = fmaf(av, xc.x, acc[i][D].x);
                            acc[i][D].y = fmaf(av, xc.y, acc[i][D].y);
                            acc[i][D].z = fmaf(av, xc.z, acc[i][D].z);
                            acc[i][D].w = fmaf(av, xc.w, acc[i][D].w);
                        }
                    }
                }
            }
            // ---- one partial copy per CTA ----
#pragma unroll
            for (int i = 0; i < BW_PP; ++i) {
                const int p = tid + i * BW_THREADS;
                if (p < npairs) {
                    const int kk = p / FCq, fq = p % FCq;
                    const int w = f0 + 4 * fq;
                    if (w < a.FW) {
#pragma unroll
                        for (int s = 0; s <= D; ++s)
                            st4(part + ((size_t)s * L + k0r + kk) * a.FW + w, acc[i][s]);
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(BW_THREADS, 1) k_bwd_w(const __grid_constant__ BwdWArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int b = blockIdx.x;
    if (b < a.cta_begin[1]) bwd_w_body<1>(a, smem, b - a.cta_begin[0], a.cta_begin[1] - a.cta_begin[0]);
    else if (b < a.cta_begin[2]) bwd_w_body<2>(a, smem, b - a.cta_begin[1], a.cta_begin[2] - a.cta_begin[1]);
    else if (b < a.cta_begin[3]) bwd_w_body<3>(a, smem, b - a.cta_begin[2], a.cta_begin[3] - a.cta_begin[2]);
    else bwd_w_body<4>(a, smem, b - a.cta_begin[3], a.cta_begin[4] - a.cta_begin[3]);
}

static const int kBwTN[4] = {64, 64, 32, 32};

// static CTA partition across degrees proportional to n_d * L_d * (d+1); at most one CTA per tile
static void bwd_w_partition(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, int nsm, int ncta[4]) {
    double cost[4], tot = 0;
    int ntiles[4];
    for (int d = 0; d < 4; ++d) {
        ntiles[d] = (plan->n[d] > 0 && layer->L[d] > 0) ? (plan->n[d] + kBwTN[d] - 1) / kBwTN[d] : 0;
        cost[d] = ntiles[d] ? (double)plan->n[d] * layer->L[d] * (d + 2) : 0.0;
        tot += cost[d];
    }
    for (int d = 0; d < 4; ++d) {
        if (!ntiles[d]) { ncta[d] = 0; continue; }
        int c = (int)(nsm * cost[d] / tot + 0.5);
        ncta[d] = std::max(1, std::min(c, ntiles[d]));
    }
}

static int64_t bwd_w_configure(const molkgnn_layer_t* layer, int budget, BwdWArgs* a) {
    const int FW = layer->Fp + EP;
    a->F = layer->F; a->Fp = layer->Fp; a->FW = FW;
    // single F chunk if the widest tile fits, otherwise 64-float chunks
    for (int attempt = 0; attempt < 2; ++attempt) {
        const int FC = attempt == 0 ? FW : 64;
        const int nfc = (FW + FC - 1) / FC;
        const int fsw = ((FC / 4) % 2 == 0) ? FC + 4 : FC;
        const int FCq = FC / 4;
        int64_t mA = 0, mC = 0, mP = 0, mR = 0;
        for (int d = 1; d <= 4; ++d) {
            const int L = std::max(1, layer->L[d - 1]);
            const int TN = kBwTN[d - 1];
            int LKc = std::min(L, std::max(1, BW_THREADS * BW_PP / FCq));
            a->cfg[d - 1] = BwdWCfg{TN, LKc, (L + LKc - 1) / LKc};
            mA = std::max<int64_t>(mA, (int64_t)(d + 1) * TN * fsw * 4);
            mC = std::max<int64_t>(mC, (int64_t)TN * LKc * 4);
            mP = std::max<int64_t>(mP, ((int64_t)TN * LKc + 15) / 16 * 16);
            mR = std::max<int64_t>(mR, (int64_t)(d + 1) * TN * 12);   // node id + clamped norm + reciprocal
        }
        int64_t off = 0;
        a->sm_A = (int)off; off += mA;
        a->sm_coef = (int)off; off += mC;
        a->sm_perm = (int)off; off += mP;
        a->sm_rownode = (int)off; off += mR;
        if (off <= budget) {
            a->FC = FC; a->nfc = nfc; a->fsw = fsw;
            return off;
        }
    }
    return -1;
}

// =============================================================================================================
// k_bwd_x
// =============================================================================================================
constexpr int BX_THREADS = 1024;   // 32 warps: the resident kernel rows allow one CTA per SM, so TLP comes from the block

struct BwdXArgs {
    const float* x; const float* xnorm; int ldx;
    const int* deg; const int* pos; const int* in_cnt; const int* in_src; const int* in_j;
    const float* coef; const uint8_t* argmax;
    int N, F, Fp;
    int L[4];
    long long scoff[4];
    const float* packed[4];
    int klo[4], khi[4];       // resident kernel range of this pass
    int sm_row0[4];           // first resident row (in units of Fp floats) of degree d in smem
    int first, last;
    float* gx; int ldgx;
};

template <int LPN, int NACC>
__global__ void __launch_bounds__(BX_THREADS, 1) k_bwd_x(const __grid_constant__ BwdXArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* Bs = reinterpret_cast<float*>(smem);
    __shared__ unsigned char permtab[4][12];
    __shared__ float wsc[4][2];   // per degree: ws/W/d, wc/W
    const int tid = threadIdx.x;
    const int Fp = a.Fp, FQ = Fp / 4;
    if (tid < 48) {
        const int d = tid / 12 + 1, p = tid % 12;
        uint32_t code = 0;
        if (d == 1) code = 0;
        else if (d == 2) code = p < 2 ? perm_code<2>(p) : 0;
        else if (d == 3) code = p < 6 ? perm_code<3>(p) : 0;
        else code = perm_code<4>(p);
        permtab[d - 1][p] = (unsigned char)code;
    }
    if (tid >= 64 && tid < 68) {
        const int d = tid - 64 + 1;
        if (a.L[d - 1] > 0) {
            PackedLayout pl(d, a.L[d - 1], Fp);
            const float* pk = a.packed[d - 1];
            wsc[d - 1][0] = pk[pl.w + 0] / pk[pl.w + 3] / (float)d;
            wsc[d - 1][1] = pk[pl.w + 1] / pk[pl.w + 3];
        }
    }
    // ---- resident rows: for each degree, support rows (s, k in range) then centre rows (k in range) ----
    for (int d = 1; d <= 4; ++d) {
        const int L = a.L[d - 1];
        const int LK = a.khi[d - 1] - a.klo[d - 1];
        if (L == 0 || LK <= 0) continue;
        PackedLayout pl(d, L, Fp);
        const float* src = a.packed[d - 1] + pl.sup;
        float* dst = Bs + (size_t)a.sm_row0[d - 1] * Fp;
        const int nq = (d + 1) * LK * FQ;
        for (int i = tid; i < nq; i += BX_THREADS) {
            const int row = i / FQ, q = i % FQ;
            const int s = row / LK, kk = row % LK;
            st4(dst + (size_t)row * Fp + 4 * q, ld4(src + ((size_t)s * L + a.klo[d - 1] + kk) * Fp + 4 * q));
        }
    }
    __syncthreads();

    constexpr int NPW = 32 / LPN;                 // nodes per warp
    const int lane = tid & 31, warp = tid >> 5;
    const int gl = lane % LPN, grp = lane / LPN;  // lane inside the node group, group inside the warp
    const unsigned gmask = LPN == 32 ? 0xffffffffu : (((1u << LPN) - 1u) << (grp * LPN));
    const int nodes_per_cta = (BX_THREADS / 32) * NPW;
    for (int base = blockIdx.x * nodes_per_cta; base < a.N; base += gridDim.x * nodes_per_cta) {
        const int v = base + warp * NPW + grp;
        const bool valid = v < a.N;
        float4 acc[NACC];
#pragma unroll
        for (int c = 0; c < NACC; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int cnt = valid ? min(a.in_cnt[v], 4) : 0;
        // ---- role table: role 0 = focal term, roles 1..cnt = neighbourhoods of the in-neighbours (edge order).
        // Lane r of the node group fetches role r's metadata, so the dependent loads of all roles overlap. ----
        int m_off = -1, m_j = 0, m_d = 0;      // m_off: offset of the role's coefficient row (floats), -1 = none
        if (valid && gl <= cnt) {
            int u = v;
            if (gl > 0) { u = a.in_src[4 * v + gl - 1]; m_j = a.in_j[4 * v + gl - 1]; }
            if (u >= 0) {
                m_d = a.deg[u];
                if (m_d >= 1 && m_d <= 4 && a.L[m_d - 1] > 0 && a.khi[m_d - 1] > a.klo[m_d - 1])
                    m_off = (int)(a.scoff[m_d - 1] + (long long)a.pos[u] * a.L[m_d - 1]);
            }
        }
        for (int t = 0; t <= cnt; ++t) {
            const int off = __shfl_sync(gmask, m_off, grp * LPN + t);
            const int j = __shfl_sync(gmask, m_j, grp * LPN + t);
            const int d = __shfl_sync(gmask, m_d, grp * LPN + t);
            if (off < 0) continue;
            const int klo = a.klo[d - 1], khi = a.khi[d - 1], LK = khi - klo;
            const float scale = t == 0 ? wsc[d - 1][1] : wsc[d - 1][0];
            const float* rows = Bs + (size_t)a.sm_row0[d - 1] * Fp + 4 * gl;
            const unsigned pj = 2 * j;
            // batches of KB kernels: every lane of the group holds KB/LPN (coefficient, row offset) pairs
            constexpr int KB = 32, MR = KB / LPN;
            for (int k0 = klo; k0 < khi; k0 += KB) {
                float my_a[MR];
                int my_r[MR];
#pragma unroll
                for (int m = 0; m < MR; ++m) {
                    const int k = k0 + m * LPN + gl;
                    my_a[m] = 0.f;
                    my_r[m] = 0;
                    if (k < khi) {
                        my_a[m] = a.coef[off + k] * scale;
                        // centre rows sit after the d support blocks; supports: s = perm[pi][j]
                        const int s = t == 0 ? d : (permtab[d - 1][a.argmax[off + k] & 0x7f] >> pj) & 3;
                        my_r[m] = (s * LK + (k - klo)) * Fp;
                    }
                }
                // coefficients beyond khi are zero and point at row 0: no branches in the FMA loop
#pragma unroll
                for (int m = 0; m < MR; ++m) {
                    if (k0 + m * LPN >= khi) break;
#pragma unroll 4
                    for (int i = 0; i < LPN; ++i) {
                        const float av = __shfl_sync(gmask, my_a[m], grp * LPN + i);
                        const int ro = __shfl_sync(gmask, my_r[m], grp * LPN + i);
                        const float* rp = rows + ro;
#pragma unroll
                        for (int c = 0; c < NACC; ++c) {
                            if (gl + c * LPN < FQ) {
                                const float4 b = ld4(rp + 4 * c * LPN);
                                acc[c].x = fmaf(av, b.x, acc[c].x);
                                acc[c].y = fmaf(av, b.y, acc[c].y);
                                acc[c].z = fmaf(av, b.z, acc[c].z);
                                acc[c].w = fmaf(av, b.w, acc[c].w);
                            }
                        }
                    }
                }
            }
        }
        if (!valid) continue;   // whole node group is invalid together (shuffles above stay inside the group)
        float* out = a.gx + (size_t)v * a.ldgx;
        if (!a.first) {
#pragma unroll
            for (int c = 0; c < NACC; ++c) {
                const int q = gl + c * LPN;
                if (q < FQ) { float4 o = ld4(out + 4 * q); acc[c].x += o.x; acc[c].y += o.y; acc[c].z += o.z; acc[c].w += o.w; }
            }
        }
        if (a.last) {
            // chain rule through xhat = x / max(|x|, eps):  gx = (g - (xhat.g) xhat) / |x|
            const float nrm = a.xnorm[v];
            const float den = fmaxf(nrm, MOLKGNN_COS_EPS);
            const float rden = 1.0f / den;
            float4 xh[NACC];
            float dot = 0.f;
#pragma unroll
            for (int c = 0; c < NACC; ++c) {
                const int q = gl + c * LPN;
                xh[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (q < FQ) {
                    float4 xv = ld4(a.x + (size_t)v * a.ldx + 4 * q);
                    xh[c] = make_float4(div_by(xv.x, den, rden), div_by(xv.y, den, rden), div_by(xv.z, den, rden),
                                        div_by(xv.w, den, rden));
                    dot += acc[c].x * xh[c].x + acc[c].y * xh[c].y + acc[c].z * xh[c].z + acc[c].w * xh[c].w;
                }
            }
#pragma unroll
            for (int o = LPN / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(gmask, dot, o);
            const bool clamped = !(nrm > MOLKGNN_COS_EPS);
#pragma unroll
            for (int c = 0; c < NACC; ++c) {
                const int q = gl + c * LPN;
                if (q < FQ) {
                    float4 g;
                    if (clamped) g = make_float4(acc[c].x * rden, acc[c].y * rden, acc[c].z * rden, acc[c].w * rden);
                    else g = make_float4((acc[c].x - dot * xh[c].x) * rden, (acc[c].y - dot * xh[c].y) * rden,
                                         (acc[c].z - dot * xh[c].z) * rden, (acc[c].w - dot * xh[c].w) * rden);
                    const int f = 4 * q;
                    if (f + 0 >= a.F) g.x = 0.f;
                    if (f + 1 >= a.F) g.y = 0.f;
                    if (f + 2 >= a.F) g.z = 0.f;
                    if (f + 3 >= a.F) g.w = 0.f;
                    st4(out + f, g);
                }
            }
            for (int f = Fp + gl; f < a.ldgx; f += LPN) out[f] = 0.f;
        } else {
#pragma unroll
            for (int c = 0; c < NACC; ++c) {
                const int q = gl + c * LPN;
                if (q < FQ) st4(out + 4 * q, acc[c]);
            }
        }
    }
}

template <int LPN, int NACC>
static int launch_bwd_x(const BwdXArgs& a, int grid, int64_t smem, cudaStream_t st) {
    static int64_t s_attr_dev[16] = {0};
    int64_t& s_attr = s_attr_dev[device_index()];      // function attributes are per device
    if (smem > s_attr) {
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_bwd_x<LPN, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        s_attr = smem;
    }
    count_launches(1);
    ProfScope prof("bwd_x", st);
    k_bwd_x<LPN, NACC><<<grid, BX_THREADS, smem, st>>>(a);
    MK_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mk

using namespace mk;

static int g_budget = 0, g_sms = 0;
static int g_bwd_path = 1;   // 1 = molecule-tile tensor-core backward when eligible, 0 = bucket-order SIMT kernels
extern "C" int molkgnn_set_bwd_path(int path) {
    const int old = g_bwd_path;
    g_bwd_path = path ? 1 : 0;
    return old;
}
extern "C" void molkgnn_path_counts(int64_t out[4]) {
    for (int i = 0; i < 4; ++i) out[i] = g_path_counts[i];
}
static int init_dev() {
    if (!g_budget) {
        g_budget = device_max_smem_optin();
        g_sms = device_num_sms();
        MK_REQUIRE(g_budget > 0 && g_sms > 0, "conv_bwd: no CUDA device");
    }
    return 0;
}

namespace mk { int64_t tile_bwd_coef_floats(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer); }

namespace mk { int tile_argmax_stride(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer); }
extern "C" int64_t molkgnn_tile_argmax_bytes(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer) {
    if (!tile_bwd_ok(plan, layer)) return 0;
    return (int64_t)plan->n_tiles * mk::tile_argmax_stride(plan, layer) + 16;
}

extern "C" int64_t molkgnn_conv_bwd_coef_floats(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer) {
    int64_t tot = 0;
    for (int d = 0; d < 4; ++d) tot += (int64_t)plan->n[d] * layer->L[d];
    return std::max<int64_t>(std::max(std::max(tot, mk::tile_bwd_coef_floats(plan, layer)), wide_bwd_coef_floats(plan, layer)), 4);
}

extern "C" int64_t molkgnn_conv_bwd_partial_floats(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer) {
    if (init_dev()) return -1;
    int ncta[4];
    bwd_w_partition(plan, layer, g_sms, ncta);
    int64_t tot = 0, rows = 0, tot_tile = 0;
    const int tgrid = tile_bwd_ok(plan, layer) ? tile_bwd_grid(plan) : 0;
    for (int d = 0; d < 4; ++d) {
        tot += (int64_t)ncta[d] * (d + 2) * layer->L[d] * (layer->Fp + EP);
        tot_tile += (int64_t)tgrid * (d + 2) * layer->L[d] * (layer->Fp + EP);
        rows += (int64_t)(d + 2) * layer->L[d];
    }
    return std::max(std::max(tot, tot_tile) + 2 * rows + 16 + 512,      // + per-CTA max |coef| of k_coef_tile (tile path)
                    wide_bwd_partial_floats(plan, layer));
}

extern "C" int molkgnn_conv_bwd(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                                const float* xnorm, const float* grad, int32_t ldg, int32_t grad_mode,
                                const uint8_t* argmax, const int64_t scoff[4], float* coef, float* partials,
                                float* grad_x, int32_t ldgx, const molkgnn_layer_grads_t* grads, int32_t phases,
                                const void* ximg, float* scratch, const uint8_t* argmax_tile, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    if (init_dev()) return -1;
    MK_REQUIRE(ldx % 4 == 0 && ldx >= layer->Fp, "conv_bwd: ldx=%d must be a multiple of 4 and >= Fp=%d", ldx, layer->Fp);
    MK_REQUIRE(!grad_x || (ldgx % 4 == 0 && ldgx >= layer->Fp), "conv_bwd: ldgx=%d must be a multiple of 4 >= Fp", ldgx);
    // ---------------- molecule-tile tensor-core path ----------------
    if (ximg && g_bwd_path != 0 && (!grad_x || ldgx == layer->Fp)) {
        int64_t t_off[4], t_total = 0;
        int t_ncta[4];
        const int rc = launch_conv_bwd_tile(plan, layer, x, ldx, xnorm, ximg, grad, ldg, grad_mode, argmax, scoff, coef,
                                            partials, scratch, grad_x, ldgx, t_off, t_ncta, &t_total, (phases & 1) != 0,
                                            argmax_tile, st);
        if (rc < 0) return rc;
        if (rc == 1) {
            if (phases & 1) ++g_path_counts[2];
            if (grads && (phases & 2)) return launch_param_finalize(layer, partials, t_off, t_ncta, partials + t_total, grads, 1, st);
            return 0;
        }
    }
    // ---------------- wide layers: tensor-core backward with streamed operands (conv_bwd_wide.cu) ----------------
    if (ximg && g_bwd_path != 0 && (!grad_x || ldgx == layer->Fp)) {
        int64_t t_off[4], t_total = 0;
        int t_ncta[4];
        const int rc = launch_conv_bwd_wide(plan, layer, x, ldx, xnorm, ximg, grad, ldg, grad_mode, argmax, scoff, coef, partials,
                                            grad_x, ldgx, t_off, t_ncta, &t_total, (phases & 1) != 0, st);
        if (rc < 0) return rc;
        if (rc == 1) {
            if (phases & 1) ++g_path_counts[2];
            if (grads && (phases & 2)) return launch_param_finalize(layer, partials, t_off, t_ncta, partials + t_total, grads, 1, st);
            return 0;
        }
    }
    if (phases & 1) ++g_path_counts[3];
    // ---------------- k_bwd_w ----------------
    BwdWArgs w;
    const int64_t smem_w = bwd_w_configure(layer, g_budget - 1024, &w);
    MK_REQUIRE(smem_w > 0, "conv_bwd: tile does not fit in shared memory");
    int ncta[4];
    bwd_w_partition(plan, layer, g_sms, ncta);
    int cb = 0;
    int64_t po = 0;
    int64_t part_off[4];
    for (int d = 0; d < 4; ++d) {
        w.cta_begin[d] = cb; cb += ncta[d];
        w.part_off[d] = part_off[d] = po;
        po += (int64_t)ncta[d] * (d + 2) * layer->L[d] * w.FW;
        w.n[d] = plan->n[d]; w.boff[d] = plan->boff[d]; w.eoff[d] = plan->eoff[d];
        w.L[d] = layer->L[d]; w.koff[d] = layer->koff[d]; w.scoff[d] = scoff[d];
    }
    w.cta_begin[4] = cb;
    w.x = x; w.xnorm = xnorm; w.ldx = ldx;
    w.sel = plan->sel; w.nei = plan->nei; w.ehat = plan->ehat;
    w.grad = grad; w.ldg = ldg; w.grad_mode = grad_mode;
    w.argmax = argmax; w.coef = coef; w.partials = partials;
    if (cb > 0 && (phases & 1)) {
        static int64_t s_attr_dev[16] = {0};
    int64_t& s_attr = s_attr_dev[device_index()];      // function attributes are per device
        if (smem_w > s_attr) {
            MK_CHECK_CUDA(cudaFuncSetAttribute(k_bwd_w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
            s_attr = smem_w;
        }
        count_launches(1);
        ProfScope prof("bwd_w", st);
        k_bwd_w<<<cb, BW_THREADS, smem_w, st>>>(w);
        MK_CHECK_CUDA(cudaGetLastError());
    }
    // ---------------- parameter gradients ----------------
    if (grads && (phases & 2)) {
        // degrees without nodes: their parameter gradients are exactly zero (ncta == 0 -> finalize sums nothing)
        int rc = launch_param_finalize(layer, partials, part_off, ncta, partials + po, grads, 0, st);
        if (rc) return rc;
    }
    // ---------------- k_bwd_x ----------------
    if (grad_x && plan->N > 0 && (phases & 4)) {
        BwdXArgs b;
        b.x = x; b.xnorm = xnorm; b.ldx = ldx;
        b.deg = plan->deg; b.pos = plan->pos; b.in_cnt = plan->in_cnt; b.in_src = plan->in_src; b.in_j = plan->in_j;
        b.coef = coef; b.argmax = argmax;
        b.N = plan->N; b.F = layer->F; b.Fp = layer->Fp;
        b.gx = grad_x; b.ldgx = ldgx;
        int64_t rows_all = 0;
        for (int d = 0; d < 4; ++d) {
            b.L[d] = layer->L[d]; b.scoff[d] = scoff[d]; b.packed[d] = layer->packed[d];
            rows_all += (int64_t)(d + 2) * layer->L[d];
        }
        const int64_t row_bytes = (int64_t)layer->Fp * 4;
        const int64_t budget = g_budget - 2048;
        int npass = (int)((rows_all * row_bytes + budget - 1) / budget);
        npass = std::max(npass, 1);
        // make sure each pass fits: rows per pass = sum_d (d+1)*ceil(L_d/npass)
        while (true) {
            int64_t r = 0;
            for (int d = 0; d < 4; ++d) r += (int64_t)(d + 2) * ((layer->L[d] + npass - 1) / npass);
            if (r * row_bytes <= budget) break;
            ++npass;
            MK_REQUIRE(npass < 4096, "conv_bwd: cannot fit kernel rows in shared memory (Fp=%d)", layer->Fp);
        }
        const int FQ = layer->Fp / 4;
        for (int p = 0; p < npass; ++p) {
            int row0 = 0;
            for (int d = 0; d < 4; ++d) {
                const int step = (layer->L[d] + npass - 1) / npass;
                b.klo[d] = std::min(layer->L[d], p * step);
                b.khi[d] = std::min(layer->L[d], (p + 1) * step);
                b.sm_row0[d] = row0;
                row0 += (d + 2) * (b.khi[d] - b.klo[d]);
            }
            b.first = (p == 0); b.last = (p == npass - 1);
            const int64_t smem = std::max<int64_t>(16, (int64_t)row0 * row_bytes);
            int rc;
            if (FQ <= 8) {
                const int npc = (BX_THREADS / 32) * 4;
                rc = launch_bwd_x<8, 1>(b, std::min(g_sms, (plan->N + npc - 1) / npc), smem, st);
            } else if (FQ <= 16) {
                const int npc = (BX_THREADS / 32) * 4;
                rc = launch_bwd_x<8, 2>(b, std::min(g_sms, (plan->N + npc - 1) / npc), smem, st);
            } else if (FQ <= 32) {
                const int npc = (BX_THREADS / 32) * 2;
                rc = launch_bwd_x<16, 2>(b, std::min(g_sms, (plan->N + npc - 1) / npc), smem, st);
            } else {
                const int npc = BX_THREADS / 32;
                const int grid = std::min(g_sms, (plan->N + npc - 1) / npc);
                if (FQ <= 64) rc = launch_bwd_x<32, 2>(b, grid, smem, st);
                else if (FQ <= 128) rc = launch_bwd_x<32, 4>(b, grid, smem, st);
                else { MK_REQUIRE(false, "conv_bwd: node_attr_dim %d > 512 not supported", layer->F); }
            }
            if (rc) return rc;
        }
    }
    return 0;
}
