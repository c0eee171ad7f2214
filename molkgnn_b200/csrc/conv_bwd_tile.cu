// Molecule-tile backward of the molecular-kernel convolution (sm_100a, tcgen05 + TMEM).  The reference relies on autograd
// over kernels.py:353-425 and KernelLayer.py:119; gradients are routed through the SAVED arg-max permutation.
//
// Formulation (tile.cuh: tiles of <= 128 nodes holding whole molecules, kernel rows in blocks of <= 128).  For one
// (tile, block) the sparse coefficient matrix
//        Wt[row, v] = sum over (n, j, k): nei(n,j) = v, row = (k, pi*_{n,k}(j))   of   alpha_d * chi*g[n,k] / scale
//                   (+ centre rows: Wt[(c,k), n] = beta_d * chi*g[n,k] / scale)
// is built in shared memory as an fp16 (hi, lo) tensor-core operand, and two GEMMs consume it:
//        G_b[row, f] += Wt[row, :] . xhat[:, f]        kernel-parameter gradients: TMEM resident over ALL tiles of the CTA
//        dxh[v, f]   += Wt[:, v]^T . khat_b[:, f]      gradient w.r.t. the normalised input rows: TMEM resident over the
//                                                      blocks of one tile
// g[n,k] is the incoming gradient (optionally summed over the node's neighbours = transpose of propagate,
// KernelLayer.py:119), chi the saved chirality sign, alpha_d = w_s/(d W), beta_d = w_c/W the softmax mixing factors
// (kernels.py:402-425) and `scale` a power of two derived from max|grad| so that every entry fits fp16 comfortably; both
// operands are unscaled (hi, lo) splits, three UMMAs per K step, fp32 accumulation.
//
//   k_coef_bond      streaming pre-pass in bucket order: coef = chi * g per (node, kernel) pair (the gather of the
//                    incoming gradient rows stays out of the tile kernel's critical path) and the bond-attribute support
//                    gradients (8 floats per support row) as one partial copy per CTA
//   k_conv_bwd_tile  tile-major persistent kernel over a subset of <= 2 kernel blocks (TMEM: 2 x G_b + dxh = 336 of the
//                    512 columns; a layer with 4 blocks runs it twice and the first launch hands its partial dxh to the
//                    second through `scratch`).  Per (tile, block): Wt by scatter with one thread per (node, kernel) pair;
//                    neighbour slots are pre-grouped by collision rank (k_tile_meta, bucket.cu) so that rank 0 is a plain
//                    store and the few higher ranks read-modify-write after a barrier -- deterministic, no atomics;
//                    then 48 UMMAs.  The block's images stream through shared memory by bulk copy.  Per tile one dxh
//                    epilogue; the last launch applies d(x/|x|)/dx and writes grad_x.
// k_param_finalize (params.cu) reduces the per-CTA copies in fixed order.
#include <algorithm>
#include "common.cuh"
#include "tc.cuh"
#include "tile.cuh"

namespace mk {

bool tile_layer_ok(const molkgnn_layer_t* layer);

constexpr int TB_THREADS = 512;
constexpr int WT_ONE = 16 * 16 * 128;          // one fp16 image of the 128 x 128 coefficient block
constexpr int CB_TN = 64;                      // nodes per step of k_coef_bond
constexpr int CB_THREADS = 256;
constexpr int CB_MAXR = 4;                     // support rows per thread of k_coef_bond (L * d <= 1024)

// =============================================================================================================
// k_coef_bond
// =============================================================================================================
struct CoefArgs {
    const int* sel; const int* nei; const float* ehat;
    int n[4], boff[4], eoff[4], L[4], koff[4];
    long long scoff[4];
    const float* grad; int ldg; int grad_mode;
    const uint8_t* argmax;
    float* coef;
    float* partials; long long part_off[4]; int FW, Fp;
    int G;                                     // CTAs (= partial copies) per degree
    float* amax;                               // device scalar (zeroed by the launcher): max |coef|
};

template <int D>
__device__ __forceinline__ void coef_bond_body(const CoefArgs& a, float* coefS, unsigned char* invS, float4* ehS, int c) {
    const int L = a.L[D - 1], n = a.n[D - 1];
    const int tid = threadIdx.x;
    const int eoff = a.eoff[D - 1], koff = a.koff[D - 1];
    const int nrows = L * D;
    float acc[CB_MAXR][EP];
#pragma unroll
    for (int r = 0; r < CB_MAXR; ++r)
#pragma unroll
        for (int e = 0; e < EP; ++e) acc[r][e] = 0.f;
    float amax = 0.f;
    const int ntiles = (n + CB_TN - 1) / CB_TN;
    for (int t = c; t < ntiles; t += a.G) {
        const int R0 = t * CB_TN;
        const int nv = min(CB_TN, n - R0);
        __syncthreads();
        // bond rows of the node tile: contiguous in the plan, staged once
        {
            const float4* src = reinterpret_cast<const float4*>(a.ehat + ((size_t)eoff + (size_t)R0 * D) * EP);
            for (int i = tid; i < nv * D * 2; i += CB_THREADS) ehS[i] = __ldg(src + i);
        }
        // coefficients: 4 pairs per thread in flight
        constexpr int U = 4;
        const int npair = nv * L;
        for (int i0 = tid; i0 < npair; i0 += CB_THREADS * U) {
            float g[U];
            uint8_t am[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * CB_THREADS;
                g[u] = 0.f; am[u] = 0;
                if (i < npair) {
                    const int nl = i / L, k = i - nl * L;
                    const int col = koff + k;
                    const int R = R0 + nl;
                    if (a.grad_mode == 0) {
                        g[u] = a.grad[(size_t)a.sel[a.boff[D - 1] + R] * a.ldg + col];
                    } else {
                        const int* nb = a.nei + (size_t)eoff + (size_t)R * D;
                        float s = a.grad[(size_t)nb[0] * a.ldg + col];
#pragma unroll
                        for (int j = 1; j < D; ++j) s += a.grad[(size_t)nb[j] * a.ldg + col];
                        g[u] = s;
                    }
                    am[u] = a.argmax[(size_t)a.scoff[D - 1] + (size_t)R0 * L + i];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * CB_THREADS;
                if (i < npair) {
                    const float av = (am[u] & 0x80) ? -g[u] : g[u];
                    amax = fmaxf(amax, fabsf(av));
                    a.coef[(size_t)a.scoff[D - 1] + (size_t)R0 * L + i] = av;
                    coefS[i] = av;
                    uint32_t inv = 0;
#pragma unroll
                    for (int p = 0; p < Perm<D>::P; ++p) if (p == (am[u] & 0x7f)) inv = perm_inv_code<D>(p);
                    invS[i] = (unsigned char)inv;
                }
            }
        }
        __syncthreads();
        // bond-attribute support gradients: thread owns rows (s, k), sums over the nodes in fixed order
#pragma unroll
        for (int r = 0; r < CB_MAXR; ++r) {
            const int row = tid + r * CB_THREADS;
            if (row < nrows) {
                const int s = row / L, k = row - s * L;
#pragma unroll 4
                for (int nl = 0; nl < nv; ++nl) {
                    const float av = coefS[nl * L + k];
                    const int j = (invS[nl * L + k] >> (2 * s)) & 3;
                    const float4 e0 = ehS[(nl * D + j) * 2], e1 = ehS[(nl * D + j) * 2 + 1];
                    acc[r][0] = fmaf(av, e0.x, acc[r][0]); acc[r][1] = fmaf(av, e0.y, acc[r][1]);
                    acc[r][2] = fmaf(av, e0.z, acc[r][2]); acc[r][3] = fmaf(av, e0.w, acc[r][3]);
                    acc[r][4] = fmaf(av, e1.x, acc[r][4]); acc[r][5] = fmaf(av, e1.y, acc[r][5]);
                    acc[r][6] = fmaf(av, e1.z, acc[r][6]); acc[r][7] = fmaf(av, e1.w, acc[r][7]);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((tid & 31) == 0 && amax > 0.f) atomicMax(reinterpret_cast<unsigned int*>(a.amax), __float_as_uint(amax));
    const int rows_x = (D + 1) * L;
    float* part = a.partials + a.part_off[D - 1] + (size_t)c * rows_x * a.FW + a.Fp;
#pragma unroll
    for (int r = 0; r < CB_MAXR; ++r) {
        const int row = tid + r * CB_THREADS;
        if (row < nrows) {
            st4(part + (size_t)row * a.FW, make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]));
            st4(part + (size_t)row * a.FW + 4, make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]));
        }
    }
    for (int row = nrows + tid; row < rows_x; row += CB_THREADS) {   // centre rows carry no bond part
        st4(part + (size_t)row * a.FW, make_float4(0.f, 0.f, 0.f, 0.f));
        st4(part + (size_t)row * a.FW + 4, make_float4(0.f, 0.f, 0.f, 0.f));
    }
}

__global__ void __launch_bounds__(CB_THREADS) k_coef_bond(const __grid_constant__ CoefArgs a) {
    extern __shared__ __align__(16) unsigned char smem_c[];
    const int d = blockIdx.x / a.G + 1, c = blockIdx.x % a.G;
    const int L = a.L[d - 1];
    if (L == 0) return;
    float4* ehS = reinterpret_cast<float4*>(smem_c);                     // [CB_TN * 4 slots][2]
    float* coefS = reinterpret_cast<float*>(smem_c + CB_TN * 4 * 32);
    unsigned char* invS = smem_c + CB_TN * 4 * 32 + (size_t)CB_TN * L * 4;
    switch (d) {
        case 1: coef_bond_body<1>(a, coefS, invS, ehS, c); break;
        case 2: coef_bond_body<2>(a, coefS, invS, ehS, c); break;
        case 3: coef_bond_body<3>(a, coefS, invS, ehS, c); break;
        default: coef_bond_body<4>(a, coefS, invS, ehS, c); break;
    }
}

// =============================================================================================================
// k_conv_bwd_tile
// =============================================================================================================
struct BwdTileArgs {
    const float* xnorm;
    int F, Fp, Fk;
    const TileMetaG* meta; const unsigned char* ximg; int n_tiles;
    int L[4];
    const float* packed[4];
    const unsigned char* img;
    TileBlocks tb;
    int img_one, x_one;
    const float* coef;                 // chi * g per (node, kernel) pair, compact bucket order (k_coef_bond)
    const uint8_t* argmax; long long scoff[4];
    const float* amax;                 // device scalar: max |coef| (k_coef_bond)
    float* partials; long long part_off[4]; int FW;
    float* scratch;                    // [N, Fk] partial dxh handed from the first launch to the second
    float* gx; int ldgx;
    int nbl, blist[2];                 // kernel blocks of this launch
    int first, last;                   // first: no partial dxh to add; last: apply the Jacobian and write grad_x
    int a_cap, buf_bytes;
    int sm_img, sm_x, sm_wt, sm_buf, sm_a, sm_am;
};

struct BSeg {
    float alpha, beta;                 // w_s/(d W), w_c/W
    int d, k0, nk, rowbase, L;
    float rnk;
};

__device__ __forceinline__ void wt_store(unsigned char* wt, int row, int col, float v) {
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    const uint32_t off = tc::il_off(row, col, 128);
    *reinterpret_cast<__half*>(wt + off) = hi;
    *reinterpret_cast<__half*>(wt + WT_ONE + off) = lo;
}
__device__ __forceinline__ void wt_add(unsigned char* wt, int row, int col, float v) {
    const uint32_t off = tc::il_off(row, col, 128);
    __half* ph = reinterpret_cast<__half*>(wt + off);
    __half* pl = reinterpret_cast<__half*>(wt + WT_ONE + off);
    v += __half2float(*ph) + __half2float(*pl);
    const __half hi = __float2half_rn(v);
    *ph = hi;
    *pl = __float2half_rn(v - __half2float(hi));
}

// thread 0: metadata record and node images of `tile`
__device__ __forceinline__ void tb_issue_copy(const BwdTileArgs& a, unsigned char* smem, unsigned char* buf, int tile,
                                              uint64_t* bar) {
    mbar_expect_tx(bar, (uint32_t)sizeof(TileMetaG) + 2u * (uint32_t)a.x_one);
    bulk_g2s(buf, a.meta + tile, (uint32_t)sizeof(TileMetaG), bar);
    bulk_g2s(smem + a.sm_x, a.ximg + (size_t)tile * 2 * a.x_one, 2u * (uint32_t)a.x_one, bar);
}
// thread 0: images of kernel block `blk`
__device__ __forceinline__ void tb_issue_img(const BwdTileArgs& a, unsigned char* smem, int blk, uint64_t* bar) {
    mbar_expect_tx(bar, 2u * (uint32_t)a.img_one);
    bulk_g2s(smem + a.sm_img, a.img + (size_t)blk * 2 * a.img_one, 2u * (uint32_t)a.img_one, bar);
}

// thread 0: G_bi (TMEM columns bi*128 ..) += Wt . xhat ;  dxh (TMEM columns 256 ..) (+)= Wt^T . khat
__device__ __forceinline__ void tb_issue_mma(const BwdTileArgs& a, unsigned char* smem, int nn, int rows, int bi,
                                             bool g_fresh, uint32_t tmem, uint64_t* bar) {
    const uint32_t whi = tc::smem_u32(smem + a.sm_wt), wlo = whi + WT_ONE;
    const uint32_t xhi = tc::smem_u32(smem + a.sm_x), xlo = xhi + (uint32_t)a.x_one;
    const uint32_t ihi = tc::smem_u32(smem + a.sm_img), ilo = ihi + (uint32_t)a.img_one;
    const uint32_t fgrp = (uint32_t)(a.Fk >> 3) * 128u;      // bytes of one 8-row group of an [R x Fk] image
    const uint32_t idesc_g = tc::idesc_f16(128, a.Fk, 0, 1);  // A = Wt K-major (K = node), B = xhat MN-major (N = feature)
    const uint32_t idesc_x = tc::idesc_f16(128, a.Fk, 1, 1);  // A = Wt MN-major (M = node, K = row), B = khat MN-major
    const uint32_t dG = tmem + (uint32_t)bi * 128u, dX = tmem + 256u;
    // G: K = nodes, 16 per step.  Wt K-major: +256 B per step (two 8-column chunks); xhat MN-major: +2 row groups per step
    const int nkg = (max(16, (nn + 15) & ~15)) >> 4;
    for (int ks = 0; ks < nkg; ++ks) {
        const uint64_t dAh = tc::smem_desc(whi + ks * 256u, 128u, 2048u), dAl = tc::smem_desc(wlo + ks * 256u, 128u, 2048u);
        const uint64_t dBh = tc::smem_desc(xhi + ks * 2u * fgrp, fgrp, 128u), dBl = tc::smem_desc(xlo + ks * 2u * fgrp, fgrp, 128u);
        tc::umma_f16(dG, dAh, dBh, idesc_g, (g_fresh && ks == 0) ? 0u : 1u);
        tc::umma_f16(dG, dAl, dBh, idesc_g, 1u);
        tc::umma_f16(dG, dAh, dBl, idesc_g, 1u);
    }
    // dxh: K = kernel rows, 16 per step.  Wt MN-major: +2 row groups (2 * 2048 B) per step; khat MN-major likewise
    const int nkx = (max(16, (rows + 15) & ~15)) >> 4;
    for (int ks = 0; ks < nkx; ++ks) {
        const uint64_t dAh = tc::smem_desc(whi + ks * 4096u, 2048u, 128u), dAl = tc::smem_desc(wlo + ks * 4096u, 2048u, 128u);
        const uint64_t dBh = tc::smem_desc(ihi + ks * 2u * fgrp, fgrp, 128u), dBl = tc::smem_desc(ilo + ks * 2u * fgrp, fgrp, 128u);
        tc::umma_f16(dX, dAh, dBh, idesc_x, (bi == 0 && ks == 0) ? 0u : 1u);
        tc::umma_f16(dX, dAl, dBh, idesc_x, 1u);
        tc::umma_f16(dX, dAh, dBl, idesc_x, 1u);
    }
    tc::umma_commit(bar);
}

#ifdef MK_PHASE_CLOCKS
__device__ unsigned long long g_ph_bwd[16];
#endif

__global__ void __launch_bounds__(TB_THREADS, 1) k_conv_bwd_tile(const __grid_constant__ BwdTileArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar_mma, bar_img, bar_cp[2];
    __shared__ uint32_t tslot;
    __shared__ BSeg s_seg[2][TILE_MAXSEG];
    __shared__ unsigned char s_lut[4][12];               // packed permutation codes (2 bits per j) per degree
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    MK_PH_DECL(tid == 0)
    if (tid == 0) {
        tc::mbar_init(&bar_mma, 1); tc::mbar_init(&bar_img, 1);
        tc::mbar_init(&bar_cp[0], 1); tc::mbar_init(&bar_cp[1], 1);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(&tslot, 512);
    if (tid < 48) {
        const int d = tid / 12 + 1, p = tid % 12;
        uint32_t code = 0;
        if (d == 2) code = p < 2 ? perm_code<2>(p) : 0;
        else if (d == 3) { for (int q = 0; q < 6; ++q) if (q == p) code = perm_code<3>(q); }
        else if (d == 4) { for (int q = 0; q < 12; ++q) if (q == p) code = perm_code<4>(q); }
        s_lut[d - 1][p] = (unsigned char)code;
    }
    if (tid >= 64 && tid < 64 + 2 * TILE_MAXSEG) {
        const int bi = (tid - 64) / TILE_MAXSEG, si = (tid - 64) % TILE_MAXSEG;
        if (bi < a.nbl && si < a.tb.nseg[a.blist[bi]]) {
            const TileSeg sg = a.tb.seg[a.blist[bi]][si];
            const int L = a.L[sg.d - 1];
            const PackedLayout pl(sg.d, L, a.Fp);
            const float* pk = a.packed[sg.d - 1];
            BSeg c;
            c.alpha = pk[pl.w + 0] / pk[pl.w + 3] / (float)sg.d;
            c.beta = pk[pl.w + 1] / pk[pl.w + 3];
            c.d = sg.d; c.k0 = sg.k0; c.nk = sg.nk; c.rowbase = sg.rowbase; c.L = L;
            c.rnk = 1.0f / (float)sg.nk;
            s_seg[bi][si] = c;
        }
    }
    unsigned char* wt = smem + a.sm_wt;
    for (int i = tid * 16; i < 2 * WT_ONE; i += TB_THREADS * 16) *reinterpret_cast<uint4*>(wt + i) = make_uint4(0, 0, 0, 0);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tslot;
    float* a_s = reinterpret_cast<float*>(smem + a.sm_a);
    unsigned char* am_s = smem + a.sm_am;
    float* red = reinterpret_cast<float*>(smem + a.sm_a);       // Jacobian reduction scratch: the coefficients are dead by then
    uint32_t ph_mma = 0u, ph_img = 0u, ph_cp[2] = {0u, 0u};
    // power-of-two scale: |alpha * chi * g| / scale <= 2^10
    float scale, rscale;
    {
        const float gm = fmaxf(*a.amax, 1e-30f);
        int e;
        frexpf(gm, &e);                                   // gm < 2^e
        scale = ldexpf(1.0f, e - 10);
        rscale = ldexpf(1.0f, 10 - e);
    }
    const int q = warp & 3, cpart = warp >> 2;            // TMEM lane quadrant / 32-column part of this warp
    bool img_pending = false;                             // an image copy is in flight (uniform across the CTA)
    if ((int)blockIdx.x < a.n_tiles) {
        if (tid == 0) {
            tb_issue_copy(a, smem, smem + a.sm_buf, blockIdx.x, &bar_cp[0]);
            tb_issue_img(a, smem, a.blist[0], &bar_img);
        }
        img_pending = true;
    }
    MK_PH(0);                                             // prologue
    int cur = 0;
    bool fresh = true;                                    // first tile of this CTA: the G accumulators start from zero

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        unsigned char* buf = smem + a.sm_buf + cur * a.buf_bytes;
        const TileMetaG& m = *reinterpret_cast<const TileMetaG*>(buf);
        tc::mbar_wait(&bar_cp[cur], ph_cp[cur]);
        ph_cp[cur] ^= 1u;
        MK_PH(1);                                         // wait for the tile's metadata + node images
        const int t0 = m.t0, nn = m.nn;
        const int tnext = tile + gridDim.x;
        for (int bi = 0; bi < a.nbl; ++bi) {
            const int blk = a.blist[bi];
            const int nseg = a.tb.nseg[blk];
            int abase[TILE_MAXSEG];
            {
                int run = 0;
                for (int si = 0; si < nseg; ++si) { abase[si] = run; run += m.cnt[s_seg[bi][si].d - 1] * s_seg[bi][si].nk; }
            }
            // ---- rank 0: one thread per (node, kernel) pair -- coefficient, centre entry, collision-free support entries ----
            for (int si = 0; si < nseg; ++si) {
                const BSeg sg = s_seg[bi][si];
                const int np = m.cnt[sg.d - 1] * sg.nk;
                for (int p = tid; p < np; p += TB_THREADS) {
                    const int ni = (int)(((float)p + 0.5f) * sg.rnk);
                    const int kl = p - ni * sg.nk;
                    const int nl_ = m.list[sg.d - 1][ni];
                    const size_t cidx = (size_t)a.scoff[sg.d - 1] + (size_t)m.posl[nl_] * sg.L + sg.k0 + kl;
                    const float av = __ldg(a.coef + cidx) * rscale;
                    const int am = a.argmax[cidx] & 0x7f;
                    a_s[abase[si] + p] = av;
                    am_s[abase[si] + p] = (unsigned char)am;
                    const uint32_t code = s_lut[sg.d - 1][am];
                    const uint32_t nw = m.nl[nl_];
                    const uint32_t cr = m.cr[nl_];
                    wt_store(wt, sg.rowbase + sg.d * sg.nk + kl, nl_, av * sg.beta);
                    const float as = av * sg.alpha;
                    for (int j = 0; j < sg.d; ++j) {
                        if (((cr >> (2 * j)) & 3u) == 0u)
                            wt_store(wt, sg.rowbase + (int)((code >> (2 * j)) & 3u) * sg.nk + kl, (int)((nw >> (8 * j)) & 0xffu), as);
                    }
                }
            }
            MK_PH(2);                                     // rank-0 scatter (thread 0's own share)
            // ---- ranks 1..3: one thread per (neighbour slot, kernel), read-modify-write after a barrier ----
            for (int r = 1; r < 4; ++r) {
                bool more = false;
                for (int si = 0; si < nseg; ++si) {
                    const int d = s_seg[bi][si].d;
                    if (m.eoffs[d - 1][4] > m.eoffs[d - 1][r]) more = true;
                }
                if (!more) break;
                __syncthreads();
                for (int si = 0; si < nseg; ++si) {
                    const BSeg sg = s_seg[bi][si];
                    const int e0 = m.eoffs[sg.d - 1][r], e1 = m.eoffs[sg.d - 1][r + 1];
                    const int ni_ = (e1 - e0) * sg.nk;
                    for (int p = tid; p < ni_; p += TB_THREADS) {
                        const int ei = (int)(((float)p + 0.5f) * sg.rnk);
                        const int kl = p - ei * sg.nk;
                        const int ent = m.elist[e0 + ei];
                        const int nl_ = ent >> 2, j = ent & 3;
                        const int pi = abase[si] + m.lidx[nl_] * sg.nk + kl;
                        const int s = (s_lut[sg.d - 1][am_s[pi]] >> (2 * j)) & 3;
                        wt_add(wt, sg.rowbase + s * sg.nk + kl, (int)((m.nl[nl_] >> (8 * j)) & 0xffu), a_s[pi] * sg.alpha);
                    }
                }
            }
            MK_PH(3);                                     // ranks 1..3 (incl. waiting for the slowest rank-0 thread)
            tc::fence_async_smem();
            __syncthreads();
            MK_PH(4);                                     // barrier before the MMAs
            // ---- tensor cores ----
            if (tid == 0) {
                if (img_pending) {                            // this block's images were requested after the previous MMAs
                    tc::mbar_wait(&bar_img, ph_img);
                    ph_img ^= 1u;
                }
                tc::fence_after_sync();
                tb_issue_mma(a, smem, nn, a.tb.rows[blk], bi, fresh, tmem, &bar_mma);
            }
            img_pending = false;
            MK_PH(5);                                     // image wait + MMA issue
            tc::mbar_wait(&bar_mma, ph_mma);
            ph_mma ^= 1u;
            tc::fence_after_sync();
            MK_PH(6);                                     // MMA completion
            // the tensor cores are done with Wt and the images: clear Wt, fetch the next block's images
            for (int i = tid * 16; i < 2 * WT_ONE; i += TB_THREADS * 16) *reinterpret_cast<uint4*>(wt + i) = make_uint4(0, 0, 0, 0);
            {
                const int nblk = bi + 1 < a.nbl ? a.blist[bi + 1] : (tnext < a.n_tiles ? a.blist[0] : blk);
                if (nblk != blk) {
                    if (tid == 0) tb_issue_img(a, smem, nblk, &bar_img);
                    img_pending = true;
                }
            }
            if (bi + 1 < a.nbl) __syncthreads();              // Wt cleared before the next block's scatter
            MK_PH(7);                                     // Wt clear
        }
        fresh = false;
        // ---- dxh epilogue: lane = node, 32 columns per warp ----
        {
            const int v = q * 32 + lane;
            const int f0 = cpart * 32;
            const bool colok = f0 < a.Fk;
            const bool rowok = v < nn;
            const int nf = min(32, a.Fk - f0);            // columns of this part (multiple of 16, or <= 0)
            float dv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) dv[i] = 0.f;
            float* sp = a.scratch + (size_t)(t0 + v) * a.Fk + f0;
            if (!a.first && rowok && colok) {             // partial dxh of the first launch
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    if (i < nf) {
                        const float4 o = __ldcg(reinterpret_cast<const float4*>(sp + i));
                        dv[i] = o.x; dv[i + 1] = o.y; dv[i + 2] = o.z; dv[i + 3] = o.w;
                    }
                }
            }
            float nrm = 1.f;
            if (a.last && rowok) nrm = a.xnorm[t0 + v];
            if (colok) {
                uint32_t u[32];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + 256u + (uint32_t)f0;
                tc::tmem_ld16(taddr, u);
                if (f0 + 16 < a.Fk) tc::tmem_ld16(taddr + 16, u + 16);
                else {
#pragma unroll
                    for (int i = 16; i < 32; ++i) u[i] = 0u;
                }
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) { asm volatile("" : "+r"(u[i])); dv[i] = fmaf(__uint_as_float(u[i]), scale, dv[i]); }
            }
            if (!a.last) {
                if (tid == 0 && tnext < a.n_tiles)
                    tb_issue_copy(a, smem, smem + a.sm_buf + (cur ^ 1) * a.buf_bytes, tnext, &bar_cp[cur ^ 1]);
                if (rowok && colok) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        if (i < nf) __stcg(reinterpret_cast<float4*>(sp + i), make_float4(dv[i], dv[i + 1], dv[i + 2], dv[i + 3]));
                }
            } else {
                // chain rule through xhat = x / max(|x|, eps):  gx = (g - (xhat . g) xhat) / |x|;  xhat = hi + lo from the
                // node images still resident in shared memory
                float xh[32];
                float dot = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) xh[i] = 0.f;
                if (colok) {
                    const unsigned char* Xhi = smem + a.sm_x;
                    const unsigned char* Xlo = Xhi + a.x_one;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (8 * c < nf) {
                            const uint32_t off = tc::il_off(v, f0 + 8 * c, a.Fk);
                            const uint4 h4 = *reinterpret_cast<const uint4*>(Xhi + off);
                            const uint4 l4 = *reinterpret_cast<const uint4*>(Xlo + off);
                            const __half2* hh = reinterpret_cast<const __half2*>(&h4);
                            const __half2* ll = reinterpret_cast<const __half2*>(&l4);
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const float2 fh = __half22float2(hh[t]), fl = __half22float2(ll[t]);
                                xh[8 * c + 2 * t] = fh.x + fl.x;
                                xh[8 * c + 2 * t + 1] = fh.y + fl.y;
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) dot = fmaf(dv[i], xh[i], dot);
                }
                red[cpart * 128 + v] = dot;
                __syncthreads();
                if (tid == 0 && tnext < a.n_tiles)      // every thread has read its xhat: the image buffer may be refilled
                    tb_issue_copy(a, smem, smem + a.sm_buf + (cur ^ 1) * a.buf_bytes, tnext, &bar_cp[cur ^ 1]);
                dot = (red[v] + red[128 + v]) + (red[256 + v] + red[384 + v]);
                const float den = fmaxf(nrm, MOLKGNN_COS_EPS);
                const float rden = 1.0f / den;
                const bool clamped = !(nrm > MOLKGNN_COS_EPS);
                if (a.gx && rowok && colok) {
                    float* out = a.gx + (size_t)(t0 + v) * a.ldgx + f0;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        if (f0 + i + 4 <= a.Fp) {
                            float o[4];
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                o[c] = clamped ? dv[i + c] * rden : (dv[i + c] - dot * xh[i + c]) * rden;
                                if (f0 + i + c >= a.F) o[c] = 0.f;
                            }
                            st4(out + i, make_float4(o[0], o[1], o[2], o[3]));
                        }
                    }
                }
            }
        }
        tc::fence_before_sync();
        __syncthreads();                 // coefficient arrays / Jacobian scratch, TMEM dxh and this tile's buffer are free again
        MK_PH(8);                                         // dxh epilogue
        cur ^= 1;
    }
    // ---- kernel-parameter partial sums of this CTA: node-attribute part of every row of this launch's blocks ----
    for (int bi = 0; bi < a.nbl; ++bi) {
        const int blk = a.blist[bi];
        const int row = q * 32 + lane;
        int d = 0, slot = 0, kk = 0, L = 0;
        for (int si = 0; si < a.tb.nseg[blk]; ++si) {
            const TileSeg sg = a.tb.seg[blk][si];
            const int r = row - sg.rowbase;
            if (r >= 0 && r < sg.nk * (sg.d + 1)) { d = sg.d; slot = r / sg.nk; kk = sg.k0 + r % sg.nk; L = a.L[sg.d - 1]; }
        }
        const int f0 = cpart * 32;
        if (f0 < a.Fk) {
            uint32_t u[32];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(bi * 128 + f0);
            tc::tmem_ld16(taddr, u);
            if (f0 + 16 < a.Fk) tc::tmem_ld16(taddr + 16, u + 16);
            else {
#pragma unroll
                for (int i = 16; i < 32; ++i) u[i] = 0u;
            }
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(u[i]));
            if (d > 0) {
                const int rows_x = (d + 1) * L;
                float* part = a.partials + a.part_off[d - 1] + ((size_t)blockIdx.x * rows_x + (size_t)slot * L + kk) * a.FW;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    if (f0 + i + 4 <= a.Fp)
                        st4(part + f0 + i, make_float4(__uint_as_float(u[i]) * scale, __uint_as_float(u[i + 1]) * scale,
                                                       __uint_as_float(u[i + 2]) * scale, __uint_as_float(u[i + 3]) * scale));
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    MK_PH(9);                                             // G partial sums -> global
    MK_PH_FLUSH(g_ph_bwd);
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ---- host side -----------------------------------------------------------------------------------------------------
static bool tile_plan_ok_b(const molkgnn_plan_t* plan) {
    return plan->n_tiles > 0 && plan->tile_start && plan->tile_meta && plan->ehat_node && plan->tile_max_nodes <= TNODES;
}

int tile_bwd_grid(const molkgnn_plan_t* plan) {
    const int sms = device_num_sms();
    return std::max(1, std::min(plan->n_tiles, sms));
}

bool tile_bwd_ok(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer) {
    if (!tile_plan_ok_b(plan) || !layer->tile_img || !tile_layer_ok(layer)) return false;
    TileBlocks tb;
    if (!tb.build(layer->L) || tb.nb > 4) return false;                 // at most two launches of two blocks
    for (int d = 0; d < 4; ++d) if (layer->L[d] * (d + 1) > CB_MAXR * CB_THREADS) return false;
    return true;
}

// returns 1 if launched, 0 if not eligible, <0 on error.  part_off / ncta describe the partial copies for k_param_finalize.
int launch_conv_bwd_tile(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                         const float* xnorm, const void* ximg, const float* grad, int32_t ldg, int32_t grad_mode,
                         const uint8_t* argmax, const int64_t scoff[4], float* coef, float* partials, float* scratch,
                         float* grad_x, int32_t ldgx, int64_t part_off[4], int ncta[4], int64_t* part_total, bool do_launch,
                         cudaStream_t st) {
    (void)x; (void)ldx;
    if (!ximg || !coef || !tile_bwd_ok(plan, layer)) return 0;
    static int s_budget = 0;
    if (!s_budget) {
        s_budget = device_max_smem_optin();
        MK_REQUIRE(s_budget > 0, "conv_bwd_tile: no CUDA device");
    }
    BwdTileArgs a;
    if (!a.tb.build(layer->L)) return 0;
    const int nlaunch = a.tb.nb > 2 ? 2 : 1;
    MK_REQUIRE(nlaunch == 1 || scratch, "conv_bwd_tile: scratch is required for layers with more than two kernel blocks");
    a.xnorm = xnorm;
    a.F = layer->F; a.Fp = layer->Fp; a.Fk = tile_fk(layer->Fp);
    a.meta = reinterpret_cast<const TileMetaG*>(plan->tile_meta);
    a.ximg = reinterpret_cast<const unsigned char*>(ximg);
    a.n_tiles = plan->n_tiles;
    const int grid = tile_bwd_grid(plan);
    a.FW = layer->Fp + EP;
    int64_t po = 0, rows_all = 0;
    for (int d = 0; d < 4; ++d) {
        a.L[d] = layer->L[d];
        a.packed[d] = layer->packed[d];
        a.scoff[d] = scoff[d];
        a.part_off[d] = part_off[d] = po;
        ncta[d] = layer->L[d] > 0 ? grid : 0;
        po += (int64_t)ncta[d] * (d + 2) * layer->L[d] * a.FW;
        rows_all += (int64_t)(d + 2) * layer->L[d];
    }
    *part_total = po;
    float* amax = partials + po + 2 * rows_all + 8;   // spare floats behind k_param_finalize's Q scratch
    a.img = reinterpret_cast<const unsigned char*>(layer->tile_img);
    a.img_one = tile_img_one(a.Fk);
    a.x_one = tile_img_one(a.Fk);
    a.coef = coef;
    a.argmax = argmax;
    a.amax = amax;
    a.partials = partials;
    a.scratch = scratch;
    a.gx = grad_x; a.ldgx = ldgx;
    // capacity from the plan: (node, kernel) pairs of the fullest tile of any block
    int a_cap = 0;
    for (int b = 0; b < a.tb.nb; ++b) {
        int c = 0;
        for (int si = 0; si < a.tb.nseg[b]; ++si) c += plan->tile_max_deg[a.tb.seg[b][si].d - 1] * a.tb.seg[b][si].nk;
        a_cap = std::max(a_cap, c);
    }
    a.a_cap = a_cap;
    a.buf_bytes = (int)((sizeof(TileMetaG) + 127) / 128 * 128);
    int64_t off = 0;
    a.sm_img = (int)off; off += 2 * (int64_t)a.img_one;
    a.sm_x = (int)off; off += 2 * (int64_t)a.x_one;
    a.sm_wt = (int)off; off += 2 * (int64_t)WT_ONE;
    a.sm_buf = (int)off; off += 2 * (int64_t)a.buf_bytes;
    a.sm_a = (int)off; off += (std::max<int64_t>((int64_t)a_cap * 4, 4 * 128 * 4) + 127) / 128 * 128;
    a.sm_am = (int)off; off += ((int64_t)a_cap + 127) / 128 * 128;
    if (off > s_budget - 2048) return 0;
    int Lmax = 1;
    for (int d = 0; d < 4; ++d) Lmax = std::max(Lmax, layer->L[d]);
    const int64_t smem_c = (int64_t)CB_TN * 4 * 32 + (int64_t)CB_TN * Lmax * 5;
    if (smem_c > s_budget - 2048) return 0;
    if (!do_launch) return 1;
    static int64_t s_attr = 0, s_attr_c = 0;
    if (off > s_attr) {
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_conv_bwd_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off));
        s_attr = off;
    }
    if (smem_c > s_attr_c) {
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_coef_bond, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        s_attr_c = smem_c;
    }
    MK_CHECK_CUDA(cudaMemsetAsync(amax, 0, sizeof(float), st));
    {
        CoefArgs c;
        c.sel = plan->sel; c.nei = plan->nei; c.ehat = plan->ehat;
        for (int d = 0; d < 4; ++d) {
            c.n[d] = plan->n[d]; c.boff[d] = plan->boff[d]; c.eoff[d] = plan->eoff[d];
            c.L[d] = layer->L[d]; c.koff[d] = layer->koff[d]; c.scoff[d] = scoff[d];
            c.part_off[d] = part_off[d];
        }
        c.grad = grad; c.ldg = ldg; c.grad_mode = grad_mode;
        c.argmax = argmax; c.coef = coef;
        c.partials = partials; c.FW = a.FW; c.Fp = layer->Fp;
        c.G = grid;
        c.amax = amax;
        count_launches(1);
        ProfScope prof("coef_bond", st);
        k_coef_bond<<<4 * grid, CB_THREADS, smem_c, st>>>(c);
        MK_CHECK_CUDA(cudaGetLastError());
    }
    for (int l = 0; l < nlaunch; ++l) {
        // blocks {0, 3} and {1, 2}: the scatter work of the two launches is about equal for the base model
        if (a.tb.nb <= 2) { a.nbl = a.tb.nb; a.blist[0] = 0; a.blist[1] = a.tb.nb > 1 ? 1 : 0; }
        else if (a.tb.nb == 3) { a.nbl = l == 0 ? 2 : 1; a.blist[0] = l == 0 ? 0 : 2; a.blist[1] = l == 0 ? 1 : 2; }
        else { a.nbl = 2; a.blist[0] = l == 0 ? 0 : 1; a.blist[1] = l == 0 ? 3 : 2; }
        a.first = l == 0; a.last = l == nlaunch - 1;
        count_launches(1);
        ProfScope prof("conv_bwd_tile", st);
        k_conv_bwd_tile<<<grid, TB_THREADS, off, st>>>(a);
        MK_CHECK_CUDA(cudaGetLastError());
    }
    return 1;
}

}  // namespace mk


#ifdef MK_PHASE_CLOCKS
// profiling build only: read (and clear) the accumulated phase clocks of k_conv_bwd_tile
extern "C" int molkgnn_debug_phase_clocks_bwd(unsigned long long* out16) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out16, mk::g_ph_bwd, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
    unsigned long long z[16] = {0};
    return cudaMemcpyToSymbol(mk::g_ph_bwd, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif
