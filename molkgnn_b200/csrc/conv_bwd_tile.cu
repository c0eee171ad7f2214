// Molecule-tile backward of the molecular-kernel convolution (sm_100a, tcgen05 + TMEM).  The reference relies on autograd
// over kernels.py:353-425 and KernelLayer.py:119; gradients are routed through the SAVED arg-max permutation.
//
// Formulation (tile.cuh: tiles of <= 128 nodes holding whole molecules, kernel rows in blocks of <= 128).  For one
// (block, tile) the sparse coefficient matrix
//        Wt[row, v] = sum over (n, j, k): nei(n,j) = v, row = (k, pi*_{n,k}(j))   of   alpha_d * chi*g[n,k] / scale
//                   (+ centre rows: Wt[(c,k), n] = beta_d * chi*g[n,k] / scale)
// is built in shared memory as an fp16 (hi, lo) tensor-core operand, and two GEMMs consume it:
//        G[row, f]   += Wt[row, :] . xhat[:, f]        kernel-parameter gradients, accumulated in TMEM over all tiles of the CTA
//        dxh[v, f]    = Wt[:, v]^T . khat[:, f]        gradient w.r.t. the normalised input rows, per tile
// g[n,k] is the incoming gradient (optionally summed over the node's neighbours = transpose of propagate, KernelLayer.py:119),
// chi the saved chirality sign, alpha_d = w_s/(d W), beta_d = w_c/W the softmax mixing factors (kernels.py:402-425) and
// `scale` a power of two derived from max|grad| so that every entry fits fp16 comfortably; both operands are unscaled
// (hi, lo) splits, three UMMAs per K step, fp32 accumulation.
//
// Wt is built by scatter with one thread per (neighbour slot, kernel); slots are pre-grouped by collision rank
// (k_tile_meta, bucket.cu) so that group 0 is a plain store and the few higher groups read-modify-write after a barrier:
// deterministic, no atomics.  Bond-attribute gradients (8 floats per support row) accumulate in registers.  The partial
// dxh of the kernel blocks are summed through an L2-resident scratch in fixed block order by the same CTA (static tile
// assignment); the last block applies d(x/|x|)/dx and writes grad_x.  k_param_finalize (params.cu) reduces the per-CTA
// copies of G in fixed order.
#include <algorithm>
#include "common.cuh"
#include "tc.cuh"
#include "tile.cuh"

namespace mk {

bool tile_layer_ok(const molkgnn_layer_t* layer);

constexpr int TB_THREADS = 512;
constexpr int TB_WARPS = TB_THREADS / 32;
constexpr int WT_ONE = 16 * 16 * 128;          // one fp16 image of the 128 x 128 coefficient block

struct BwdTileArgs {
    const float* x; const float* xnorm; int ldx;
    int F, Fp, Fk;
    const TileMetaG* meta; const float* ehat_node; const unsigned char* ximg; int n_tiles;
    int L[4], koff[4];
    const float* packed[4];
    const unsigned char* img;
    TileBlocks tb;
    int img_one, x_one;
    const float* coef;                 // chi * g per (node, kernel) pair, compact bucket order (k_coef)
    const uint8_t* argmax; long long scoff[4];
    const float* grad_absmax;          // device scalar: max |grad|
    float* partials; long long part_off[4]; int FW;
    float* scratch;                    // [N, Fk] partial dxh between kernel blocks
    float* gx; int ldgx;
    float* gx_absmax;                  // device scalar (nullable): max |grad_x|, for the next layer down
    int ne_cap, a_cap, buf_bytes;
    int sm_img, sm_x, sm_wt, sm_buf, sm_a, sm_am, sm_red;
};

struct BSeg {
    float alpha, beta;                 // w_s/(d W), w_c/W
    int d, k0, nk, rowbase, L, abase;
    float rnk;
};

__device__ __forceinline__ void tb_copy16(unsigned char* dst, const unsigned char* src, int64_t bytes) {
    for (int64_t i = (int64_t)threadIdx.x * 16; i < bytes; i += (int64_t)TB_THREADS * 16)
        *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(src + i);
}

__device__ __forceinline__ uint32_t tb_perm_code(int d, int p) {
    uint32_t c = 0;
    if (d == 4) {
#pragma unroll
        for (int q = 0; q < 12; ++q) if (q == p) c = perm_code<4>(q);
    } else if (d == 3) {
#pragma unroll
        for (int q = 0; q < 6; ++q) if (q == p) c = perm_code<3>(q);
    } else if (d == 2) {
        c = p == 0 ? perm_code<2>(0) : perm_code<2>(1);
    }
    return c;
}
__device__ __forceinline__ uint32_t tb_perm_inv_code(int d, int p) {
    uint32_t c = 0;
    if (d == 4) {
#pragma unroll
        for (int q = 0; q < 12; ++q) if (q == p) c = perm_inv_code<4>(q);
    } else if (d == 3) {
#pragma unroll
        for (int q = 0; q < 6; ++q) if (q == p) c = perm_inv_code<3>(q);
    } else if (d == 2) {
        c = p == 0 ? perm_inv_code<2>(0) : perm_inv_code<2>(1);
    }
    return c;
}

__device__ __forceinline__ void tb_split(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

// thread 0: metadata record, bond rows and node images of `tile`
__device__ __forceinline__ void tb_issue_copy(const BwdTileArgs& a, unsigned char* smem, unsigned char* buf, int tile,
                                              uint64_t* bar) {
    const TileMetaG* g = a.meta + tile;
    const int e0 = g->e0, ne = g->ne;
    const uint32_t eb = (uint32_t)ne * EP * 4u;
    mbar_expect_tx(bar, (uint32_t)sizeof(TileMetaG) + eb + 2u * (uint32_t)a.x_one);
    bulk_g2s(buf, g, (uint32_t)sizeof(TileMetaG), bar);
    if (eb) bulk_g2s(buf + sizeof(TileMetaG), a.ehat_node + (size_t)e0 * EP, eb, bar);
    bulk_g2s(smem + a.sm_x, a.ximg + (size_t)tile * 2 * a.x_one, 2u * (uint32_t)a.x_one, bar);
}

// thread 0: G (TMEM columns 0..Fk) += Wt . xhat ;  dxh (TMEM columns 128..128+Fk) = Wt^T . khat
__device__ __forceinline__ void tb_issue_mma(const BwdTileArgs& a, unsigned char* smem, int nn, int rows, bool first,
                                             uint32_t tmem, uint64_t* bar) {
    const uint32_t whi = tc::smem_u32(smem + a.sm_wt), wlo = whi + WT_ONE;
    const uint32_t xhi = tc::smem_u32(smem + a.sm_x), xlo = xhi + (uint32_t)a.x_one;
    const uint32_t ihi = tc::smem_u32(smem + a.sm_img), ilo = ihi + (uint32_t)a.img_one;
    const uint32_t fgrp = (uint32_t)(a.Fk >> 3) * 128u;      // bytes of one 8-row group of an [R x Fk] image
    const uint32_t idesc_g = tc::idesc_f16(128, a.Fk, 0, 1);  // A = Wt K-major (K = node), B = xhat MN-major (N = feature)
    const uint32_t idesc_x = tc::idesc_f16(128, a.Fk, 1, 1);  // A = Wt MN-major (M = node, K = row), B = khat MN-major
    const uint32_t dG = tmem, dX = tmem + 128u;
    // G: K = nodes, 16 per step.  Wt K-major: +256 B per step (two 8-column chunks); xhat MN-major: +2 row groups per step
    const int nkg = (max(16, (nn + 15) & ~15)) >> 4;
    for (int ks = 0; ks < nkg; ++ks) {
        const uint64_t dAh = tc::smem_desc(whi + ks * 256u, 128u, 2048u), dAl = tc::smem_desc(wlo + ks * 256u, 128u, 2048u);
        const uint64_t dBh = tc::smem_desc(xhi + ks * 2u * fgrp, fgrp, 128u), dBl = tc::smem_desc(xlo + ks * 2u * fgrp, fgrp, 128u);
        tc::umma_f16(dG, dAh, dBh, idesc_g, (first && ks == 0) ? 0u : 1u);
        tc::umma_f16(dG, dAl, dBh, idesc_g, 1u);
        tc::umma_f16(dG, dAh, dBl, idesc_g, 1u);
    }
    // dxh: K = kernel rows, 16 per step.  Wt MN-major: +2 row groups (2 * 2048 B) per step; khat MN-major likewise
    const int nkx = (max(16, (rows + 15) & ~15)) >> 4;
    for (int ks = 0; ks < nkx; ++ks) {
        const uint64_t dAh = tc::smem_desc(whi + ks * 4096u, 2048u, 128u), dAl = tc::smem_desc(wlo + ks * 4096u, 2048u, 128u);
        const uint64_t dBh = tc::smem_desc(ihi + ks * 2u * fgrp, fgrp, 128u), dBl = tc::smem_desc(ilo + ks * 2u * fgrp, fgrp, 128u);
        tc::umma_f16(dX, dAh, dBh, idesc_x, ks == 0 ? 0u : 1u);
        tc::umma_f16(dX, dAl, dBh, idesc_x, 1u);
        tc::umma_f16(dX, dAh, dBl, idesc_x, 1u);
    }
    tc::umma_commit(bar);
}

// coef[pair] = chi * g for every (node, kernel) pair in compact bucket order (streaming, fully occupied: keeps the gather
// of the incoming gradient rows out of the tile kernel's critical path)
struct CoefArgs {
    const int* sel; const int* nei;
    int n[4], boff[4], eoff[4], L[4], koff[4];
    long long scoff[4], tot;
    const float* grad; int ldg; int grad_mode;
    const uint8_t* argmax;
    float* coef;
};

__global__ void __launch_bounds__(256) k_coef(const CoefArgs a) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= a.tot) return;
    int d = 4;
    while (d > 1 && i < a.scoff[d - 1]) --d;
    const int L = a.L[d - 1];
    const long long li = i - a.scoff[d - 1];
    const int R = (int)(li / L), k = (int)(li - (long long)R * L);
    const int col = a.koff[d - 1] + k;
    float g;
    if (a.grad_mode == 0) {
        g = a.grad[(size_t)a.sel[a.boff[d - 1] + R] * a.ldg + col];
    } else {
        const int* nb = a.nei + (size_t)a.eoff[d - 1] + (size_t)R * d;
        g = a.grad[(size_t)nb[0] * a.ldg + col];
        for (int j = 1; j < d; ++j) g += a.grad[(size_t)nb[j] * a.ldg + col];
    }
    a.coef[i] = (a.argmax[i] & 0x80) ? -g : g;
}

__global__ void __launch_bounds__(TB_THREADS, 1) k_conv_bwd_tile(const __grid_constant__ BwdTileArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar_mma, bar_cp[2];
    __shared__ uint32_t tslot;
    __shared__ BSeg s_seg[TILE_MAXSEG];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        tc::mbar_init(&bar_mma, 1);
        tc::mbar_init(&bar_cp[0], 1); tc::mbar_init(&bar_cp[1], 1);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(&tslot, 256);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tslot;
    unsigned char* wt = smem + a.sm_wt;
    float* a_s = reinterpret_cast<float*>(smem + a.sm_a);
    unsigned char* am_s = smem + a.sm_am;
    float* red = reinterpret_cast<float*>(smem + a.sm_a);       // Jacobian reduction scratch: the coefficients are dead by then
    float (*s_eacc)[128][EP] = reinterpret_cast<float (*)[128][EP]>(wt);   // block end: bond partial sums (Wt is dead by then)
    uint32_t ph_mma = 0u, ph_cp[2] = {0u, 0u};
    // power-of-two scale: |alpha * chi * g| / scale <= 2^10 (g sums at most 4 gradient entries)
    float scale, rscale;
    {
        const float gm = fmaxf(*a.grad_absmax, 1e-30f) * 4.0f;
        int e;
        frexpf(gm, &e);                                   // gm < 2^e
        scale = ldexpf(1.0f, e - 10);
        rscale = ldexpf(1.0f, 10 - e);
    }
    float gmax_local = 0.f;
    const int q = warp & 3, cpart = warp >> 2;            // TMEM lane quadrant / 32-column part of this warp

    for (int blk = 0; blk < a.tb.nb; ++blk) {
        __syncthreads();
        // ---- block set-up ----
        tb_copy16(smem + a.sm_img, a.img + (size_t)blk * 2 * a.img_one, 2 * (int64_t)a.img_one);
        for (int i = tid * 16; i < 2 * WT_ONE; i += TB_THREADS * 16) *reinterpret_cast<uint4*>(wt + i) = make_uint4(0, 0, 0, 0);
        const int nseg = a.tb.nseg[blk];
        const int rows = a.tb.rows[blk];
        if (tid < nseg) {
            const TileSeg sg = a.tb.seg[blk][tid];
            const int L = a.L[sg.d - 1];
            const PackedLayout pl(sg.d, L, a.Fp);
            const float* pk = a.packed[sg.d - 1];
            BSeg c;
            c.alpha = pk[pl.w + 0] / pk[pl.w + 3] / (float)sg.d;
            c.beta = pk[pl.w + 1] / pk[pl.w + 3];
            c.d = sg.d; c.k0 = sg.k0; c.nk = sg.nk; c.rowbase = sg.rowbase; c.L = L; c.abase = 0;
            c.rnk = 1.0f / (float)sg.nk;
            s_seg[tid] = c;
        }
        // this thread's kernel row for the bond-gradient accumulation: row = tid & 127, node quarter = tid >> 7
        int er_seg = -1, er_slot = 0, er_kl = 0;
        {
            const int row = tid & 127;
            for (int si = 0; si < nseg; ++si) {
                const TileSeg sg = a.tb.seg[blk][si];
                const int r = row - sg.rowbase;
                if (r >= 0 && r < sg.nk * sg.d) { er_seg = si; er_slot = r / sg.nk; er_kl = r % sg.nk; }
            }
        }
        float eacc[EP];
#pragma unroll
        for (int c = 0; c < EP; ++c) eacc[c] = 0.f;
        tc::fence_async_smem();
        __syncthreads();
        bool first = true;
        int cur = 0;
        if (tid == 0 && (int)blockIdx.x < a.n_tiles) tb_issue_copy(a, smem, smem + a.sm_buf, blockIdx.x, &bar_cp[0]);

        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
            unsigned char* buf = smem + a.sm_buf + cur * a.buf_bytes;
            const TileMetaG& m = *reinterpret_cast<const TileMetaG*>(buf);
            const float* ehat = reinterpret_cast<const float*>(buf + sizeof(TileMetaG));
            tc::mbar_wait(&bar_cp[cur], ph_cp[cur]);
            ph_cp[cur] ^= 1u;
            const int t0 = m.t0, nn = m.nn;
            // segment offsets into the coefficient arrays
            int abase[TILE_MAXSEG];
            {
                int run = 0;
                for (int si = 0; si < nseg; ++si) { abase[si] = run; run += m.cnt[s_seg[si].d - 1] * s_seg[si].nk; }
            }
            // ---- (A) coefficients a = chi * g / scale, one thread per (node, kernel) pair; centre entries of Wt ----
            for (int si = 0; si < nseg; ++si) {
                const BSeg sg = s_seg[si];
                const int np = m.cnt[sg.d - 1] * sg.nk;
                for (int p = tid; p < np; p += TB_THREADS) {
                    const int ni = (int)(((float)p + 0.5f) * sg.rnk);
                    const int kl = p - ni * sg.nk;
                    const int nl_ = m.list[sg.d - 1][ni];
                    const int k = sg.k0 + kl;
                    const size_t cidx = (size_t)a.scoff[sg.d - 1] + (size_t)m.posl[nl_] * sg.L + k;
                    const float av = a.coef[cidx] * rscale;
                    const uint8_t am = a.argmax[cidx];
                    a_s[abase[si] + p] = av;
                    am_s[abase[si] + p] = am & 0x7f;
                    __half hi, lo;
                    tb_split(av * sg.beta, hi, lo);
                    const uint32_t off = tc::il_off(sg.rowbase + sg.d * sg.nk + kl, nl_, 128);
                    *reinterpret_cast<__half*>(wt + off) = hi;
                    *reinterpret_cast<__half*>(wt + WT_ONE + off) = lo;
                }
            }
            __syncthreads();
            // ---- (B) support entries of Wt: one thread per (neighbour slot, kernel), groups of rising collision rank ----
            for (int r = 0; r < 4; ++r) {
                bool any = false;
                for (int si = 0; si < nseg; ++si) {
                    const BSeg sg = s_seg[si];
                    const int e0 = m.eoffs[sg.d - 1][r], e1 = m.eoffs[sg.d - 1][r + 1];
                    const int ni_ = (e1 - e0) * sg.nk;
                    if (ni_ > 0) any = true;
                    for (int p = tid; p < ni_; p += TB_THREADS) {
                        const int ei = (int)(((float)p + 0.5f) * sg.rnk);
                        const int kl = p - ei * sg.nk;
                        const int ent = m.elist[e0 + ei];
                        const int nl_ = ent >> 2, j = ent & 3;
                        const int pi = abase[si] + m.lidx[nl_] * sg.nk + kl;
                        const float av = a_s[pi] * sg.alpha;
                        const int s = (tb_perm_code(sg.d, am_s[pi]) >> (2 * j)) & 3;
                        const int colv = (m.nl[nl_] >> (8 * j)) & 0xff;
                        const uint32_t off = tc::il_off(sg.rowbase + s * sg.nk + kl, colv, 128);
                        __half* ph = reinterpret_cast<__half*>(wt + off);
                        __half* pl_ = reinterpret_cast<__half*>(wt + WT_ONE + off);
                        float v = av;
                        if (r > 0) v += __half2float(*ph) + __half2float(*pl_);
                        __half hi, lo;
                        tb_split(v, hi, lo);
                        *ph = hi;
                        *pl_ = lo;
                    }
                }
                if (r < 3) {
                    bool more = false;
                    for (int si = 0; si < nseg; ++si) {
                        const int d = s_seg[si].d;
                        if (m.eoffs[d - 1][4] > m.eoffs[d - 1][r + 1]) more = true;
                    }
                    if (!more) break;
                    __syncthreads();
                }
                (void)any;
            }
            // ---- (C) bond-attribute gradients: thread (kernel row, node quarter), fixed node order ----
            if (er_seg >= 0) {
                const BSeg sg = s_seg[er_seg];
                const int cnt = m.cnt[sg.d - 1];
                const int qtr = tid >> 7;
                for (int ni = qtr; ni < cnt; ni += 4) {
                    const int nl_ = m.list[sg.d - 1][ni];
                    const int pi = abase[er_seg] + ni * sg.nk + er_kl;
                    const float av = a_s[pi];
                    const int j = (tb_perm_inv_code(sg.d, am_s[pi]) >> (2 * er_slot)) & 3;
                    const float* e = ehat + (size_t)(m.eslot[nl_] + j) * EP;
                    const float4 e0v = *reinterpret_cast<const float4*>(e), e1v = *reinterpret_cast<const float4*>(e + 4);
                    eacc[0] = fmaf(av, e0v.x, eacc[0]); eacc[1] = fmaf(av, e0v.y, eacc[1]);
                    eacc[2] = fmaf(av, e0v.z, eacc[2]); eacc[3] = fmaf(av, e0v.w, eacc[3]);
                    eacc[4] = fmaf(av, e1v.x, eacc[4]); eacc[5] = fmaf(av, e1v.y, eacc[5]);
                    eacc[6] = fmaf(av, e1v.z, eacc[6]); eacc[7] = fmaf(av, e1v.w, eacc[7]);
                }
            }
            tc::fence_async_smem();
            __syncthreads();
            // ---- (D) tensor cores ----
            if (tid == 0) {
                tc::fence_after_sync();
                tb_issue_mma(a, smem, nn, rows, first, tmem, &bar_mma);
            }
            first = false;
            // ---- (E) dxh epilogue: lane = node, 32 columns per warp.  Loads that do not depend on the tensor cores go first ----
            const bool lastb = blk + 1 == a.tb.nb;
            const int v = q * 32 + lane;
            const int f0 = cpart * 32;
            const bool colok = f0 < a.Fk;
            const bool rowok = v < nn;
            const int nf = min(32, a.Fk - f0);            // columns of this part (multiple of 16, or <= 0)
            float dv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) dv[i] = 0.f;
            float* sp = a.scratch + (size_t)(t0 + v) * a.Fk + f0;
            if (blk > 0 && rowok && colok) {              // partial sums of the previous kernel blocks
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    if (i < nf) {
                        const float4 o = __ldcg(reinterpret_cast<const float4*>(sp + i));
                        dv[i] = o.x; dv[i + 1] = o.y; dv[i + 2] = o.z; dv[i + 3] = o.w;
                    }
                }
            }
            float nrm = 1.f;
            if (lastb && rowok) nrm = a.xnorm[t0 + v];
            tc::mbar_wait(&bar_mma, ph_mma);
            ph_mma ^= 1u;
            tc::fence_after_sync();
            // the tensor cores are done with Wt: clear it for the next tile (ordered by the barrier that ends this tile)
            for (int i = tid * 16; i < 2 * WT_ONE; i += TB_THREADS * 16) *reinterpret_cast<uint4*>(wt + i) = make_uint4(0, 0, 0, 0);
            const int tnext = tile + gridDim.x;
            if (!lastb && tid == 0 && tnext < a.n_tiles)
                tb_issue_copy(a, smem, smem + a.sm_buf + (cur ^ 1) * a.buf_bytes, tnext, &bar_cp[cur ^ 1]);
            if (colok) {
                uint32_t u[32];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + 128u + (uint32_t)f0;
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
            if (!lastb) {
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
                                gmax_local = fmaxf(gmax_local, fabsf(o[c]));
                            }
                            st4(out + i, make_float4(o[0], o[1], o[2], o[3]));
                        }
                    }
                }
            }
            tc::fence_before_sync();
            __syncthreads();                 // Wt, coefficient arrays, TMEM dxh and this tile's buffer are free again
            cur ^= 1;
        }
        // ---- block end: kernel-parameter partial sums of this CTA ----
        {
            // bond part: sum the four node quarters in fixed order
#pragma unroll
            for (int c = 0; c < EP; ++c) s_eacc[tid >> 7][tid & 127][c] = eacc[c];
            __syncthreads();
            const int row = q * 32 + lane;
            int d = 0, slot = 0, kk = 0, L = 0;
            for (int si = 0; si < nseg; ++si) {
                const TileSeg sg = a.tb.seg[blk][si];
                const int r = row - sg.rowbase;
                if (r >= 0 && r < sg.nk * (sg.d + 1)) { d = sg.d; slot = r / sg.nk; kk = sg.k0 + r % sg.nk; L = a.L[sg.d - 1]; }
            }
            const int f0 = cpart * 32;
            if (f0 < a.Fk) {
                uint32_t u[32];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)f0;
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
                    const bool has_tiles = (int)blockIdx.x < a.n_tiles;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        if (f0 + i + 4 <= a.Fp) {
                            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (has_tiles)
                                o = make_float4(__uint_as_float(u[i]) * scale, __uint_as_float(u[i + 1]) * scale,
                                                __uint_as_float(u[i + 2]) * scale, __uint_as_float(u[i + 3]) * scale);
                            st4(part + f0 + i, o);
                        }
                    }
                    if (cpart == 0) {
                        float e[EP];
#pragma unroll
                        for (int c = 0; c < EP; ++c) {
                            e[c] = 0.f;
                            if (slot < d) e[c] = ((s_eacc[0][row][c] + s_eacc[1][row][c]) + (s_eacc[2][row][c] + s_eacc[3][row][c])) * scale;
                        }
                        st4(part + a.Fp, make_float4(e[0], e[1], e[2], e[3]));
                        st4(part + a.Fp + 4, make_float4(e[4], e[5], e[6], e[7]));
                    }
                }
            }
            tc::fence_before_sync();
        }
    }
    if (a.gx_absmax) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gmax_local = fmaxf(gmax_local, __shfl_xor_sync(0xffffffffu, gmax_local, o));
        if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(a.gx_absmax), __float_as_uint(gmax_local));
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// max |x| of a buffer into a device scalar (the caller zeroes it): the scale of the fp16 coefficient operand
__global__ void __launch_bounds__(256) k_absmax(const float* __restrict__ x, long long n4, long long n, float* out) {
    float m = 0.f;
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (long long i = n4 * 4 + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(m));
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
    return tile_plan_ok_b(plan) && layer->tile_img && tile_layer_ok(layer);
}

// returns 1 if launched, 0 if not eligible, <0 on error.  part_off / ncta describe the partial copies for k_param_finalize.
int launch_conv_bwd_tile(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                         const float* xnorm, const void* ximg, const float* grad, int32_t ldg, int32_t grad_mode,
                         const float* grad_absmax, const uint8_t* argmax, const int64_t scoff[4], float* coef,
                         float* partials, float* scratch, float* grad_x, int32_t ldgx, float* gx_absmax, int64_t part_off[4],
                         int ncta[4], int64_t* part_total, bool do_launch, cudaStream_t st) {
    if (!ximg || !grad_absmax || !tile_bwd_ok(plan, layer)) return 0;
    static int s_budget = 0;
    if (!s_budget) {
        s_budget = device_max_smem_optin();
        MK_REQUIRE(s_budget > 0, "conv_bwd_tile: no CUDA device");
    }
    BwdTileArgs a;
    if (!a.tb.build(layer->L)) return 0;
    MK_REQUIRE(a.tb.nb == 1 || scratch, "conv_bwd_tile: scratch is required for layers with more than one kernel block");
    a.x = x; a.xnorm = xnorm; a.ldx = ldx;
    a.F = layer->F; a.Fp = layer->Fp; a.Fk = tile_fk(layer->Fp);
    a.meta = reinterpret_cast<const TileMetaG*>(plan->tile_meta);
    a.ehat_node = plan->ehat_node;
    a.ximg = reinterpret_cast<const unsigned char*>(ximg);
    a.n_tiles = plan->n_tiles;
    const int grid = tile_bwd_grid(plan);
    a.FW = layer->Fp + EP;
    int64_t po = 0;
    for (int d = 0; d < 4; ++d) {
        a.L[d] = layer->L[d]; a.koff[d] = layer->koff[d];
        a.packed[d] = layer->packed[d];
        a.scoff[d] = scoff[d];
        a.part_off[d] = part_off[d] = po;
        ncta[d] = layer->L[d] > 0 ? grid : 0;
        po += (int64_t)ncta[d] * (d + 2) * layer->L[d] * a.FW;
    }
    *part_total = po;
    a.img = reinterpret_cast<const unsigned char*>(layer->tile_img);
    a.img_one = tile_img_one(a.Fk);
    a.x_one = tile_img_one(a.Fk);
    a.coef = coef;
    a.argmax = argmax;
    a.grad_absmax = grad_absmax;
    a.partials = partials;
    a.scratch = scratch;
    a.gx = grad_x; a.ldgx = ldgx; a.gx_absmax = gx_absmax;
    // capacities from the plan: bond slots and (node, kernel) pairs of the fullest tile
    int ne_cap = 0, a_cap = 0;
    for (int d = 0; d < 4; ++d) ne_cap += plan->tile_max_deg[d] * (d + 1);
    ne_cap = std::min(ne_cap, TILE_ESLOTS);
    for (int b = 0; b < a.tb.nb; ++b) {
        int c = 0;
        for (int si = 0; si < a.tb.nseg[b]; ++si) c += plan->tile_max_deg[a.tb.seg[b][si].d - 1] * a.tb.seg[b][si].nk;
        a_cap = std::max(a_cap, c);
    }
    a.ne_cap = ne_cap; a.a_cap = a_cap;
    a.buf_bytes = (int)((sizeof(TileMetaG) + (size_t)ne_cap * EP * 4 + 127) / 128 * 128);
    int64_t off = 0;
    a.sm_img = (int)off; off += 2 * (int64_t)a.img_one;
    a.sm_x = (int)off; off += 2 * (int64_t)a.x_one;
    a.sm_wt = (int)off; off += 2 * (int64_t)WT_ONE;
    a.sm_buf = (int)off; off += 2 * (int64_t)a.buf_bytes;
    a.sm_a = (int)off; off += (std::max<int64_t>((int64_t)a_cap * 4, 4 * 128 * 4) + 127) / 128 * 128;
    a.sm_am = (int)off; off += ((int64_t)a_cap + 127) / 128 * 128;
    a.sm_red = a.sm_a;
    if (off > s_budget - 2048) return 0;
    if (!do_launch) return 1;
    static int64_t s_attr = 0;
    if (off > s_attr) {
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_conv_bwd_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off));
        s_attr = off;
    }
    if (gx_absmax) MK_CHECK_CUDA(cudaMemsetAsync(gx_absmax, 0, sizeof(float), st));
    {
        CoefArgs c;
        c.sel = plan->sel; c.nei = plan->nei;
        long long tot = 0;
        for (int d = 0; d < 4; ++d) {
            c.n[d] = plan->n[d]; c.boff[d] = plan->boff[d]; c.eoff[d] = plan->eoff[d];
            c.L[d] = layer->L[d]; c.koff[d] = layer->koff[d]; c.scoff[d] = scoff[d];
            tot = std::max<long long>(tot, scoff[d] + (long long)plan->n[d] * layer->L[d]);
        }
        c.tot = tot;
        c.grad = grad; c.ldg = ldg; c.grad_mode = grad_mode;
        c.argmax = argmax; c.coef = coef;
        if (tot > 0) {
            count_launches(1);
            k_coef<<<(int)((tot + 255) / 256), 256, 0, st>>>(c);
            MK_CHECK_CUDA(cudaGetLastError());
        }
    }
    count_launches(1);
    k_conv_bwd_tile<<<grid, TB_THREADS, off, st>>>(a);
    MK_CHECK_CUDA(cudaGetLastError());
    return 1;
}

}  // namespace mk

using namespace mk;

extern "C" int molkgnn_absmax(const float* x, int64_t n, float* out, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    MK_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "absmax: x must be 16-byte aligned");
    MK_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
    if (n <= 0) return 0;
    const int grid = (int)std::min<int64_t>(148 * 8, (n / 4 + 255) / 256 + 1);
    count_launches(1);
    k_absmax<<<grid, 256, 0, st>>>(x, n / 4, n, out);
    MK_CHECK_CUDA(cudaGetLastError());
    return 0;
}
