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
//   k_conv_bwd_tile  tile-major persistent kernel, ONE launch per layer.  TMEM holds 2 x G_b + dxh = 336 of the 512 columns, so a
//                    layer with 4 blocks at Fk = 112 is two PASSES over the CTA's tiles inside the launch -- blocks {0, 1} (base
//                    model: the degree-4 blocks), flush of their G, blocks {2, 3} -- and the first pass hands its partial dxh to the
//                    second through `scratch` (only the rows its blocks touch; written and re-read by the same thread); layer 0
//                    (Fk = 32) holds all accumulators at once.  Per (tile, block): Wt by scatter with one thread per (node,
//                    kernel) pair; neighbour slots are pre-grouped by collision rank (k_tile_meta, bucket.cu) so that rank 0 is a
//                    plain store and the followers of a collision chain are added by one thread per kernel after a barrier --
//                    deterministic, no atomics; then 48 UMMAs from two issuing threads.  The block's images stream through
//                    shared memory by bulk copy; metadata / coefficients are double buffered and fetched a tile ahead.  Per tile
//                    one dxh epilogue; the last pass applies d(x/|x|)/dx and writes grad_x.
// k_param_finalize (params.cu) reduces the per-CTA copies in fixed order.
#include <algorithm>
#include <cstring>
#include "common.cuh"
#include "tc.cuh"
#include "tile.cuh"

namespace mk {

bool tile_layer_ok(const molkgnn_layer_t* layer);

static int64_t g_coef_attr_dev[16] = {0};      // dynamic shared memory k_coef_tile is configured for, per device (two launchers)

constexpr int TB_THREADS = 512;
constexpr int WT_ONE = 16 * 16 * 128;          // one fp16 image of the 128 x 128 coefficient block

// =============================================================================================================
// k_coef_tile
// =============================================================================================================
// Tile-ordered pre-pass.  Per tile the incoming gradient rows of its nodes are ONE contiguous range of grad (tiles hold
// whole molecules), staged in shared memory by cp.async one tile ahead; g[n,k] = sum over the node's neighbours (all inside
// the tile) is then shared-memory arithmetic.  Outputs per tile, in the order the tile kernel consumes them (one bulk copy):
//   coefT[tile * stride + off_d + i * L_d + k] = chi * g   for the i-th degree-d node of the tile (list[d-1][i]),
//   amT  [tile * stride_am + same index]       = saved arg-max permutation,
// off_d = sum_{d' < d} cnt_d' * L_d'.  Bond-attribute support gradients: thread owns support rows, sums over the tile's
// nodes in fixed order, one partial copy per CTA (k_param_finalize reduces them).
constexpr int CT_THREADS = 512;
constexpr int CT_MAXR = 4;                     // bond-gradient items per thread (sum_d d * L_d * nch_d <= 2048)
constexpr int CT_CHUNK = 12;                   // nodes per item and tile

struct CoefTileArgs {
    const TileMetaG* meta; const float* ehat_node; int n_tiles;
    const int* order; int order_grid;          // balanced tile schedule (tile.cuh TileWalk), nullable
    int L[4], koff[4];
    long long scoff[4];
    const float* grad; int ldg; int grad_mode; int vec;     // (vec: unused since the staging went to bulk copies)
    long long grad_floats;                                  // floats of the gradient array the staging may read (N * ldg)
    const uint8_t* argmax;
    float* coefT; uint8_t* amT; int stride, stride_am;
    int nch[4];                                // node chunks per degree for the bond gradients
    const uint8_t* amT_in;                     // tile-ordered arg-max written by the forward (nullable: read `argmax`, write amT)
    float* partials; long long part_off[4]; int FW, Fp;
    float* amax;                               // [gridDim.x] max |coef| seen by every CTA (no zeroing, no atomics)
    int buf_bytes, sm_grad, sm_eh, sm_am, sm_coef, sm_inv;   // per-buffer size / offsets inside a buffer / pair arrays
    // wide layers (conv_bwd_wide.cu): the gradient rows are read from global memory (a tile's rows do not fit shared memory)
    // and coefT / amT are written in BLOCKED tile order -- the pairs of one kernel block (WideBlocks: degree d, kernels
    // k0 .. k0+nk-1) contiguous: off_d + cnt_d * k0 + i * nk + (k - k0).  wb_base / wb_rem: kernels per block of degree d
    // (the first wb_rem blocks hold wb_base + 1).
    int wide, wb_base[4], wb_rem[4];
};

__device__ __forceinline__ void cp_async_n(void* dst, const void* src, int bytes) {
    const uint32_t d = tc::smem_u32(dst);
    if (bytes == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    else if (bytes == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// thread 0: metadata record, gradient rows, bond rows and arg-max codes of `tile` -> buffer, as bulk copies completing on `bar`
// (the first version had all 512 threads issue 16-byte cp.async: 4.3 k cycles of pure issue per tile).  The gradient rows of a
// tile are one contiguous range of `grad`; the copy starts at the 16-byte boundary below it and ends at the one above it -- or
// at the last whole 16 bytes of the array, the <= 3 floats behind which are moved by plain loads / stores.
__device__ __forceinline__ void ct_issue(const CoefTileArgs& a, unsigned char* buf, int tile, uint64_t* bar, const int4 hdr,
                                         const int4 c) {               // hdr = (t0, nn, e0, ne), c = nodes per degree
    const TileMetaG* g = a.meta + tile;
    uint32_t tx = (uint32_t)sizeof(TileMetaG) + (uint32_t)hdr.w * 32u;
    long long f0 = 0, f1 = 0, fend = 0;
    if (!a.wide) {
        const long long first = (long long)hdr.x * a.ldg;
        fend = first + (long long)hdr.y * a.ldg;
        f0 = first & ~3ll;
        f1 = min((fend + 3) & ~3ll, a.grad_floats & ~3ll);
        if (f1 > f0) tx += (uint32_t)(f1 - f0) * 4u;
    }
    uint32_t ab = 0;
    if (a.amT_in) {
        ab = (uint32_t)((c.x * a.L[0] + c.y * a.L[1] + c.z * a.L[2] + c.w * a.L[3] + 15) & ~15);
        tx += ab;
    }
    mbar_expect_tx(bar, tx);
    bulk_g2s(buf, g, (uint32_t)sizeof(TileMetaG), bar);
    if (hdr.w > 0) bulk_g2s(buf + a.sm_eh, a.ehat_node + (size_t)hdr.z * EP, (uint32_t)hdr.w * 32u, bar);
    if (!a.wide) {
        float* dst = reinterpret_cast<float*>(buf + a.sm_grad);
        if (f1 > f0) bulk_g2s(dst, a.grad + f0, (uint32_t)(f1 - f0) * 4u, bar);
        for (long long f = max(f1, f0); f < fend; ++f) dst[f - f0] = __ldg(a.grad + f);       // tail of the array (<= 3 floats)
    }
    if (ab) bulk_g2s(buf + a.sm_am, a.amT_in + (size_t)tile * a.stride_am, ab, bar);
}

// (node, kernel) pairs of degree D of one tile: chi * g from the staged gradient rows, tile-ordered outputs
template <int D>
__device__ __forceinline__ void ct_pairs(const CoefTileArgs& a, const TileMetaG& m, const float* gS, const unsigned char* amS,
                                         const unsigned char* inv_lut, int off, float* cT, uint8_t* aT, float* coefS,
                                         unsigned char* invS, float& amax) {
    const int L = a.L[D - 1];
    if (L == 0) return;
    const int np = m.cnt[D - 1] * L;
    const float rL = 1.0f / (float)L;
    const int koff = a.koff[D - 1];
    const int wbase = a.wb_base[D - 1], wrem = a.wb_rem[D - 1], cntd = m.cnt[D - 1];
#pragma unroll 2
    for (int p = threadIdx.x; p < np; p += CT_THREADS) {
        const int i = (int)(((float)p + 0.5f) * rL);
        const int k = p - i * L;
        const int nl_ = m.list[D - 1][i];
        int po = p;                                     // position inside the degree's part of the tile-ordered arrays
        if (a.wide) {
            const int big = wrem * (wbase + 1);
            int k0, nk;
            if (k < big) { const int j = k / (wbase + 1); k0 = j * (wbase + 1); nk = wbase + 1; }
            else { const int j = (k - big) / wbase; k0 = big + j * wbase; nk = wbase; }
            po = cntd * k0 + i * nk + (k - k0);
        }
        const uint8_t am = amS ? amS[off + p] : a.argmax[(size_t)a.scoff[D - 1] + (size_t)m.posl[nl_] * L + k];
        float g;
        if (a.grad_mode == 0) {
            g = gS[nl_ * a.ldg + koff + k];
        } else {
            const uint32_t nw = m.nl[nl_];
            g = gS[(int)(nw & 0xffu) * a.ldg + koff + k];
#pragma unroll
            for (int j = 1; j < D; ++j) g += gS[(int)((nw >> (8 * j)) & 0xffu) * a.ldg + koff + k];
        }
        const float av = (am & 0x80) ? -g : g;
        amax = fmaxf(amax, fabsf(av));
        cT[off + po] = av;
        if (aT) aT[off + po] = am & 0x7f;
        coefS[off + p] = av;
        invS[off + p] = inv_lut[am & 0x7f];
    }
}

#ifdef MK_PHASE_CLOCKS
__device__ unsigned long long g_ph_coef[16];
#endif

__global__ void __launch_bounds__(CT_THREADS, 1) k_coef_tile(const __grid_constant__ CoefTileArgs a) {
    extern __shared__ __align__(128) unsigned char smem_c[];
    __shared__ unsigned char s_inv[4][12];               // inverse permutation codes per degree
    __shared__ uint64_t bar_ct[2];                       // completion of the bulk copies into tile buffer 0 / 1
    const int tid = threadIdx.x;
    MK_PH_DECL(tid == 0)
    if (tid == 0) {
        tc::mbar_init(&bar_ct[0], 1); tc::mbar_init(&bar_ct[1], 1);
        tc::fence_mbar_init();
    }
    if (tid < 48) {
        s_inv[tid / 12][tid % 12] = c_perm_inv_code[tid / 12][tid % 12];
    }
    float* coefS = reinterpret_cast<float*>(smem_c + a.sm_coef);
    unsigned char* invS = smem_c + a.sm_inv;
    // Bond-gradient work items owned by this thread: item r = tid + q * CT_THREADS over the concatenation, degree by degree,
    // of (node chunk c, support row (s, k)).  A long degree bucket is cut into nch[d] chunks of CT_CHUNK nodes so that no
    // thread walks more than a few nodes per tile (a thread per whole row left two warps walking ~50 nodes while the rest
    // idled); the chunks' sums are combined in chunk order at the end.
    int rd[CT_MAXR], rs[CT_MAXR], rk[CT_MAXR], rc[CT_MAXR];
    float acc[CT_MAXR][EP];
#pragma unroll
    for (int q = 0; q < CT_MAXR; ++q) {
        int r = tid + q * CT_THREADS;
        rd[q] = 0; rs[q] = 0; rk[q] = 0; rc[q] = 0;
#pragma unroll
        for (int d = 1; d <= 4; ++d) {
            const int rows = d * a.L[d - 1];
            const int items = rows * a.nch[d - 1];
            if (rd[q] == 0 && r >= 0) {
                if (r < items) {
                    rd[q] = d; rc[q] = r / rows;
                    const int rr = r - rc[q] * rows;
                    rs[q] = rr / a.L[d - 1]; rk[q] = rr - rs[q] * a.L[d - 1];
                } else r -= items;
            }
        }
#pragma unroll
        for (int e = 0; e < EP; ++e) acc[q][e] = 0.f;
    }
    float amax = 0.f;
    int cur = 0;
    const TileWalk walk(a.order, a.order_grid, a.n_tiles);
    // thread 0 keeps the header and the degree counts of the NEXT tile to issue in registers, loaded one tile earlier: the
    // global-memory latency of the two loads stays off its (= every barrier's) critical path
    int4 nh = make_int4(0, 0, 0, 0), nc = make_int4(0, 0, 0, 0);
    auto ct_load_hdr = [&](int k) {
        if (k < walk.cnt) {
            const TileMetaG* g = a.meta + walk.tile(k);
            nh = __ldg(reinterpret_cast<const int4*>(g));
            nc = __ldg(reinterpret_cast<const int4*>(&g->cnt[0]));
        }
    };
    if (tid == 0 && walk.cnt > 0) {
        ct_load_hdr(0);
        ct_issue(a, smem_c, walk.tile(0), &bar_ct[0], nh, nc);
        ct_load_hdr(1);
    }
    MK_PH(0);
    for (int wk = 0; wk < walk.cnt; ++wk) {
        const int tile = walk.tile(wk);
        unsigned char* buf = smem_c + (size_t)cur * a.buf_bytes;
        const int tnext = wk + 1 < walk.cnt ? walk.tile(wk + 1) : a.n_tiles;
        __syncthreads();                                 // the previous tile's readers are done with the other buffer; coefS / invS are free
        if (tid == 0 && tnext < a.n_tiles) {
            ct_issue(a, smem_c + (size_t)(cur ^ 1) * a.buf_bytes, tnext, &bar_ct[cur ^ 1], nh, nc);
            ct_load_hdr(wk + 2);
        }
        MK_PH(1);                                        // barrier + issue of the next tile's copies
        tc::mbar_wait(&bar_ct[cur], ((uint32_t)wk >> 1) & 1u);   // this tile's buffer is complete
        MK_PH(2);                                        // wait for this tile's data
        const TileMetaG& m = *reinterpret_cast<const TileMetaG*>(buf);
        const float* gS = a.wide ? a.grad + (size_t)m.t0 * a.ldg
                                 : reinterpret_cast<const float*>(buf + a.sm_grad) + (int)(((long long)m.t0 * a.ldg) & 3ll);
        const float4* ehS = reinterpret_cast<const float4*>(buf + a.sm_eh);
        int off[5];
        off[0] = 0;
#pragma unroll
        for (int d = 0; d < 4; ++d) off[d + 1] = off[d] + m.cnt[d] * a.L[d];
        float* cT = a.coefT + (size_t)tile * a.stride;
        uint8_t* aT = a.amT ? a.amT + (size_t)tile * a.stride_am : nullptr;
        const unsigned char* amS = a.amT_in ? buf + a.sm_am : nullptr;
        // ---- coefficients: one thread per (node, kernel) pair, degree by degree ----
        ct_pairs<1>(a, m, gS, amS, s_inv[0], off[0], cT, aT, coefS, invS, amax);
        ct_pairs<2>(a, m, gS, amS, s_inv[1], off[1], cT, aT, coefS, invS, amax);
        ct_pairs<3>(a, m, gS, amS, s_inv[2], off[2], cT, aT, coefS, invS, amax);
        ct_pairs<4>(a, m, gS, amS, s_inv[3], off[3], cT, aT, coefS, invS, amax);
        MK_PH(3);                                        // pairs (thread 0's share)
        __syncthreads();
        MK_PH(4);                                        // waiting for the slowest pair thread
        // ---- bond-attribute support gradients ----
#pragma unroll
        for (int q = 0; q < CT_MAXR; ++q) {
            const int d = rd[q];
            if (d == 0) continue;
            const int L = a.L[d - 1], cnt = m.cnt[d - 1];
            const float* cs = coefS + off[d - 1] + rk[q];
            const unsigned char* is = invS + off[d - 1] + rk[q];
            const unsigned char* lst = m.list[d - 1];
            const int sh = 2 * rs[q];
            const int i0 = rc[q] * CT_CHUNK;
            const int i1 = rc[q] == a.nch[d - 1] - 1 ? cnt : min(cnt, i0 + CT_CHUNK);   // the last chunk takes the rest
#pragma unroll 4
            for (int i = i0; i < i1; ++i) {
                const float av = cs[i * L];
                const int j = (is[i * L] >> sh) & 3;
                const int e = m.eslot[lst[i]] + j;
                const float4 e0 = ehS[2 * e], e1 = ehS[2 * e + 1];
                acc[q][0] = fmaf(av, e0.x, acc[q][0]); acc[q][1] = fmaf(av, e0.y, acc[q][1]);
                acc[q][2] = fmaf(av, e0.z, acc[q][2]); acc[q][3] = fmaf(av, e0.w, acc[q][3]);
                acc[q][4] = fmaf(av, e1.x, acc[q][4]); acc[q][5] = fmaf(av, e1.y, acc[q][5]);
                acc[q][6] = fmaf(av, e1.z, acc[q][6]); acc[q][7] = fmaf(av, e1.w, acc[q][7]);
            }
        }
        MK_PH(5);                                        // bond sums (thread 0's rows)
        cur ^= 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    {
        __shared__ float s_amax[CT_THREADS / 32];
        if ((tid & 31) == 0) s_amax[tid >> 5] = amax;
        __syncthreads();
        if (tid == 0) {
            float m = s_amax[0];
            for (int w = 1; w < CT_THREADS / 32; ++w) m = fmaxf(m, s_amax[w]);
            a.amax[blockIdx.x] = m;
        }
    }
    // one partial copy per CTA: bond columns of the support rows (chunks combined in chunk order through shared memory,
    // the tile buffers are free now); the centre rows carry no bond part
    __syncthreads();
    float4* stage = reinterpret_cast<float4*>(smem_c);
#pragma unroll
    for (int q = 0; q < CT_MAXR; ++q) {
        if (rd[q] == 0) continue;
        const int r = tid + q * CT_THREADS;
        stage[2 * r] = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
        stage[2 * r + 1] = make_float4(acc[q][4], acc[q][5], acc[q][6], acc[q][7]);
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < CT_MAXR; ++q) {
        const int d = rd[q];
        if (d == 0 || rc[q] != 0) continue;                  // the chunk-0 owner of a row writes it
        const int L = a.L[d - 1], rows = d * L;
        const int r = tid + q * CT_THREADS;
        float4 s0 = stage[2 * r], s1 = stage[2 * r + 1];
        for (int c = 1; c < a.nch[d - 1]; ++c) {
            const float4 t0 = stage[2 * (r + c * rows)], t1 = stage[2 * (r + c * rows) + 1];
            s0.x += t0.x; s0.y += t0.y; s0.z += t0.z; s0.w += t0.w;
            s1.x += t1.x; s1.y += t1.y; s1.z += t1.z; s1.w += t1.w;
        }
        float* part = a.partials + a.part_off[d - 1] + ((size_t)blockIdx.x * (d + 1) * L + (size_t)rs[q] * L + rk[q]) * a.FW + a.Fp;
        st4(part, s0);
        st4(part + 4, s1);
    }
    for (int d = 1; d <= 4; ++d) {
        const int L = a.L[d - 1];
        for (int k = tid; k < L; k += CT_THREADS) {
            float* part = a.partials + a.part_off[d - 1] + ((size_t)blockIdx.x * (d + 1) * L + (size_t)d * L + k) * a.FW + a.Fp;
            st4(part, make_float4(0.f, 0.f, 0.f, 0.f));
            st4(part + 4, make_float4(0.f, 0.f, 0.f, 0.f));
        }
    }
    MK_PH(6);
    MK_PH_FLUSH(g_ph_coef);
}

// =============================================================================================================
// k_conv_bwd_tile
// =============================================================================================================
struct BwdTileArgs {
    const float* xnorm;
    int F, Fp, Fk;
    const TileMetaG* meta; const unsigned char* ximg; int n_tiles;
    const int* order; int order_grid;  // balanced tile schedule (tile.cuh TileWalk), nullable
    int L[4];
    const float* packed[4];
    const unsigned char* img;
    TileBlocks tb;
    int img_one, x_one;
    const float* coefT;                // chi * g per (node, kernel) pair, tile order (k_coef_tile)
    const uint8_t* amT; int stride, stride_am;
    const float* amax;                 // [gridDim.x] per-CTA max |coef| of k_coef_tile (same grid)
    float* partials; long long part_off[4]; int FW;
    float* scratch;                    // [N, Fk] partial dxh handed from the first launch to the second
    float* gx; int ldgx;
    int nbl, blist[4];                 // kernel blocks of this launch (k_conv_bwd_pipe; k_conv_bwd_tile reads the per-pass lists)
    int npass, p_nbl[2], p_blist[2][4]; // k_conv_bwd_tile: passes of this launch and their kernel blocks
    int gstride, dxcol;                // TMEM columns: G of block bi at bi * gstride, dxh at dxcol
    int nimg;                          // kernel-block image buffers in shared memory (1..4)
    int first, last;                   // first: no partial dxh to add; last: apply the Jacobian and write grad_x
    int gflush;                        // tiles per accumulation chunk of the G accumulators (see g_flush in k_conv_bwd_tile)
    int buf_bytes;
    int sm_img, sm_x, sm_wt, sm_buf, sm_a, sm_am, sm_red;
    int dmask_first;                   // two-launch layers: bit d set if the FIRST launch's blocks hold kernels of degree d
    int dbuf, a_bytes, am_bytes;       // dbuf: the coefficient / arg-max arrays are double buffered (copy b at sm_a + b * a_bytes)
    // pipelined kernel (k_conv_bwd_pipe): ring of K-step stages, two Wt buffers, two tile buffers [meta | coef | arg-max]
    int nstages, stage_bytes, sm_ring, tbuf_bytes, tb_a, tb_am, cs;   // cs: TMEM column stride of an accumulator
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
__device__ __forceinline__ void wt_store_hl(unsigned char* wt, int row, int col, __half hi, __half lo) {
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

// Rank-0 scatter of one segment (degree D): thread t owns the (node, kernel) pairs t, t + TB_THREADS, ...  A pair is one long
// dependent chain (list -> neighbour words -> coefficient, arg-max code -> permutation -> conversions -> stores, ~130
// instructions), and with 16 warps on 4 schedulers the phase is bound by that chain's latency times the pairs per thread.  The
// pairs of a thread are therefore taken TB_NU at a time: all their loads first (the compiler cannot hoist loads over the
// shared-memory stores of the previous pair on its own), then the conversions, then the stores.
constexpr int TB_NU = 3;
template <int D>
__device__ __forceinline__ void tb_scatter_seg(const BSeg& sg, const TileMetaG& m, int np, int abase, const float* a_s,
                                               const unsigned char* am_s, const unsigned char* lut, unsigned char* wt, float rscale,
                                               int tid) {
    for (int p0 = tid; p0 < np; p0 += TB_NU * TB_THREADS) {
        int kl[TB_NU], nl_[TB_NU];
        float av[TB_NU];
        uint32_t code[TB_NU], nw[TB_NU], cr[TB_NU];
        bool ok[TB_NU];
#pragma unroll
        for (int u = 0; u < TB_NU; ++u) {
            const int p = p0 + u * TB_THREADS;
            ok[u] = p < np;
            const int pc = ok[u] ? p : p0;
            const int ni = (int)(((float)pc + 0.5f) * sg.rnk);
            kl[u] = pc - ni * sg.nk;
            nl_[u] = m.list[D - 1][ni];
            const int pi = abase + ni * sg.L + kl[u];
            av[u] = a_s[pi] * rscale;
            code[u] = lut[am_s[pi] & 0x7f];
            nw[u] = m.nl[nl_[u]];
            cr[u] = m.cr[nl_[u]];
        }
#pragma unroll
        for (int u = 0; u < TB_NU; ++u) {
            if (!ok[u]) continue;
            wt_store(wt, sg.rowbase + D * sg.nk + kl[u], nl_[u], av[u] * sg.beta);
            const float as = av[u] * sg.alpha;                // the same value goes to all D support entries: split it once
            const __half ah = __float2half_rn(as);
            const __half al = __float2half_rn(as - __half2float(ah));
#pragma unroll
            for (int j = 0; j < D; ++j) {
                if (((cr[u] >> (2 * j)) & 3u) == 0u)
                    wt_store_hl(wt, sg.rowbase + (int)((code[u] >> (2 * j)) & 3u) * sg.nk + kl[u], (int)((nw[u] >> (8 * j)) & 0xffu), ah, al);
            }
        }
    }
}

// thread 0: metadata record, node images and the tile-ordered coefficients / arg-max codes of `tile` (np pairs).  The
// coefficient arrays and the image buffer are single: the caller issues this only after the previous tile's last use of them.
// The node images complete on their OWN barrier: only the first G MMA of the tile (and the Jacobian) need them, so the 57 KB
// land while the tile's first block is scattered instead of being waited for at the top of the tile.
// Split in two: the node images (single buffer: issued once the previous tile's last reader is done) and the metadata +
// coefficient arrays into buffer `b` -- with a.dbuf these are double buffered and fetched a whole tile ahead.
__device__ __forceinline__ void tb_issue_x(const BwdTileArgs& a, unsigned char* smem, int tile, uint64_t* bar_x) {
    mbar_expect_tx(bar_x, 2u * (uint32_t)a.x_one);
    bulk_g2s(smem + a.sm_x, a.ximg + (size_t)tile * 2 * a.x_one, 2u * (uint32_t)a.x_one, bar_x);
}
__device__ __forceinline__ void tb_issue_meta(const BwdTileArgs& a, unsigned char* smem, int b, int tile, int np, uint64_t* bar) {
    const uint32_t cb = (uint32_t)((np * 4 + 15) & ~15), ab = (uint32_t)((np + 15) & ~15);
    const int cb_i = a.dbuf ? b : 0;
    mbar_expect_tx(bar, (uint32_t)sizeof(TileMetaG) + cb + ab);
    bulk_g2s(smem + a.sm_buf + (size_t)b * a.buf_bytes, a.meta + tile, (uint32_t)sizeof(TileMetaG), bar);
    if (np > 0) {
        bulk_g2s(smem + a.sm_a + (size_t)cb_i * a.a_bytes, a.coefT + (size_t)tile * a.stride, cb, bar);
        bulk_g2s(smem + a.sm_am + (size_t)cb_i * a.am_bytes, a.amT + (size_t)tile * a.stride_am, ab, bar);
    }
}
__device__ __forceinline__ int tb_tile_pairs(const BwdTileArgs& a, int tile) {
    const int4 c = __ldg(reinterpret_cast<const int4*>(&a.meta[tile].cnt[0]));
    return c.x * a.L[0] + c.y * a.L[1] + c.z * a.L[2] + c.w * a.L[3];
}
// thread 0: images of kernel block `blk` -> image buffer `ib`
__device__ __forceinline__ void tb_issue_img(const BwdTileArgs& a, unsigned char* smem, int blk, int ib, uint64_t* bar) {
    mbar_expect_tx(bar, 2u * (uint32_t)a.img_one);
    bulk_g2s(smem + a.sm_img + (size_t)ib * 2 * a.img_one, a.img + (size_t)blk * 2 * a.img_one, 2u * (uint32_t)a.img_one, bar);
}

// thread 0: G_bi (TMEM columns bi * gstride ..) += Wt . xhat
__device__ __forceinline__ void tb_issue_mma_g(const BwdTileArgs& a, unsigned char* smem, int nn, int bi, bool g_fresh,
                                               uint32_t tmem) {
    const uint32_t whi = tc::smem_u32(smem + a.sm_wt), wlo = whi + WT_ONE;
    const uint32_t xhi = tc::smem_u32(smem + a.sm_x), xlo = xhi + (uint32_t)a.x_one;
    const uint32_t fgrp = (uint32_t)(a.Fk >> 3) * 128u;      // bytes of one 8-row group of an [R x Fk] image
    const uint32_t idesc_g = tc::idesc_f16(128, a.Fk, 0, 1);  // A = Wt K-major (K = node), B = xhat MN-major (N = feature)
    const uint32_t dG = tmem + (uint32_t)(bi * a.gstride);
    // K = nodes, 16 per step.  Wt K-major: +256 B per step (two 8-column chunks); xhat MN-major: +2 row groups per step
    const int nkg = (max(16, (nn + 15) & ~15)) >> 4;
    for (int ks = 0; ks < nkg; ++ks) {
        const uint64_t dAh = tc::smem_desc(whi + ks * 256u, 128u, 2048u), dAl = tc::smem_desc(wlo + ks * 256u, 128u, 2048u);
        const uint64_t dBh = tc::smem_desc(xhi + ks * 2u * fgrp, fgrp, 128u), dBl = tc::smem_desc(xlo + ks * 2u * fgrp, fgrp, 128u);
        tc::umma_f16(dG, dAh, dBh, idesc_g, (g_fresh && ks == 0) ? 0u : 1u);
        tc::umma_f16(dG, dAl, dBh, idesc_g, 1u);
        tc::umma_f16(dG, dAh, dBl, idesc_g, 1u);
    }
}
// thread 0: dxh (TMEM columns dxcol ..) (+)= Wt^T . khat (image buffer ib), then commit everything issued so far
__device__ __forceinline__ void tb_issue_mma_x(const BwdTileArgs& a, unsigned char* smem, int rows, bool x_fresh, int ib,
                                               uint32_t tmem, uint64_t* bar) {
    const uint32_t whi = tc::smem_u32(smem + a.sm_wt), wlo = whi + WT_ONE;
    const uint32_t ihi = tc::smem_u32(smem + a.sm_img + (size_t)ib * 2 * a.img_one), ilo = ihi + (uint32_t)a.img_one;
    const uint32_t fgrp = (uint32_t)(a.Fk >> 3) * 128u;
    const uint32_t idesc_x = tc::idesc_f16(128, a.Fk, 1, 1);  // A = Wt MN-major (M = node, K = row), B = khat MN-major
    const uint32_t dX = tmem + (uint32_t)a.dxcol;
    // K = kernel rows, 16 per step.  Wt MN-major: +2 row groups (2 * 2048 B) per step; khat MN-major likewise
    const int nkx = (max(16, (rows + 15) & ~15)) >> 4;
    for (int ks = 0; ks < nkx; ++ks) {
        const uint64_t dAh = tc::smem_desc(whi + ks * 4096u, 2048u, 128u), dAl = tc::smem_desc(wlo + ks * 4096u, 2048u, 128u);
        const uint64_t dBh = tc::smem_desc(ihi + ks * 2u * fgrp, fgrp, 128u), dBl = tc::smem_desc(ilo + ks * 2u * fgrp, fgrp, 128u);
        tc::umma_f16(dX, dAh, dBh, idesc_x, (x_fresh && ks == 0) ? 0u : 1u);
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
    __shared__ uint64_t bar_mma, bar_img[4], bar_cp[2], bar_xi;
    __shared__ uint32_t tslot;
    __shared__ BSeg s_seg[4][TILE_MAXSEG];
    __shared__ unsigned char s_lut[4][12];               // packed permutation codes (2 bits per j) per degree
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    MK_PH_DECL(tid == 0)
    if (tid == 0) {
        tc::mbar_init(&bar_xi, 1);
        tc::mbar_init(&bar_mma, 2);                       // one commit per issuing thread (G group, dxh group)
        for (int i = 0; i < 4; ++i) tc::mbar_init(&bar_img[i], 1);
        tc::mbar_init(&bar_cp[0], 1); tc::mbar_init(&bar_cp[1], 1);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(&tslot, 512);
    if (tid < 48) {
        s_lut[tid / 12][tid % 12] = c_perm_code[tid / 12][tid % 12];
    }
    __shared__ float s_gmax;
    if (warp == 3) {                                      // max |coef| over the per-CTA values of k_coef_tile
        float gm = 0.f;
        for (int i = lane; i < (int)gridDim.x; i += 32) gm = fmaxf(gm, __ldg(a.amax + i));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
        if (lane == 0) s_gmax = gm;
    }
    unsigned char* wt = smem + a.sm_wt;
    for (int i = tid * 16; i < 2 * WT_ONE; i += TB_THREADS * 16) *reinterpret_cast<uint4*>(wt + i) = make_uint4(0, 0, 0, 0);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tslot;
    float* red = reinterpret_cast<float*>(smem + a.sm_red);     // Jacobian reduction scratch
    uint32_t ph_mma = 0u, ph_cp[2] = {0u, 0u};
    // power-of-two scale: |alpha * chi * g| / scale <= 2^10
    float scale, rscale;
    {
        const float gm = fmaxf(s_gmax, 1e-30f);
        int e;
        frexpf(gm, &e);                                   // gm < 2^e
        scale = ldexpf(1.0f, e - 10);
        rscale = ldexpf(1.0f, 10 - e);
    }
    const int q = warp & 3, cpart = warp >> 2;            // TMEM lane quadrant / 32-column part of this warp
    // image uses u = 0, 1, ...: block blist[u % nbl] in buffer u % nimg.  With nbl <= nimg the images stay resident.
    const TileWalk walk(a.order, a.order_grid, a.n_tiles);
    const int my_tiles = walk.cnt;
    int np_next = 0;
    int cur = 0;
    uint32_t ph_img = 0u;                                 // phase bit of every image barrier (thread 32 waits on them)
    MK_PH(0);                                             // prologue
  // One launch runs a.npass passes over the CTA's tiles, each over its own <= 2 kernel blocks (TMEM: 2 x G_b + dxh = 336 of the 512
  // columns): a layer with 4 blocks at Fk = 112 is pass 0 over blocks {0, 1}, G flush, pass 1 over blocks {2, 3} -- the two
  // LAUNCHES this used to be, without the second launch's prologue and without every CTA waiting for the slowest one in between.
  // A pass hands its partial dxh to the next through `scratch` (rows written and re-read by the same thread).
  for (int pass = 0; pass < a.npass; ++pass) {
    const int nbl = a.p_nbl[pass];
    const int* blist = a.p_blist[pass];
    const bool pfirst = a.first && pass == 0, plast = a.last && pass == a.npass - 1;
    __syncthreads();                                      // the previous pass is done with s_seg
    if (tid >= 64 && tid < 64 + 4 * TILE_MAXSEG) {
        const int bi = (tid - 64) / TILE_MAXSEG, si = (tid - 64) % TILE_MAXSEG;
        if (bi < nbl && si < a.tb.nseg[blist[bi]]) {
            const TileSeg sg = a.tb.seg[blist[bi]][si];
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
    __syncthreads();
    const int total_uses = my_tiles * nbl;
    const bool resident = nbl <= a.nimg;
    const uint32_t xi_base = (uint32_t)(pass * my_tiles);  // bar_xi completes once per tile, over all passes
    if (tid == 0 && my_tiles > 0) {
        tb_issue_meta(a, smem, cur, walk.tile(0), tb_tile_pairs(a, walk.tile(0)), &bar_cp[cur]);
        tb_issue_x(a, smem, walk.tile(0), &bar_xi);
        const int n0 = resident ? nbl : min(a.nimg, total_uses);
        for (int u = 0; u < n0; ++u) tb_issue_img(a, smem, blist[u % nbl], u % a.nimg, &bar_img[u % a.nimg]);
    }
    if (tid == 64 && a.dbuf && my_tiles > 1) np_next = tb_tile_pairs(a, walk.tile(1));
    int use = 0;
    bool fresh = true;                                    // first tile of this pass: the G accumulators start from zero
    // Kernel-parameter partial sums of this CTA: node-attribute part of every row of this launch's blocks, tensor memory ->
    // the CTA's partial copy.  The tensor core TRUNCATES every accumulate (measured: against the fp32 SIMT backward the sums come
    // out smaller in magnitude, the error growing with the batch: 4e-6 at 6 tiles per CTA, 5e-5 at 96), so the accumulators are
    // flushed every `gflush` tiles -- chains of <= gflush * 24 MMAs -- and the chunks are added in fp32 (round to nearest) by
    // fire-and-forget reductions; every element has one writer thread, the order of its reductions is program order.
    auto g_flush = [&](bool add, bool valid) {
        for (int bi = 0; bi < nbl; ++bi) {
            const int blk = blist[bi];
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
#pragma unroll
                for (int i = 0; i < 32; ++i) u[i] = 0u;
                if (valid) {               // a CTA that saw no tile never ran an MMA: its accumulators are undefined, its sums zero
                    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(bi * a.gstride + f0);
                    tc::tmem_ld16(taddr, u);
                    if (f0 + 16 < a.Fk) tc::tmem_ld16(taddr + 16, u + 16);
                    tc::tmem_ld_wait();
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(u[i]));
                if (d > 0) {
                    const int rows_x = (d + 1) * L;
                    float* part = a.partials + a.part_off[d - 1] + ((size_t)blockIdx.x * rows_x + (size_t)slot * L + kk) * a.FW;
                    const float sc2 = valid ? scale : 0.f;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        if (f0 + i + 4 <= a.Fp) {
                            if (add) {
#pragma unroll
                                for (int c = 0; c < 4; ++c) atomicAdd(part + f0 + i + c, __uint_as_float(u[i + c]) * sc2);
                            } else {
                                st4(part + f0 + i, make_float4(__uint_as_float(u[i]) * sc2, __uint_as_float(u[i + 1]) * sc2,
                                                               __uint_as_float(u[i + 2]) * sc2, __uint_as_float(u[i + 3]) * sc2));
                            }
                        }
                    }
                }
            }
        }
    };
    int nflushed = 0;

    for (int wk = 0; wk < walk.cnt; ++wk) {
        const int tile = walk.tile(wk);
        unsigned char* buf = smem + a.sm_buf + cur * a.buf_bytes;
        const TileMetaG& m = *reinterpret_cast<const TileMetaG*>(buf);
        const int tnext = wk + 1 < walk.cnt ? walk.tile(wk + 1) : a.n_tiles;
        if (!a.dbuf && tid == 0 && tnext < a.n_tiles) np_next = tb_tile_pairs(a, tnext);   // latency hidden behind this tile's work
        const float* a_s = reinterpret_cast<const float*>(smem + a.sm_a + (size_t)(a.dbuf ? cur : 0) * a.a_bytes);
        const unsigned char* am_s = smem + a.sm_am + (size_t)(a.dbuf ? cur : 0) * a.am_bytes;
        tc::mbar_wait(&bar_cp[cur], ph_cp[cur]);
        ph_cp[cur] ^= 1u;
        MK_PH(1);                                         // wait for the tile's metadata, node images, coefficients
        const int t0 = m.t0, nn = m.nn;
        int doff[4];                                      // first pair of degree d in the tile-ordered arrays
        doff[0] = 0;
#pragma unroll
        for (int d = 1; d < 4; ++d) doff[d] = doff[d - 1] + m.cnt[d - 1] * a.L[d - 1];
        // Two-launch layers: the first launch's partial dxh is non-zero only in the rows its blocks touch -- nodes of those degrees
        // (centre entries) and their neighbours (support entries).  With the first launch on the degree-4 blocks that is ~30 % of the
        // rows for drug-like molecules; only those rows make the round trip through `scratch`.
        bool touched = true;
        if (pfirst != plast) {
            const int v = q * 32 + lane;
            touched = false;
            if (v < nn) {
                touched = ((a.dmask_first >> m.degl[v]) & 1) != 0;
                const uint32_t iw = m.inl[v];
                const int ic = m.incnt[v];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < ic && ((a.dmask_first >> m.degl[(iw >> (8 * j)) & 0xffu]) & 1)) touched = true;
            }
        }
        float dv[32];
        float nrm = 1.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) dv[i] = 0.f;
        for (int bi = 0; bi < nbl; ++bi) {
            const int blk = blist[bi];
            const int nseg = a.tb.nseg[blk];
            int abase[TILE_MAXSEG];
            for (int si = 0; si < nseg; ++si) abase[si] = doff[s_seg[bi][si].d - 1] + s_seg[bi][si].k0;
            // The image buffer the previous visit freed is refilled NOW, not at the start of that visit's Wt clear: the 57 KB of the
            // bulk copy and the 64 KB of clearing stores shared the 128 B/cycle of shared-memory bandwidth (clear + set-up 1.9 k
            // cycles per visit); the copy has the whole scatter to land.  Thread 96: not a thread the barriers wait for.
            if (tid == 96 && !resident && use >= 1 && use - 1 + a.nimg < total_uses)
                tb_issue_img(a, smem, blist[(use - 1 + a.nimg) % nbl], (use - 1) % a.nimg, &bar_img[(use - 1) % a.nimg]);
            MK_PH(14);                                    // tile / block set-up in front of the scatter
            // ---- rank 0: one thread per (node, kernel) pair -- centre entry, collision-free support entries ----
            for (int si = 0; si < nseg; ++si) {
                const BSeg sg = s_seg[bi][si];
                const int np = m.cnt[sg.d - 1] * sg.nk;
                switch (sg.d) {
                    case 1: tb_scatter_seg<1>(sg, m, np, abase[si], a_s, am_s, s_lut[0], wt, rscale, tid); break;
                    case 2: tb_scatter_seg<2>(sg, m, np, abase[si], a_s, am_s, s_lut[1], wt, rscale, tid); break;
                    case 3: tb_scatter_seg<3>(sg, m, np, abase[si], a_s, am_s, s_lut[2], wt, rscale, tid); break;
                    default: tb_scatter_seg<4>(sg, m, np, abase[si], a_s, am_s, s_lut[3], wt, rscale, tid); break;
                }
            }
            MK_PH(10 + blk);                              // rank-0 scatter (thread 0's own share), per kernel block
            // ---- collision chains: one thread per (chain, kernel) adds the chain's followers in in-edge order after ONE barrier ----
            {
                bool more = false;
                for (int si = 0; si < nseg; ++si) {
                    const int d = s_seg[bi][si].d;
                    if (m.choff[d] > m.choff[d - 1]) more = true;
                }
                if (more) {
                    __syncthreads();
                    for (int si = 0; si < nseg; ++si) {
                        const BSeg sg = s_seg[bi][si];
                        const int c0 = m.choff[sg.d - 1];
                        const int ni_ = (m.choff[sg.d] - c0) * sg.nk;
                        for (int p = tid; p < ni_; p += TB_THREADS) {
                            const int ci = (int)(((float)p + 0.5f) * sg.rnk);
                            const int kl = p - ci * sg.nk;
                            const uint32_t ch = m.chains[c0 + ci];
                            const int nf = (int)((ch >> 27) & 3u) + 1;
                            for (int f = 0; f < nf; ++f) {
                                const int ent = (int)((ch >> (9 * f)) & 0x1ffu);
                                const int nl_ = ent >> 2, j = ent & 3;
                                const int pi = abase[si] + m.lidx[nl_] * sg.L + kl;
                                const int s = (s_lut[sg.d - 1][am_s[pi] & 0x7f] >> (2 * j)) & 3;
                                wt_add(wt, sg.rowbase + s * sg.nk + kl, (int)((m.nl[nl_] >> (8 * j)) & 0xffu), (a_s[pi] * rscale) * sg.alpha);
                            }
                        }
                    }
                }
            }
            MK_PH(3);                                     // ranks 1..3 (incl. waiting for the slowest rank-0 thread)
            tc::fence_async_smem();
            __syncthreads();
            MK_PH(4);                                     // barrier before the MMAs
            // ---- tensor cores ----
            // TWO issuing threads: a tcgen05.mma costs the thread that issues it ~100 cycles (descriptor set-up, uniform-register
            // moves), about twice the tensor time of these shapes; the G MMAs and the dxh MMAs write different accumulators, so
            // warp 0 issues the first and warp 1 the second group and the tensor pipe sees both streams (bar_mma counts 2)
            if (tid == 0) {
                if (bi == 0) tc::mbar_wait(&bar_xi, (xi_base + (uint32_t)wk) & 1u);   // node images of this tile (one completion per tile)
                tc::fence_after_sync();
                tb_issue_mma_g(a, smem, nn, bi, fresh, tmem);
                tc::umma_commit(&bar_mma);
            } else if (tid == 32) {
                tc::fence_after_sync();
                const int ib = resident ? use % nbl : use % a.nimg;
                if (!resident || use < nbl) {                 // (resident images are waited for once per pass)
                    tc::mbar_wait(&bar_img[ib], (ph_img >> ib) & 1u);
                    ph_img ^= 1u << ib;
                }
                tb_issue_mma_x(a, smem, a.tb.rows[blk], bi == 0, ib, tmem, &bar_mma);
            } else if (tid == 64 && bi == 0 && a.dbuf) {
                // double buffered metadata / coefficients: the other buffer's readers finished with the previous tile.  Thread 64
                // fetches the next tile's a whole tile ahead while its warp would only sleep in the barrier below (any work on
                // thread 0 sits on every barrier's critical path), then loads the pair count of the tile after it.
                if (tnext < a.n_tiles) tb_issue_meta(a, smem, cur ^ 1, tnext, np_next, &bar_cp[cur ^ 1]);
                if (wk + 2 < walk.cnt) np_next = tb_tile_pairs(a, walk.tile(wk + 2));
            }
            ++use;
            MK_PH(5);                                     // image wait + MMA issue
            // Epilogue operands that come from global memory (partial dxh of the first launch, row norm) are fetched while the
            // tile's LAST block is in the tensor cores: ~3 k cycles ahead of the epilogue.  (Fetched at the top of the tile they cost
            // 6 k cycles per tile in front of the first scatter -- measured with a phase clock between the two: 128 16-byte loads
            // per warp ahead of the scatter's shared-memory loads in the same load / store queue.)
            // The 57 KB of a tile take ~6 k cycles to get through the load / store unit, two tensor-core phases: the first half of the
            // columns goes with the second-to-last block, the second half with the last.
            if (bi >= nbl - 2) {
                const int v = q * 32 + lane, f0 = cpart * 32;
                const int nf = min(32, a.Fk - f0);
                const bool h0 = nbl == 1 || bi == nbl - 2, h1 = bi == nbl - 1;
                if (!pfirst && touched && v < nn && f0 < a.Fk) {
                    const float* sp = a.scratch + (size_t)(t0 + v) * a.Fk + f0;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        if (i < nf && (i < 16 ? h0 : h1)) {
                            const float4 o = __ldcg(reinterpret_cast<const float4*>(sp + i));
                            dv[i] = o.x; dv[i + 1] = o.y; dv[i + 2] = o.z; dv[i + 3] = o.w;
                        }
                    }
                }
                if (h1 && plast && v < nn) nrm = __ldg(a.xnorm + t0 + v);
            }
            // the issuing thread alone polls for completion, everybody else sleeps in the hardware barrier: 511 threads
            // spinning on try_wait take issue slots from the one thread that feeds the tensor core
            if (tid == 0) tc::mbar_wait(&bar_mma, ph_mma);
            __syncthreads();
            ph_mma ^= 1u;
            tc::fence_after_sync();
            MK_PH(6);                                     // MMA completion
            // the tensor cores are done with Wt and this block's images: fetch the images of a later use, clear Wt
            for (int i = tid * 16; i < 2 * WT_ONE; i += TB_THREADS * 16) *reinterpret_cast<uint4*>(wt + i) = make_uint4(0, 0, 0, 0);
            if (bi + 1 < nbl) __syncthreads();              // Wt cleared before the next block's scatter
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
            float* sp = a.scratch + (size_t)(t0 + v) * a.Fk + f0;
            if (colok) {
                uint32_t u[32];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(a.dxcol + f0);
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
            if (!plast) {
                if (tid == 0 && tnext < a.n_tiles) {
                    if (!a.dbuf) tb_issue_meta(a, smem, cur ^ 1, tnext, np_next, &bar_cp[cur ^ 1]);
                    tb_issue_x(a, smem, tnext, &bar_xi);
                }
                if (rowok && colok && touched) {
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
                tc::mbar_wait(&bar_xi, (xi_base + (uint32_t)wk) & 1u);        // (complete since the first G MMA; acquire for this thread)
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
                if (tid == 0 && tnext < a.n_tiles) {    // every thread has read its xhat: the image buffer may be refilled
                    if (!a.dbuf) tb_issue_meta(a, smem, cur ^ 1, tnext, np_next, &bar_cp[cur ^ 1]);
                    tb_issue_x(a, smem, tnext, &bar_xi);
                }
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
        __syncthreads();                 // Jacobian scratch, TMEM dxh and this tile's buffer are free again
        MK_PH(8);                                         // dxh epilogue
        if ((wk + 1) % a.gflush == 0 && wk + 1 < walk.cnt) {      // end of an accumulation chunk, more tiles follow
            tc::fence_after_sync();
            g_flush(nflushed > 0, true);
            ++nflushed;
            fresh = true;
            tc::fence_before_sync();
            __syncthreads();
            MK_PH(9);
        }
        cur ^= 1;
    }
    g_flush(nflushed > 0, my_tiles > 0);
    tc::fence_before_sync();
    __syncthreads();
    MK_PH(9);                                             // G partial sums -> global
  }
    MK_PH_FLUSH(g_ph_bwd);
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// =============================================================================================================
// k_conv_bwd_pipe: the same computation as k_conv_bwd_tile, software pipelined
// =============================================================================================================
// k_conv_bwd_tile runs scatter -> MMA -> clear strictly one after the other (every thread waits for the 48 MMAs of a visit).
// Here the coefficient block Wt is double buffered and the roles are split:
//   16 worker warps  build Wt of visit i+1 (clear, rank-0 scatter, rank passes) while the tensor core consumes visit i; they
//                    run the dxh epilogue of tile t-1 after handing over the first block of tile t (two dxh accumulators);
//   MMA warp         per visit: G_b += Wt . xhat, dxh (+)= Wt^T . khat_b -- both B operands arrive K step by K step ...
//   ring warp        ... through a ring of bulk copies straight out of the tile-ordered node images and the kernel-block
//                    images (16 rows = 2 row groups of an image are contiguous), so neither image is resident in shared
//                    memory: that is what makes room for the second Wt buffer.
// The Jacobian reads xhat from the (L2 resident) global image, prefetched into registers before the accumulator is waited for.
constexpr int TP_WORK = 512;
constexpr int TP_THREADS = TP_WORK + 128;     // + two ring warps + two MMA warps (G group, dxh group)
constexpr int TP_SPS = 1;                      // K steps (of 16 rows) per ring stage
constexpr int TP_MAXSTAGES = 8;

__device__ __forceinline__ void tp_worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TP_WORK) : "memory"); }
__device__ __forceinline__ void tp_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

#ifdef MK_PHASE_CLOCKS
__device__ unsigned long long g_ph_bwdp[48];     // [0..15] worker thread 0, [16..31] ring lane, [32..47] MMA lane
#endif

__global__ void __launch_bounds__(TP_THREADS, 1) k_conv_bwd_pipe(const __grid_constant__ BwdTileArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar_cp[2], bar_wfull[2], bar_wfree[2], bar_dxrdy[2], bar_dxfree[2], bar_rfull[TP_MAXSTAGES],
        bar_rfree[TP_MAXSTAGES], bar_done;
    __shared__ uint32_t tslot;
    __shared__ BSeg s_seg[4][TILE_MAXSEG];
    __shared__ unsigned char s_lut[4][12];               // packed permutation codes (2 bits per j) per degree
    __shared__ float s_gmax;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    MK_PH_DECL(tid == 0 || tid == TP_WORK || tid == TP_WORK + 64)      // worker 0, first ring lane, G-MMA lane
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&bar_cp[i], 1); tc::mbar_init(&bar_wfull[i], 1); tc::mbar_init(&bar_wfree[i], 2);   // wfree: both MMA warps
            tc::mbar_init(&bar_dxrdy[i], 1); tc::mbar_init(&bar_dxfree[i], 1);
        }
        for (int i = 0; i < TP_MAXSTAGES; ++i) { tc::mbar_init(&bar_rfull[i], 1); tc::mbar_init(&bar_rfree[i], 1); }
        tc::mbar_init(&bar_done, 2);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(&tslot, 512);
    if (tid < 48) {
        s_lut[tid / 12][tid % 12] = c_perm_code[tid / 12][tid % 12];
    }
    if (tid >= 64 && tid < 64 + 4 * TILE_MAXSEG) {
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
    if (warp == 3) {                                      // max |coef| over the per-CTA values of k_coef_tile
        float gm = 0.f;
        for (int i = lane; i < (int)gridDim.x; i += 32) gm = fmaxf(gm, __ldg(a.amax + i));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
        if (lane == 0) s_gmax = gm;
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tslot;
    const TileWalk walk(a.order, a.order_grid, a.n_tiles);
    const int my_tiles = walk.cnt;
    const int NS = a.nstages;
    const uint32_t fgrp = (uint32_t)(a.Fk >> 3) * 128u;  // bytes of one 8-row group of an [R x Fk] image
    unsigned char* ring = smem + a.sm_ring;
    const int dxcol = a.nbl * a.cs;                       // dxh accumulators behind the G accumulators

    // TWO independent rings (single producer, single consumer each): ring 0 carries the node-image K steps to the G warp, ring 1
    // the kernel-block-image K steps to the dxh warp -- slots [0, NSh) and [NSh, 2 NSh).  (One shared ring with two consumers
    // running at their own pace aliases the mbarrier phase parity: a consumer a whole ring ahead of the other sees "complete".)
    const int NSh = NS / 2;
    if (warp == TP_WORK / 32 || warp == TP_WORK / 32 + 1) {
        // ================= ring warps =================
        if (lane == 0) {
            const bool isx = warp == TP_WORK / 32;                    // node image (true) / kernel-block image (false)
            const uint32_t s0 = isx ? 0u : (uint32_t)NSh;
            uint32_t q = 0;
            for (int wk = 0; wk < my_tiles; ++wk) {
                const int tile = walk.tile(wk);
                const int nn = __ldg(&a.meta[tile].nn);
                const int nkg = max(16, (nn + 15) & ~15) >> 4;
                for (int bi = 0; bi < a.nbl; ++bi) {
                    const int blk = a.blist[bi];
                    const int nk = isx ? nkg : (max(16, (a.tb.rows[blk] + 15) & ~15) >> 4);
                    const unsigned char* base = isx ? a.ximg + (size_t)tile * 2 * a.x_one : a.img + (size_t)blk * 2 * a.img_one;
                    const uint32_t one = (uint32_t)(isx ? a.x_one : a.img_one);
                    for (int ks = 0; ks < nk; ++ks, ++q) {
                        const uint32_t slot = s0 + q % (uint32_t)NSh, use = q / (uint32_t)NSh;
                        MK_PH(0);
                        tc::mbar_wait(&bar_rfree[slot], (use & 1u) ^ 1u);
                        MK_PH(1);
                        unsigned char* dst = ring + (size_t)slot * a.stage_bytes;
                        const unsigned char* src = base + (size_t)ks * 2 * fgrp;
                        mbar_expect_tx(&bar_rfull[slot], 4u * fgrp);
                        bulk_g2s(dst, src, 2u * fgrp, &bar_rfull[slot]);
                        bulk_g2s(dst + 2u * fgrp, src + one, 2u * fgrp, &bar_rfull[slot]);
                    }
                }
            }
        }
    } else if (warp == TP_WORK / 32 + 2 || warp == TP_WORK / 32 + 3) {
        // ================= MMA warps: warp G issues G_b += Wt . xhat, warp X issues dxh (+)= Wt^T . khat -- one thread needs ~100
        // cycles per tcgen05.mma, two issuing threads keep the tensor pipe busy; bar_wfree / bar_done count both =================
        const bool isG = warp == TP_WORK / 32 + 2;
        if (lane == 0) {
            uint32_t q = 0, vis = 0;
            const uint32_t idesc_g = tc::idesc_f16(128, a.Fk, 0, 1);  // A = Wt K-major (K = node), B = xhat MN-major (N = feature)
            const uint32_t idesc_x = tc::idesc_f16(128, a.Fk, 1, 1);  // A = Wt MN-major (M = node, K = row), B = khat MN-major
            for (int wk = 0; wk < my_tiles; ++wk) {
                const int tile = walk.tile(wk);
                const int nn = __ldg(&a.meta[tile].nn);
                const int nkg = max(16, (nn + 15) & ~15) >> 4;
                const uint32_t par_t = (uint32_t)wk & 1u;
                const uint32_t dX = tmem + (uint32_t)(dxcol + (int)par_t * a.cs);
                for (int bi = 0; bi < a.nbl; ++bi, ++vis) {
                    const int blk = a.blist[bi];
                    const int nkx = max(16, (a.tb.rows[blk] + 15) & ~15) >> 4;
                    const uint32_t buf = vis & 1u;
                    const uint32_t whi = tc::smem_u32(smem + a.sm_wt + (size_t)buf * 2 * WT_ONE), wlo = whi + WT_ONE;
                    MK_PH(0);
                    tc::mbar_wait(&bar_wfull[buf], (vis >> 1) & 1u);
                    MK_PH(1);                                             // MMA: waiting for the workers' Wt
                    if (!isG && bi == 0) tc::mbar_wait(&bar_dxfree[par_t], (((uint32_t)wk >> 1) & 1u) ^ 1u);
                    tc::fence_after_sync();
                    if (isG) {
                        const uint32_t dG = tmem + (uint32_t)(bi * a.cs);
                        for (int ks = 0; ks < nkg; ++ks, ++q) {           // G_bi += Wt . xhat   (K = nodes, 16 per step)
                            const uint32_t slot = q % (uint32_t)NSh, use = q / (uint32_t)NSh;
                            MK_PH(3);
                            tc::mbar_wait(&bar_rfull[slot], use & 1u);
                            MK_PH(2);                                     // MMA: waiting for a ring stage
                            const uint32_t st = tc::smem_u32(ring + (size_t)slot * a.stage_bytes);
                            const uint64_t dAh = tc::smem_desc(whi + ks * 256u, 128u, 2048u), dAl = tc::smem_desc(wlo + ks * 256u, 128u, 2048u);
                            const uint64_t dBh = tc::smem_desc(st, fgrp, 128u), dBl = tc::smem_desc(st + 2u * fgrp, fgrp, 128u);
                            tc::umma_f16(dG, dAh, dBh, idesc_g, (wk == 0 && ks == 0) ? 0u : 1u);
                            tc::umma_f16(dG, dAl, dBh, idesc_g, 1u);
                            tc::umma_f16(dG, dAh, dBl, idesc_g, 1u);
                            tc::umma_commit(&bar_rfree[slot]);
                        }
                    } else {
                        for (int ks = 0; ks < nkx; ++ks, ++q) {           // dxh (+)= Wt^T . khat   (K = kernel rows, 16 per step)
                            const uint32_t slot = (uint32_t)NSh + q % (uint32_t)NSh, use = q / (uint32_t)NSh;
                            MK_PH(3);
                            tc::mbar_wait(&bar_rfull[slot], use & 1u);
                            MK_PH(2);
                            const uint32_t st = tc::smem_u32(ring + (size_t)slot * a.stage_bytes);
                            const uint64_t dAh = tc::smem_desc(whi + ks * 4096u, 2048u, 128u), dAl = tc::smem_desc(wlo + ks * 4096u, 2048u, 128u);
                            const uint64_t dBh = tc::smem_desc(st, fgrp, 128u), dBl = tc::smem_desc(st + 2u * fgrp, fgrp, 128u);
                            tc::umma_f16(dX, dAh, dBh, idesc_x, (bi == 0 && ks == 0) ? 0u : 1u);
                            tc::umma_f16(dX, dAl, dBh, idesc_x, 1u);
                            tc::umma_f16(dX, dAh, dBl, idesc_x, 1u);
                            tc::umma_commit(&bar_rfree[slot]);
                        }
                    }
                    tc::umma_commit(&bar_wfree[buf]);                     // (both warps) Wt buffer free once these MMAs have read it
                    if (!isG && bi == a.nbl - 1) tc::umma_commit(&bar_dxrdy[par_t]);
                    MK_PH(3);
                }
            }
            tc::umma_commit(&bar_done);
        }
    } else {
        // ================= workers =================
        float* red = reinterpret_cast<float*>(smem + a.sm_red);     // Jacobian reduction scratch
        // power-of-two scale: |alpha * chi * g| / scale <= 2^10
        float scale, rscale;
        {
            const float gm = fmaxf(s_gmax, 1e-30f);
            int e;
            frexpf(gm, &e);                                   // gm < 2^e
            scale = ldexpf(1.0f, e - 10);
            rscale = ldexpf(1.0f, 10 - e);
        }
        const int q = warp & 3, cpart = warp >> 2;            // TMEM lane quadrant / 32-column part of this warp
        auto issue_tile = [&](int wk) {                       // thread 0: metadata, coefficients, arg-max codes of tile wk
            const int tile = walk.tile(wk);
            const int np = tb_tile_pairs(a, tile);
            unsigned char* buf = smem + a.sm_buf + (size_t)(wk & 1) * a.tbuf_bytes;
            uint64_t* bar = &bar_cp[wk & 1];
            const uint32_t cb = (uint32_t)((np * 4 + 15) & ~15), ab = (uint32_t)((np + 15) & ~15);
            mbar_expect_tx(bar, (uint32_t)sizeof(TileMetaG) + cb + ab);
            bulk_g2s(buf, a.meta + tile, (uint32_t)sizeof(TileMetaG), bar);
            if (np > 0) {
                bulk_g2s(buf + a.tb_a, a.coefT + (size_t)tile * a.stride, cb, bar);
                bulk_g2s(buf + a.tb_am, a.amT + (size_t)tile * a.stride_am, ab, bar);
            }
        };
        // dxh epilogue of one tile (lane = node, 32 columns per warp): partial of the first launch / Jacobian + grad_x
        auto epilogue = [&](int wk_t, int tile, int t0, int nn) {
            const uint32_t par = (uint32_t)wk_t & 1u;
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
            uint4 xh4[4], xl4[4];                          // xhat = hi + lo from the global node image (prefetched)
#pragma unroll
            for (int c = 0; c < 4; ++c) { xh4[c] = make_uint4(0u, 0u, 0u, 0u); xl4[c] = make_uint4(0u, 0u, 0u, 0u); }
            if (a.last && rowok) {
                nrm = a.xnorm[t0 + v];
                if (colok) {
                    const unsigned char* Xhi = a.ximg + (size_t)tile * 2 * a.x_one;
                    const unsigned char* Xlo = Xhi + a.x_one;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (8 * c < nf) {
                            const uint32_t off = tc::il_off(v, f0 + 8 * c, a.Fk);
                            xh4[c] = __ldg(reinterpret_cast<const uint4*>(Xhi + off));
                            xl4[c] = __ldg(reinterpret_cast<const uint4*>(Xlo + off));
                        }
                    }
                }
            }
            tc::mbar_wait(&bar_dxrdy[par], ((uint32_t)wk_t >> 1) & 1u);
            tc::fence_after_sync();
            if (colok) {
                uint32_t u[32];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(dxcol + (int)par * a.cs + f0);
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
            float xh[32];
            float dot = 0.f;
            if (a.last) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const __half2* hh = reinterpret_cast<const __half2*>(&xh4[c]);
                    const __half2* ll = reinterpret_cast<const __half2*>(&xl4[c]);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const float2 fh = __half22float2(hh[t]), fl = __half22float2(ll[t]);
                        xh[8 * c + 2 * t] = fh.x + fl.x;
                        xh[8 * c + 2 * t + 1] = fh.y + fl.y;
                    }
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) dot = fmaf(dv[i], xh[i], dot);
                red[cpart * 128 + v] = dot;
            }
            tc::fence_before_sync();
            tp_worker_sync();                             // every worker has drained its part of the accumulator
            if (tid == 0) tp_arrive(&bar_dxfree[par]);
            if (!a.last) {
                if (rowok && colok) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        if (i < nf) __stcg(reinterpret_cast<float4*>(sp + i), make_float4(dv[i], dv[i + 1], dv[i + 2], dv[i + 3]));
                }
            } else {
                // chain rule through xhat = x / max(|x|, eps):  gx = (g - (xhat . g) xhat) / |x|
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
                tp_worker_sync();                         // `red` is free again
            }
        };

        if (tid == 0) {
            if (my_tiles > 0) issue_tile(0);
            if (my_tiles > 1) issue_tile(1);
        }
        MK_PH(0);
        uint32_t vis = 0;
        int t0p = 0, nnp = 0, tilep = 0;
        for (int wk = 0; wk < my_tiles; ++wk) {
            const int tile = walk.tile(wk);
            const unsigned char* tbuf = smem + a.sm_buf + (size_t)(wk & 1) * a.tbuf_bytes;
            const TileMetaG& m = *reinterpret_cast<const TileMetaG*>(tbuf);
            const float* a_s = reinterpret_cast<const float*>(tbuf + a.tb_a);
            const unsigned char* am_s = tbuf + a.tb_am;
            tc::mbar_wait(&bar_cp[wk & 1], ((uint32_t)wk >> 1) & 1u);
            MK_PH(1);                                         // wait for the tile's metadata, coefficients, arg-max codes
            const int t0c = m.t0, nnc = m.nn;                 // latched: the buffer is refilled at the end of the tile
            int doff[4];                                      // first pair of degree d in the tile-ordered arrays
            doff[0] = 0;
#pragma unroll
            for (int d = 1; d < 4; ++d) doff[d] = doff[d - 1] + m.cnt[d - 1] * a.L[d - 1];
            for (int bi = 0; bi < a.nbl; ++bi, ++vis) {
                const int blk = a.blist[bi];
                const int nseg = a.tb.nseg[blk];
                unsigned char* wt = smem + a.sm_wt + (size_t)(vis & 1u) * 2 * WT_ONE;
                int abase[TILE_MAXSEG];
                for (int si = 0; si < nseg; ++si) abase[si] = doff[s_seg[bi][si].d - 1] + s_seg[bi][si].k0;
                tc::mbar_wait(&bar_wfree[vis & 1u], ((vis >> 1) & 1u) ^ 1u);   // the MMAs of visit vis - 2 have read this buffer
                MK_PH(2);
                for (int i = tid * 16; i < 2 * WT_ONE; i += TP_WORK * 16) *reinterpret_cast<uint4*>(wt + i) = make_uint4(0, 0, 0, 0);
                tp_worker_sync();
                MK_PH(3);                                     // Wt clear
                // ---- rank 0: one thread per (node, kernel) pair -- centre entry, collision-free support entries ----
                for (int si = 0; si < nseg; ++si) {
                    const BSeg sg = s_seg[bi][si];
                    const int np = m.cnt[sg.d - 1] * sg.nk;
                    for (int p = tid; p < np; p += TP_WORK) {
                        const int ni = (int)(((float)p + 0.5f) * sg.rnk);
                        const int kl = p - ni * sg.nk;
                        const int nl_ = m.list[sg.d - 1][ni];
                        const int pi = abase[si] + ni * sg.L + kl;
                        const float av = a_s[pi] * rscale;
                        const uint32_t code = s_lut[sg.d - 1][am_s[pi] & 0x7f];
                        const uint32_t nw = m.nl[nl_];
                        const uint32_t cr = m.cr[nl_];
                        wt_store(wt, sg.rowbase + sg.d * sg.nk + kl, nl_, av * sg.beta);
                        const float as = av * sg.alpha;           // the same value goes to all d support entries: split it once
                        const __half ah = __float2half_rn(as);
                        const __half al = __float2half_rn(as - __half2float(ah));
                        for (int j = 0; j < sg.d; ++j) {
                            if (((cr >> (2 * j)) & 3u) == 0u)
                                wt_store_hl(wt, sg.rowbase + (int)((code >> (2 * j)) & 3u) * sg.nk + kl, (int)((nw >> (8 * j)) & 0xffu), ah, al);
                        }
                    }
                }
                MK_PH(4);                                     // rank-0 scatter
                // ---- collision chains: one thread per (chain, kernel) adds the followers in in-edge order after ONE barrier ----
                {
                    bool more = false;
                    for (int si = 0; si < nseg; ++si) {
                        const int d = s_seg[bi][si].d;
                        if (m.choff[d] > m.choff[d - 1]) more = true;
                    }
                    if (more) {
                        tp_worker_sync();
                        for (int si = 0; si < nseg; ++si) {
                            const BSeg sg = s_seg[bi][si];
                            const int c0 = m.choff[sg.d - 1];
                            const int ni_ = (m.choff[sg.d] - c0) * sg.nk;
                            for (int p = tid; p < ni_; p += TP_WORK) {
                                const int ci = (int)(((float)p + 0.5f) * sg.rnk);
                                const int kl = p - ci * sg.nk;
                                const uint32_t ch = m.chains[c0 + ci];
                                const int nf = (int)((ch >> 27) & 3u) + 1;
                                for (int f = 0; f < nf; ++f) {
                                    const int ent = (int)((ch >> (9 * f)) & 0x1ffu);
                                    const int nl_ = ent >> 2, j = ent & 3;
                                    const int pi = abase[si] + m.lidx[nl_] * sg.L + kl;
                                    const int s = (s_lut[sg.d - 1][am_s[pi] & 0x7f] >> (2 * j)) & 3;
                                    wt_add(wt, sg.rowbase + s * sg.nk + kl, (int)((m.nl[nl_] >> (8 * j)) & 0xffu), (a_s[pi] * rscale) * sg.alpha);
                                }
                            }
                        }
                    }
                }
                tc::fence_async_smem();
                tp_worker_sync();
                if (tid == 0) tp_arrive(&bar_wfull[vis & 1u]);
                MK_PH(5);                                     // ranks 1..3 + hand-over
                // the MMAs of the previous tile's last block ran while this block was scattered: its dxh is (about) ready
                if (bi == 0 && wk > 0) { epilogue(wk - 1, tilep, t0p, nnp); MK_PH(6); }
            }
            t0p = t0c; nnp = nnc; tilep = tile;
            // this tile buffer is free (the hand-over barrier above ended the last reads): fetch the tile after the next
            if (tid == 0 && wk + 2 < my_tiles) issue_tile(wk + 2);
        }
        if (my_tiles > 0) { epilogue(my_tiles - 1, tilep, t0p, nnp); MK_PH(6); }
        // ---- kernel-parameter partial sums of this CTA: node-attribute part of every row of this launch's blocks ----
        if (my_tiles > 0) { tc::mbar_wait(&bar_done, 0u); tc::fence_after_sync(); }
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
#pragma unroll
                for (int i = 0; i < 32; ++i) u[i] = 0u;
                if (my_tiles > 0) {
                    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(bi * a.cs + f0);
                    tc::tmem_ld16(taddr, u);
                    if (f0 + 16 < a.Fk) tc::tmem_ld16(taddr + 16, u + 16);
                    tc::tmem_ld_wait();
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(u[i]));
                if (d > 0) {
                    const int rows_x = (d + 1) * L;
                    float* part = a.partials + a.part_off[d - 1] + ((size_t)blockIdx.x * rows_x + (size_t)slot * L + kk) * a.FW;
                    const float sc2 = my_tiles > 0 ? scale : 0.f;     // a CTA without tiles never ran an MMA: zero partial sums
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        if (f0 + i + 4 <= a.Fp)
                            st4(part + f0 + i, make_float4(__uint_as_float(u[i]) * sc2, __uint_as_float(u[i + 1]) * sc2,
                                                           __uint_as_float(u[i + 2]) * sc2, __uint_as_float(u[i + 3]) * sc2));
                    }
                }
            }
        }
        MK_PH(7);                                             // G partial sums -> global
    }
    tc::fence_before_sync();
    __syncthreads();
#ifdef MK_PHASE_CLOCKS
    MK_PH_FLUSH(g_ph_bwdp + (tid == 0 ? 0 : tid == TP_WORK ? 16 : 32));      // ring clocks: the first ring warp
#endif
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

// tile-ordered coefficient layout: floats per tile / bytes per tile of the arg-max codes
static void coef_strides(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, int* stride, int* stride_am) {
    int64_t st = 0;
    for (int d = 0; d < 4; ++d) st += (int64_t)plan->tile_max_deg[d] * layer->L[d];
    *stride = (int)((st + 3) / 4 * 4);
    *stride_am = (*stride + 15) / 16 * 16;
}

bool tile_bwd_ok(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer) {
    if (!tile_plan_ok_b(plan) || !layer->tile_img || !tile_layer_ok(layer)) return false;
    TileBlocks tb;
    if (!tb.build(layer->L) || tb.nb > 4) return false;                 // at most two launches of two blocks
    int rows = 0;
    for (int d = 0; d < 4; ++d) rows += layer->L[d] * (d + 1);
    return rows <= CT_MAXR * CT_THREADS;
}

int tile_argmax_stride(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer) {
    int stride, stride_am;
    coef_strides(plan, layer, &stride, &stride_am);
    return stride_am;
}

// floats of the `coef` scratch the tile path needs (tile-ordered coefficients + arg-max codes behind them); 0 = not eligible
int64_t tile_bwd_coef_floats(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer) {
    if (!tile_bwd_ok(plan, layer)) return 0;
    int stride, stride_am;
    coef_strides(plan, layer, &stride, &stride_am);
    return (int64_t)plan->n_tiles * stride + ((int64_t)plan->n_tiles * stride_am + 3) / 4 + 4;
}

// k_coef_tile for wide layers (conv_bwd_wide.cu): coefficients / arg-max codes in blocked tile order, max |coef| per CTA, and the
// bond-attribute sums as `grid` partial copies in bondP -- [degree][CTA][(d + 1) L_d rows][8 floats] (centre rows zero).
// Returns the shared-memory bytes it needs (<= 0: does not fit) when `do_launch` is false.
int64_t launch_coef_wide(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* grad, int32_t ldg, int32_t grad_mode,
                         const uint8_t* argmax, const int64_t scoff[4], float* coefT, uint8_t* amT, int stride, int stride_am,
                         float* bondP, float* amax, int grid, bool do_launch, cudaStream_t st) {
    CoefTileArgs c;
    memset(&c, 0, sizeof(c));
    WideBlocks wb;
    if (!wb.build(layer->L)) return -1;
    c.meta = reinterpret_cast<const TileMetaG*>(plan->tile_meta); c.ehat_node = plan->ehat_node; c.n_tiles = plan->n_tiles;
    c.order = plan->tile_grid == grid ? plan->tile_order : nullptr;
    c.order_grid = c.order ? grid : 0;
    int64_t po = 0;
    for (int d = 0; d < 4; ++d) {
        c.L[d] = layer->L[d]; c.koff[d] = layer->koff[d]; c.scoff[d] = scoff[d];
        c.part_off[d] = po;
        po += (int64_t)grid * (d + 2) * layer->L[d] * EP;
        int need = 0;
        for (int b = 0; b < wb.nb; ++b) if (wb.d[b] == d + 1) ++need;
        c.wb_base[d] = need ? layer->L[d] / need : 0;
        c.wb_rem[d] = need ? layer->L[d] % need : 0;
    }
    c.wide = 1;
    c.grad = grad; c.ldg = ldg; c.grad_mode = grad_mode; c.vec = 1;
    c.grad_floats = (long long)plan->N * ldg;
    c.argmax = argmax;
    c.coefT = coefT; c.amT = amT; c.stride = stride; c.stride_am = stride_am;
    c.amT_in = nullptr;
    c.partials = bondP; c.FW = EP; c.Fp = 0;              // bond columns only: rows of 8 floats
    c.amax = amax;
    for (int div = 1;; ++div) {
        int items = 0;
        for (int d = 0; d < 4; ++d) {
            c.nch[d] = std::max(1, std::min(8, ((plan->tile_max_deg[d] + CT_CHUNK - 1) / CT_CHUNK + div - 1) / div));
            items += (d + 1) * layer->L[d] * c.nch[d];
        }
        if (items <= CT_MAXR * CT_THREADS) break;
        if (div >= 64) return -1;
    }
    {
        int64_t o = (sizeof(TileMetaG) + 127) / 128 * 128;
        c.sm_grad = (int)o;
        c.sm_eh = (int)o; o += (int64_t)TILE_ESLOTS * EP * 4;
        c.sm_am = (int)o;
        c.buf_bytes = (int)o;
        c.sm_coef = (int)(2 * o);
        c.sm_inv = c.sm_coef + (int)(((int64_t)stride * 4 + 127) / 128 * 128);
    }
    // the chunk sums are combined through shared memory at the end: 32 bytes per item
    int items = 0;
    for (int d = 0; d < 4; ++d) items += (d + 1) * layer->L[d] * c.nch[d];
    const int64_t smem_c = std::max<int64_t>(c.sm_inv + ((int64_t)stride + 127) / 128 * 128, (int64_t)items * 32 + 128);
    static int s_budget = 0;
    if (!s_budget) s_budget = device_max_smem_optin();
    if (smem_c > s_budget - 2048) return -1;
    if (!do_launch) return smem_c;
    int64_t& s_attr_c = g_coef_attr_dev[device_index()];
    if (smem_c > s_attr_c) {
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_coef_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        s_attr_c = smem_c;
    }
    count_launches(1);
    ProfScope prof("coef_tile", st);
    k_coef_tile<<<grid, CT_THREADS, smem_c, st>>>(c);
    MK_CHECK_CUDA(cudaGetLastError());
    return smem_c;
}

// returns 1 if launched, 0 if not eligible, <0 on error.  part_off / ncta describe the partial copies for k_param_finalize.
int launch_conv_bwd_tile(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                         const float* xnorm, const void* ximg, const float* grad, int32_t ldg, int32_t grad_mode,
                         const uint8_t* argmax, const int64_t scoff[4], float* coef, float* partials, float* scratch,
                         float* grad_x, int32_t ldgx, int64_t part_off[4], int ncta[4], int64_t* part_total, bool do_launch,
                         const uint8_t* argmax_tile, cudaStream_t st) {
    (void)x; (void)ldx;
    if (!ximg || !coef || !tile_bwd_ok(plan, layer)) return 0;
    static int s_budget = 0;
    if (!s_budget) {
        s_budget = device_max_smem_optin();
        MK_REQUIRE(s_budget > 0, "conv_bwd_tile: no CUDA device");
    }
    BwdTileArgs a;
    if (!a.tb.build(layer->L)) return 0;
    a.xnorm = xnorm;
    a.F = layer->F; a.Fp = layer->Fp; a.Fk = tile_fk(layer->Fp);
    a.meta = reinterpret_cast<const TileMetaG*>(plan->tile_meta);
    a.ximg = reinterpret_cast<const unsigned char*>(ximg);
    a.n_tiles = plan->n_tiles;
    const int grid = tile_bwd_grid(plan);
    a.order = plan->tile_grid == grid ? plan->tile_order : nullptr;
    a.order_grid = a.order ? grid : 0;
    a.FW = layer->Fp + EP;
    int64_t po = 0, rows_all = 0;
    for (int d = 0; d < 4; ++d) {
        a.L[d] = layer->L[d];
        a.packed[d] = layer->packed[d];
        a.part_off[d] = part_off[d] = po;
        ncta[d] = layer->L[d] > 0 ? grid : 0;
        po += (int64_t)ncta[d] * (d + 2) * layer->L[d] * a.FW;
        rows_all += (int64_t)(d + 2) * layer->L[d];
    }
    *part_total = po;
    float* amax = partials + po + 2 * rows_all + 16;  // [grid] floats behind k_param_finalize's Q scratch
    MK_REQUIRE(grid <= 512, "conv_bwd_tile: %d CTAs (the per-CTA max |coef| array holds 512)", grid);
    a.img = reinterpret_cast<const unsigned char*>(layer->tile_img);
    a.img_one = tile_img_one(a.Fk);
    a.x_one = tile_img_one(a.Fk);
    a.amax = amax;
    a.partials = partials;
    a.scratch = scratch;
    a.gx = grad_x; a.ldgx = ldgx;
    int stride, stride_am;
    coef_strides(plan, layer, &stride, &stride_am);
    a.coefT = coef;
    // arg-max bytes in tile order: the forward's copy if the caller kept one, else k_coef_tile writes its own behind coefT
    a.amT = argmax_tile ? argmax_tile : reinterpret_cast<const uint8_t*>(coef + (size_t)plan->n_tiles * stride);
    a.stride = stride; a.stride_am = stride_am;
    a.buf_bytes = (int)((sizeof(TileMetaG) + 127) / 128 * 128);
    // TMEM: G of block bi at bi * gstride, dxh behind them.  A layer whose feature width lets (blocks + 1) accumulators fit
    // the 512 columns runs ONE launch over all blocks; else two launches of two blocks, the first handing its partial dxh
    // to the second through `scratch`.
    static int s_merge = -1;
    if (s_merge < 0) { const char* e = getenv("MOLKGNN_BWD_MERGE"); s_merge = (e && e[0] == '0') ? 0 : 1; }
    const bool one_launch = a.tb.nb <= 2 || (s_merge && (a.tb.nb + 1) * a.Fk <= 512);
    const int nlaunch = one_launch ? 1 : 2;
    const int nbl_max = one_launch ? a.tb.nb : 2;
    int64_t off = 0;
    a.sm_x = (int)off; off += 2 * (int64_t)a.x_one;
    a.sm_wt = (int)off; off += 2 * (int64_t)WT_ONE;
    a.sm_buf = (int)off; off += 2 * (int64_t)a.buf_bytes;
    a.a_bytes = (int)(((int64_t)stride * 4 + 127) / 128 * 128);
    a.am_bytes = (int)(((int64_t)stride_am + 127) / 128 * 128);
    // the static shared memory of the kernel is < 1 KB: dynamic budget = opt-in maximum - 1 KB
    const int64_t budget = (int64_t)s_budget - 1024;
    static int s_dbuf = -1;
    if (s_dbuf < 0) { const char* e = getenv("MOLKGNN_BWD_DBUF"); s_dbuf = (e && e[0] == '0') ? 0 : 1; }
    const int64_t fixed = off + 4 * 128 * 4 + 2 * (int64_t)a.img_one;
    // double buffered coefficient / arg-max arrays (the next tile's arrive while this tile is scattered) when they fit
    a.dbuf = (s_dbuf && fixed + 2 * ((int64_t)a.a_bytes + a.am_bytes) <= budget) ? 1 : 0;
    a.sm_a = (int)off; off += (int64_t)(a.dbuf + 1) * a.a_bytes;
    a.sm_am = (int)off; off += (int64_t)(a.dbuf + 1) * a.am_bytes;
    a.sm_red = (int)off; off += 4 * 128 * 4;
    a.sm_img = (int)off; off += 2 * (int64_t)a.img_one;
    if (off > budget) return 0;
    a.nimg = 1;                      // as many image buffers as fit (all blocks resident if possible)
    while (a.nimg < nbl_max && off + 2 * (int64_t)a.img_one <= budget) { ++a.nimg; off += 2 * (int64_t)a.img_one; }
    // k_coef_tile: two tile buffers (metadata + gradient rows + bond rows) and the pair arrays of the current tile
    CoefTileArgs c;
    c.vec = ((int64_t)plan->N * ldg) % 4 == 0 ? 4 : ((int64_t)plan->N * ldg) % 2 == 0 ? 2 : 1;
    {
        int64_t o = (sizeof(TileMetaG) + 127) / 128 * 128;
        c.sm_grad = (int)o; o += (((int64_t)TNODES * ldg + 8) * 4 + 127) / 128 * 128;
        c.sm_eh = (int)o; o += (int64_t)TILE_ESLOTS * EP * 4;
        c.sm_am = (int)o; o += ((int64_t)stride_am + 127) / 128 * 128;
        c.buf_bytes = (int)o;
        c.sm_coef = (int)(2 * o);
        c.sm_inv = c.sm_coef + (int)(((int64_t)stride * 4 + 127) / 128 * 128);
    }
    const int64_t smem_c = c.sm_inv + ((int64_t)stride + 127) / 128 * 128;
    if (smem_c > s_budget - 2048) return 0;
    // pipelined kernel (k_conv_bwd_pipe): two Wt buffers, two tile buffers [meta | coef | arg-max], ring of K-step stages
    static int s_pipe = -1;
    // opt-in (MOLKGNN_BWD_PIPE=1): measured equal to the serial kernel (0.653 vs 0.643 ms per step at 4096 molecules) -- the
    // tcgen05.mma of both kernels read their operands from shared memory (SS mode: 7.5 KB per 128 x 112 x 16 MMA = the whole
    // 128 B/cycle of the SM), so overlapping the scatter with the MMAs only makes both slower
    if (s_pipe < 0) { const char* e = getenv("MOLKGNN_BWD_PIPE"); s_pipe = (e && e[0] == '1') ? 1 : 0; }
    int64_t poff = 0;
    bool use_pipe = s_pipe != 0;
    BwdTileArgs pa = a;
    {
        pa.cs = a.Fk <= 32 ? 32 : a.Fk <= 64 ? 64 : 128;
        if ((nbl_max + 2) * pa.cs > 512) use_pipe = false;
        pa.sm_wt = (int)poff; poff += 2 * 2 * (int64_t)WT_ONE;
        pa.tb_a = (int)((sizeof(TileMetaG) + 127) / 128 * 128);
        pa.tb_am = pa.tb_a + (int)(((int64_t)stride * 4 + 127) / 128 * 128);
        pa.tbuf_bytes = pa.tb_am + (int)(((int64_t)stride_am + 127) / 128 * 128);
        pa.sm_buf = (int)poff; poff += 2 * (int64_t)pa.tbuf_bytes;
        pa.sm_red = (int)poff; poff += 4 * 128 * 4;
        pa.sm_ring = (int)poff;
        pa.stage_bytes = TP_SPS * 4 * (a.Fk >> 3) * 128;
        const int64_t room = (int64_t)s_budget - 2048 - poff;
        pa.nstages = (int)std::min<int64_t>(TP_MAXSTAGES, room / pa.stage_bytes);
        if (pa.nstages < 6) use_pipe = false;                 // two rings of >= 3 stages
        poff += (int64_t)std::max(pa.nstages, 0) * pa.stage_bytes;
    }
    static int s_gflush = -1;
    if (s_gflush < 0) { const char* e = getenv("MOLKGNN_BWD_GFLUSH"); s_gflush = e ? std::max(1, atoi(e)) : 6; }
    a.gflush = s_gflush;
    if ((plan->n_tiles + grid - 1) / grid > s_gflush) use_pipe = false;      // the pipelined variant keeps ONE chain per CTA
    if (!do_launch) return 1;
    MK_REQUIRE((reinterpret_cast<uintptr_t>(grad) & 15) == 0 && (reinterpret_cast<uintptr_t>(coef) & 15) == 0,
               "conv_bwd_tile: grad and coef must be 16-byte aligned");
    static int64_t s_attr_dev[16] = {0};
    int64_t& s_attr = s_attr_dev[device_index()];
    int64_t& s_attr_c = g_coef_attr_dev[device_index()];   // function attributes are per device
    if (off > s_attr) {
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_conv_bwd_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off));
        s_attr = off;
    }
    if (smem_c > s_attr_c) {
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_coef_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        s_attr_c = smem_c;
    }
    static int64_t s_attr_p_dev[16] = {0};
    int64_t& s_attr_p = s_attr_p_dev[device_index()];
    if (use_pipe && poff > s_attr_p) {
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_conv_bwd_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)poff));
        s_attr_p = poff;
    }
    {
        c.meta = a.meta; c.ehat_node = plan->ehat_node; c.n_tiles = plan->n_tiles;
        c.order = a.order; c.order_grid = a.order_grid;
        for (int d = 0; d < 4; ++d) {
            c.L[d] = layer->L[d]; c.koff[d] = layer->koff[d]; c.scoff[d] = scoff[d];
            c.part_off[d] = part_off[d];
        }
        c.grad = grad; c.ldg = ldg; c.grad_mode = grad_mode;
        c.grad_floats = (long long)plan->N * ldg;
        c.argmax = argmax;
        c.coefT = coef; c.stride = stride; c.stride_am = stride_am;
        c.amT_in = argmax_tile; c.amT = argmax_tile ? nullptr : const_cast<uint8_t*>(a.amT);
        c.partials = partials; c.FW = a.FW; c.Fp = layer->Fp;
        c.amax = amax;
        c.wide = 0;
        for (int d = 0; d < 4; ++d) { c.wb_base[d] = 0; c.wb_rem[d] = 0; }
        // node chunks per degree: as many as the fullest tile needs, fewer if the items would exceed the threads' capacity
        for (int div = 1;; ++div) {
            int items = 0;
            for (int d = 0; d < 4; ++d) {
                c.nch[d] = std::max(1, std::min(8, ((plan->tile_max_deg[d] + CT_CHUNK - 1) / CT_CHUNK + div - 1) / div));
                items += (d + 1) * layer->L[d] * c.nch[d];
            }
            if (items <= CT_MAXR * CT_THREADS || div >= 8) break;
        }
        count_launches(1);
        ProfScope prof("coef_tile", st);
        k_coef_tile<<<grid, CT_THREADS, smem_c, st>>>(c);
        MK_CHECK_CUDA(cudaGetLastError());
    }
    MK_REQUIRE(nlaunch == 1 || scratch, "conv_bwd_tile: scratch is required for layers with more than two kernel blocks");
    // Kernel blocks of the layer in (at most) two groups: {0, 1} first (base model: the two degree-4 blocks, whose partial dxh touches
    // the fewest rows), then {2, 3}.  k_conv_bwd_tile runs both groups as two PASSES of one launch (MOLKGNN_BWD_FUSE=0 and the
    // pipelined kernel: two launches).
    int g_nbl[2] = {0, 0}, g_blist[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    if (one_launch) {
        g_nbl[0] = a.tb.nb;
        for (int i = 0; i < a.tb.nb; ++i) g_blist[0][i] = i;
        a.gstride = a.tb.nb <= 2 ? 128 : a.Fk;
        a.dxcol = a.tb.nb <= 2 ? 256 : a.tb.nb * a.Fk;
    } else {
        g_nbl[0] = 2; g_blist[0][0] = 0; g_blist[0][1] = 1;
        g_nbl[1] = a.tb.nb - 2; g_blist[1][0] = 2; g_blist[1][1] = a.tb.nb == 3 ? 2 : 3;
        a.gstride = 128; a.dxcol = 256;
    }
    a.dmask_first = 0;
    for (int i = 0; i < g_nbl[0]; ++i)
        for (int si = 0; si < a.tb.nseg[g_blist[0][i]]; ++si) a.dmask_first |= 1 << a.tb.seg[g_blist[0][i]][si].d;
    static int s_fuse = -1;
    if (s_fuse < 0) { const char* e = getenv("MOLKGNN_BWD_FUSE"); s_fuse = (e && e[0] == '0') ? 0 : 1; }
    const bool fuse = nlaunch == 2 && s_fuse && !use_pipe;
    const int nl_eff = fuse ? 1 : nlaunch;
    for (int l = 0; l < nl_eff; ++l) {
        a.npass = fuse ? 2 : 1;
        for (int ps = 0; ps < 2; ++ps) {
            const int g = fuse ? ps : l;
            a.p_nbl[ps] = ps < a.npass ? g_nbl[g] : 0;
            for (int i = 0; i < 4; ++i) a.p_blist[ps][i] = ps < a.npass ? g_blist[g][i] : 0;
        }
        a.nbl = g_nbl[l];
        for (int i = 0; i < 4; ++i) a.blist[i] = g_blist[l][i];
        a.first = fuse || l == 0; a.last = fuse || l == nlaunch - 1;
        count_launches(1);
        ProfScope prof("conv_bwd_tile", st);
        if (use_pipe) {
            pa.nbl = a.nbl; pa.first = a.first; pa.last = a.last;
            for (int i = 0; i < 4; ++i) pa.blist[i] = a.blist[i];
            k_conv_bwd_pipe<<<grid, TP_THREADS, poff, st>>>(pa);
        } else {
            k_conv_bwd_tile<<<grid, TB_THREADS, off, st>>>(a);
        }
        MK_CHECK_CUDA(cudaGetLastError());
    }
    return 1;
}

}  // namespace mk


#ifdef MK_PHASE_CLOCKS
extern "C" int molkgnn_debug_phase_clocks_coef(unsigned long long* out16) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out16, mk::g_ph_coef, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
    unsigned long long z[16] = {0};
    return cudaMemcpyToSymbol(mk::g_ph_coef, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
// profiling build only: read (and clear) the accumulated phase clocks of k_conv_bwd_tile
extern "C" int molkgnn_debug_phase_clocks_bwd(unsigned long long* out16) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out16, mk::g_ph_bwd, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
    unsigned long long z[16] = {0};
    return cudaMemcpyToSymbol(mk::g_ph_bwd, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif
#ifdef MK_PHASE_CLOCKS
extern "C" int molkgnn_debug_phase_clocks_bwdp(unsigned long long* out48) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out48, mk::g_ph_bwdp, sizeof(unsigned long long) * 48) != cudaSuccess) return -1;
    unsigned long long z[48] = {0};
    return cudaMemcpyToSymbol(mk::g_ph_bwdp, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif
