// Molecule-tile forward of the molecular-kernel convolution (sm_100a, tcgen05 + TMEM).  Same contract as k_conv_fwd
// (conv_fwd.cu): replaces KernelConv.calculate_total_score (reference kernels.py:353-425) and the bucket gathers / output
// assembly of BaseKernelSetConv.forward (kernels.py:519-548, 674-747) for plans that carry molecule tiles.
//
// Formulation.  A tile is a run of <= 128 consecutive nodes holding whole molecules, so every neighbour of a tile node
// is itself a tile node.  Per (kernel block, tile) ONE dense GEMM on the tensor cores gives every cosine needed,
//        T[row, v] = khat[row,:] . xhat[v,:]              (M = 128 kernel rows, N = tile nodes, K = F)
// with no gather of neighbour feature rows: the normalised fp16 images of x are built once per layer in tile order
// (k_x_images) and fetched with one bulk async copy per tile visit, like the per-tile metadata (k_tile_meta, bucket.cu).
// fp32 accuracy: both operands are unscaled fp16 pairs v = hi + lo, three UMMAs per K step (hi*hi, lo*hi, hi*lo) into one
// fp32 accumulator.
//
// Block-major persistent kernel (tile.cuh): a CTA keeps one kernel block's images resident in shared memory, pulls tiles
// from a queue and per tile: (1) waits for the tile's MMAs, (2) starts the bulk copies of the next tile (the x image
// buffer is free as soon as the MMAs are done), (3) dumps the 128 x N accumulator TMEM -> shared memory ([column][row],
// conflict free both ways), (4) runs the epilogue with ONE THREAD PER (node, kernel) PAIR: the d x d similarity tile is
// 16 / 9 / 4 / 1 shared-memory reads, then exactly the reference arithmetic in registers -- sequential mean per
// permutation, first-max arg-max (kernels.py:373), bond cosine at the arg-max (kernels.py:382-390), chirality
// (kernels.py:279-350), softmax mix (kernels.py:402-425).  The next tile's MMAs are issued between the two halves of the
// epilogue and run on the second TMEM accumulator.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "tc.cuh"
#include "tile.cuh"

namespace mk {

bool tile_layer_ok(const molkgnn_layer_t* layer);
int tile_argmax_stride(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer);

// set by the stack driver while it runs the layer loop: it zeroes the tile queues of all layers with ONE memset up front
// (a memset between two kernels of the chain costs ~3 us of stream time)
bool g_fwd_counters_zeroed = false;

constexpr int TF_THREADS = 512;
constexpr int TF_STEAL_MIN = 8;                // tiles left in another block's queue that justify a block set-up
constexpr int TF_WARPS = TF_THREADS / 32;

// ---- normalised fp16 images of the activations in tile order --------------------------------------------------------
struct XImgArgs {
    const float* x; const float* xnorm; int ldx;
    // fused pad + norm (layer 0): x is the raw input with F valid columns and any row stride; the kernel also writes the
    // zero-padded copy out[N, Fp] and the row norms (what k_pad_norm would have produced) -- x is read once
    int F; float* out; float* norm_out;
    int Fp, Fk;
    const int* tile_start;
    unsigned char* ximg;
    int x_one;
};

// fused variant for narrow inputs (Fp <= 64): a quarter-warp (8 lanes x 8 columns) owns a node row, computes the norm with
// the summation order of k_pad_norm (lane-strided partial sums, butterfly) -- bitwise the same norms -- and emits the padded
// row, the norm and the (hi, lo) images
__device__ __forceinline__ void k_x_images_fused(const XImgArgs& a, int tile, int t0, int nn) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char* Xhi = a.ximg + (size_t)tile * 2 * a.x_one;
    unsigned char* Xlo = Xhi + a.x_one;
    const int rend = min(TNODES, (nn + 15) & ~15);
    for (int r = warp; r < rend; r += 16) {
        float v[2] = {0.f, 0.f};                  // columns lane, lane + 32
        const bool ok = r < nn;
        if (ok) {
            const float* xr = a.x + (size_t)(t0 + r) * a.ldx;
            if (lane < a.F) v[0] = xr[lane];
            if (lane + 32 < a.F) v[1] = xr[lane + 32];
        }
        float ss = v[0] * v[0];
        if (lane + 32 < a.Fp) ss += v[1] * v[1];
        ss = warp_sum(ss);
        const float nrm = sqrtf(ss);
        const float rinv = 1.0f / fmaxf(nrm, MOLKGNN_COS_EPS);
        if (ok) {
            if (lane < a.Fp) a.out[(size_t)(t0 + r) * a.Fp + lane] = v[0];
            if (lane + 32 < a.Fp) a.out[(size_t)(t0 + r) * a.Fp + lane + 32] = v[1];
            if (lane == 0) a.norm_out[t0 + r] = nrm;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int col = lane + 32 * h;
            if (col < a.Fk) {
                const float xv = v[h] * rinv;
                const __half hi = __float2half_rn(xv);
                const __half lo = __float2half_rn(xv - __half2float(hi));
                const uint32_t off = tc::il_off(r, col, a.Fk);
                *reinterpret_cast<__half*>(Xhi + off) = hi;
                *reinterpret_cast<__half*>(Xlo + off) = lo;
            }
        }
    }
}

__global__ void __launch_bounds__(512) k_x_images(const XImgArgs a) {
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int t0 = a.tile_start[tile], nn = a.tile_start[tile + 1] - t0;
    const int r = tid & (TNODES - 1), cg = tid >> 7;
    const bool ok = r < nn;
    if (a.out) { k_x_images_fused(a, tile, t0, nn); return; }
    float rinv = 0.f;
    const float* xr = a.x;
    if (ok) {
        rinv = 1.0f / fmaxf(a.xnorm[t0 + r], MOLKGNN_COS_EPS);
        xr = a.x + (size_t)(t0 + r) * a.ldx;
    }
    unsigned char* Xhi = a.ximg + (size_t)tile * 2 * a.x_one;
    unsigned char* Xlo = Xhi + a.x_one;
    const int nch = a.Fk >> 3;
    constexpr int UNR = 4;
    for (int c0 = cg; c0 < nch; c0 += 4 * UNR) {
        float4 v[UNR][2];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int c = c0 + 4 * u;
            v[u][0] = v[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok && c < nch) {
                if (8 * c + 4 <= a.Fp) v[u][0] = ld4(xr + 8 * c);
                if (8 * c + 8 <= a.Fp) v[u][1] = ld4(xr + 8 * c + 4);
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int c = c0 + 4 * u;
            if (c < nch) {
                __align__(16) __half2 hi[4];
                __align__(16) __half2 lo[4];
                tc::split_u2(v[u][0].x * rinv, v[u][0].y * rinv, hi[0], lo[0]);
                tc::split_u2(v[u][0].z * rinv, v[u][0].w * rinv, hi[1], lo[1]);
                tc::split_u2(v[u][1].x * rinv, v[u][1].y * rinv, hi[2], lo[2]);
                tc::split_u2(v[u][1].z * rinv, v[u][1].w * rinv, hi[3], lo[3]);
                const uint32_t off = tc::il_off(r, 8 * c, a.Fk);
                *reinterpret_cast<uint4*>(Xhi + off) = *reinterpret_cast<const uint4*>(hi);
                *reinterpret_cast<uint4*>(Xlo + off) = *reinterpret_cast<const uint4*>(lo);
            }
        }
    }
}

// ---- forward kernel ------------------------------------------------------------------------------------------------------
struct __align__(128) TileBuf {          // one in-flight tile: metadata record + its bond rows
    TileMetaG m;
    float ehat[TILE_ESLOTS][EP];
};

struct FwdTileArgs {
    const float* x; int ldx;
    int F, Fp, Fk;
    const TileMetaG* meta; const float* ehat_node;
    const unsigned char* ximg;
    int n_tiles;
    int L[4], koff[4];
    const float* packed[4];
    const unsigned char* img;
    TileBlocks tb;
    int img_one, x_one;
    int is_last;
    float* sc; int sc_mode; int ld_sc; long long scoff[4];
    uint8_t* argmax; uint8_t* argmax_free; const uint8_t* argmax_in;
    uint8_t* amT; int stride_am;      // nullable: the arg-max bytes once more in tile order (the backward's bulk-copy operand)
    int* counter;                     // [TILE_MAXB] tile queues
    int steal_min;                    // tiles left in another block's queue that justify a block set-up there
    int home_split[TILE_MAXB + 1];    // CTAs [home_split[b], home_split[b + 1]) have block b as their home
    int sm_img, sm_x, sm_dump, sm_buf, sm_es, sm_dup;   // byte offsets into dynamic shared memory
};

// segment constants kept in shared memory for the epilogue
struct SegConst {
    float ws, wc, we, W, rW;
    int d, k0, nk, rowbase, L;
    int es_off;                       // float4 offset of the segment's bond-support table [slot][half][kl]
    float rnk;                        // 1 / nk
    const int8_t* supsign;
};

__device__ __forceinline__ void tf_copy16(unsigned char* dst, const unsigned char* src, int64_t bytes) {
    for (int64_t i = (int64_t)threadIdx.x * 16; i < bytes; i += (int64_t)TF_THREADS * 16)
        *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(src + i);
}

// thread 0: metadata record, bond rows and node images of `tile` -> buffer, all completing on `bar`
__device__ __forceinline__ void tf_issue_copy(const FwdTileArgs& a, unsigned char* smem, TileBuf* buf, int xb, int tile,
                                              int e0, int ne, uint64_t* bar) {
    const TileMetaG* g = a.meta + tile;
    const uint32_t eb = (uint32_t)ne * EP * 4u;
    mbar_expect_tx(bar, (uint32_t)sizeof(TileMetaG) + eb + 2u * (uint32_t)a.x_one);
    bulk_g2s(&buf->m, g, (uint32_t)sizeof(TileMetaG), bar);
    if (eb) bulk_g2s(&buf->ehat[0][0], a.ehat_node + (size_t)e0 * EP, eb, bar);
    bulk_g2s(smem + a.sm_x + (size_t)xb * 2 * a.x_one, a.ximg + (size_t)tile * 2 * a.x_one, 2u * (uint32_t)a.x_one, bar);
}

// thread 0: (Fk/16) K steps x 3 UMMAs into accumulator `set`
// The kernel block's images (the A operand) live in TENSOR MEMORY (TF_A_HI / TF_A_LO, written once per block by the
// consumer warps): that frees 57 KB of shared memory for a second node-image buffer, so the copy of tile t+1 overlaps the
// MMAs and the epilogue of tile t instead of waiting for the single buffer.
constexpr uint32_t TF_A_HI = 256u, TF_A_LO = 320u;      // TMEM columns: accumulators 0..255, A hi 256.., A lo 320.. (Fk/2 <= 64 each)

__device__ __forceinline__ void tf_issue_mma(const FwdTileArgs& a, unsigned char* smem, int nn, uint32_t tmem, int set,
                                             uint64_t* bar) {
    const uint32_t sbo = (uint32_t)(a.Fk >> 3) * 128u;
    const uint32_t xhi = tc::smem_u32(smem + a.sm_x + (size_t)set * 2 * a.x_one), xlo = xhi + (uint32_t)a.x_one;
    const int N = max(16, (nn + 15) & ~15);
    const uint32_t idesc = tc::idesc_f16(128, N, 0, 0);
    const uint32_t d = tmem + (uint32_t)(set * TNODES);
    const int nks = a.Fk >> 4;
    for (int ks = 0; ks < nks; ++ks) {
        const uint32_t o = (uint32_t)ks * 256u;
        const uint32_t aH = tmem + TF_A_HI + (uint32_t)ks * 8u, aL = tmem + TF_A_LO + (uint32_t)ks * 8u;
        const uint64_t dBh = tc::smem_desc(xhi + o, 128u, sbo), dBl = tc::smem_desc(xlo + o, 128u, sbo);
        tc::umma_f16_ts(d, aH, dBh, idesc, ks > 0 ? 1u : 0u);
        tc::umma_f16_ts(d, aL, dBh, idesc, 1u);
        tc::umma_f16_ts(d, aH, dBl, idesc, 1u);
    }
    tc::umma_commit(bar);
}

// accumulator set -> dump[column][row]: warp (quadrant q, column block cb) moves 32 rows x 32 columns
__device__ __forceinline__ void tf_dump(float* dump, uint32_t tmem, int set, int nn) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = warp & 3, cb = warp >> 2;
    if (cb * 32 >= nn) return;
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * TNODES + cb * 32);
    tc::tmem_ld16(taddr, v);
    tc::tmem_ld16(taddr + 16, v + 16);
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(v[i]));
    float* dst = dump + (size_t)(cb * 32) * 128 + q * 32 + lane;
#pragma unroll
    for (int i = 0; i < 32; ++i) dst[i * 128] = __uint_as_float(v[i]);
}

template <int D> __device__ __forceinline__ uint32_t perm_code_rt(int p) {
    uint32_t c = 0;
#pragma unroll
    for (int q = 0; q < Perm<D>::P; ++q) if (q == p) c = perm_code<D>(q);
    return c;
}

// one (node, kernel) pair: the reference arithmetic on its d x d similarity tile.  Results are returned, the caller
// stores them: two pairs per thread are evaluated back to back so that their dependency chains interleave.
struct PairOut { float sc; size_t cidx, oidx; int tidx; uint8_t am, free; };

template <int D, bool FORCED>
__device__ __forceinline__ PairOut tf_pair(const FwdTileArgs& a, const TileBuf* tb, const float* dump, const float4* estab,
                                           const unsigned char* dupf, const SegConst& sg, int nl_, int kl) {
    constexpr int P = Perm<D>::P;
    const TileMetaG& m = tb->m;
    const uint32_t nw = m.nl[nl_];
    const int R = m.posl[nl_];
    const int e0 = m.eslot[nl_];
    const int k = sg.k0 + kl;
    const float* col0 = dump + sg.rowbase + kl;
    float T[D][D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const float* c = col0 + ((nw >> (8 * j)) & 0xffu) * 128;
#pragma unroll
        for (int s = 0; s < D; ++s) T[j][s] = c[s * sg.nk];
    }
    const float cdot = col0[nl_ * 128 + D * sg.nk];
    PairOut o;
    o.cidx = (size_t)a.scoff[D - 1] + (size_t)R * sg.L + k;
    {
        int t = m.lidx[nl_] * sg.L + k;              // tile order: degree blocks, node-of-degree major, kernel minor
#pragma unroll
        for (int dd = 1; dd < D; ++dd) t += m.cnt[dd - 1] * a.L[dd - 1];
        o.tidx = t;
    }
    const int forced = FORCED && a.argmax_in ? (a.argmax_in[o.cidx] & 0x7f) : -1;
    // mean over j for every permutation: sequential sum, then true division (kernels.py:194)
    float best = 0.f, used = 0.f;
    int bi = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        float s = T[0][Perm<D>::at(p, 0)];
#pragma unroll
        for (int j = 1; j < D; ++j) s += T[j][Perm<D>::at(p, j)];
        s = div_deg<D>(s);
        if (p == 0 || s > best) { best = s; bi = p; }   // first maximum wins (torch.max, kernels.py:373)
        if (FORCED && p == forced) used = s;
    }
    o.free = (uint8_t)bi;
    if (FORCED && forced >= 0 && forced < P) { bi = forced; best = used; }
    // bond-attribute cosine at the chosen permutation (kernels.py:382-390)
    const uint32_t code = perm_code_rt<D>(bi);
    float esum = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const int s = (code >> (2 * j)) & 3;
        const float4 e0v = *reinterpret_cast<const float4*>(&tb->ehat[e0 + j][0]);
        const float4 e1v = *reinterpret_cast<const float4*>(&tb->ehat[e0 + j][4]);
        const float4 s0 = estab[sg.es_off + (s * 2 + 0) * sg.nk + kl];
        const float4 s1 = estab[sg.es_off + (s * 2 + 1) * sg.nk + kl];
        float dd = 0.f;
        dd = fmaf(e0v.x, s0.x, dd); dd = fmaf(e0v.y, s0.y, dd); dd = fmaf(e0v.z, s0.z, dd); dd = fmaf(e0v.w, s0.w, dd);
        dd = fmaf(e1v.x, s1.x, dd); dd = fmaf(e1v.y, s1.y, dd); dd = fmaf(e1v.z, s1.z, dd); dd = fmaf(e1v.w, s1.w, dd);
        esum = j == 0 ? dd : esum + dd;
    }
    const float E = div_deg<D>(esum);
    float sc = div_by((best * sg.ws + cdot * sg.wc) + E * sg.we, sg.W, sg.rW);
    uint8_t am = (uint8_t)bi;
    if (D == 4 && a.is_last) {
        // chirality (kernels.py:279-350): +1 if any two neighbours are identical, else sign agreement
        int chi = 1;
        if (!dupf[nl_]) chi = (m.tsg[nl_] == sg.supsign[k * 12 + bi]) ? 1 : -1;
        if (chi < 0) { sc = -sc; am |= 0x80; }
    }
    o.sc = sc; o.am = am;
    o.oidx = a.sc_mode == 0 ? o.cidx : (size_t)(m.t0 + nl_) * a.ld_sc + a.koff[D - 1] + k;
    return o;
}

template <int D, bool FORCED>
__device__ __forceinline__ void tf_pairs_of_thread(const FwdTileArgs& a, const TileBuf* tb, const float* dump,
                                                   const float4* estab, const unsigned char* dupf, const SegConst& sg,
                                                   int np, int tile) {
    // the segment's pairs (node-major, then kernel), two per thread and iteration
    for (int p = (int)threadIdx.x; p < np; p += 2 * TF_THREADS) {
        const int p2 = p + TF_THREADS;
        const int ni = (int)(((float)p + 0.5f) * sg.rnk);
        const PairOut o1 = tf_pair<D, FORCED>(a, tb, dump, estab, dupf, sg, tb->m.list[D - 1][ni], p - ni * sg.nk);
        PairOut o2;
        const bool has2 = p2 < np;
        if (has2) {
            const int ni2 = (int)(((float)p2 + 0.5f) * sg.rnk);
            o2 = tf_pair<D, FORCED>(a, tb, dump, estab, dupf, sg, tb->m.list[D - 1][ni2], p2 - ni2 * sg.nk);
        }
        if (FORCED && a.argmax_free) a.argmax_free[o1.cidx] = o1.free;
        a.argmax[o1.cidx] = o1.am;
        if (a.amT) a.amT[(size_t)tile * a.stride_am + o1.tidx] = o1.am;
        a.sc[o1.oidx] = o1.sc;
        if (has2) {
            if (FORCED && a.argmax_free) a.argmax_free[o2.cidx] = o2.free;
            a.argmax[o2.cidx] = o2.am;
            if (a.amT) a.amT[(size_t)tile * a.stride_am + o2.tidx] = o2.am;
            a.sc[o2.oidx] = o2.sc;
        }
    }
}

// the block's pairs, segment by segment
template <bool FORCED>
__device__ __forceinline__ void tf_epilogue(const FwdTileArgs& a, const TileBuf* tb, const float* dump, const float4* estab,
                                            const unsigned char* dupf, const SegConst* segs, int nseg, int tile) {
    for (int si = 0; si < nseg; ++si) {
        const SegConst sg = segs[si];
        const int np = tb->m.cnt[sg.d - 1] * sg.nk;
        switch (sg.d) {
            case 1: tf_pairs_of_thread<1, FORCED>(a, tb, dump, estab, dupf, sg, np, tile); break;
            case 2: tf_pairs_of_thread<2, FORCED>(a, tb, dump, estab, dupf, sg, np, tile); break;
            case 3: tf_pairs_of_thread<3, FORCED>(a, tb, dump, estab, dupf, sg, np, tile); break;
            default: tf_pairs_of_thread<4, FORCED>(a, tb, dump, estab, dupf, sg, np, tile); break;
        }
    }
}

// chirality gate of the degree-4 nodes of a tile: any two of the four neighbour feature rows bit-equal (torch.equal,
// kernels.py:310-317); one warp per node, raw rows from global memory
__device__ __forceinline__ void tf_dup_flags(const FwdTileArgs& a, const TileBuf* tb, unsigned char* dupf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n4 = tb->m.cnt[3], t0 = tb->m.t0;
    for (int i = warp; i < n4; i += TF_WARPS) {
        const int nl_ = tb->m.list[3][i];
        const uint32_t w = tb->m.nl[nl_];
        unsigned neq = 0;
        for (int f = lane; f < a.F; f += 32) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = a.x[(size_t)(t0 + ((w >> (8 * j)) & 0xff)) * a.ldx + f];
            int b = 0;
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = p + 1; q < 4; ++q, ++b) if (!(v[p] == v[q])) neq |= 1u << b;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) neq |= __shfl_xor_sync(0xffffffffu, neq, o);
        if (lane == 0) dupf[nl_] = (neq != 0x3fu) ? 1 : 0;
    }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TF_THREADS) : "memory"); }

// 16 consumer warps (dump + epilogue) and one producer warp (tile queue, bulk copies, UMMA issue): the producer runs one
// tile ahead, bounded only by the single node-image buffer (free when the previous tile's MMAs are done), the two TMEM
// accumulators and the two metadata buffers.
#ifdef MK_PHASE_CLOCKS
__device__ unsigned long long g_ph_fwd[32];      // [0..15] consumer thread 0, [16..31] producer lane
#endif

template <bool FORCED>
__global__ void __launch_bounds__(TF_THREADS + 32, 1) k_conv_fwd_tile(const __grid_constant__ FwdTileArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar_mma[2], bar_cp[2], bar_tfree[2], bar_bfree[2];
    __shared__ uint32_t tslot;
    __shared__ int s_tile[2];
    __shared__ SegConst s_seg[TILE_MAXSEG];
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool producer = warp == TF_WARPS;
    MK_PH_DECL(tid == 0 || tid == TF_THREADS)
    if (warp == 0) tc::tmem_alloc(&tslot, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tslot;
    TileBuf* bufs = reinterpret_cast<TileBuf*>(smem + a.sm_buf);
    float* dump = reinterpret_cast<float*>(smem + a.sm_dump);
    float4* estab = reinterpret_cast<float4*>(smem + a.sm_es);
    unsigned char* dupf = smem + a.sm_dup;

    // Every CTA has a HOME block (blockIdx % nb) whose queue it drains first: with the CTAs spread over the blocks a CTA pays
    // one block set-up + pipeline fill per launch instead of one per block.  Afterwards it helps with the other blocks'
    // queues, but only where enough tiles are left to be worth a set-up; whatever it leaves is drained by that block's home
    // CTAs, which never leave their queue before it is empty.
    __shared__ int s_left;
    const int nb = a.tb.nb;
    int home = 0;
    while (home + 1 < nb && (int)blockIdx.x >= a.home_split[home + 1]) ++home;
    for (int bi = 0; bi < nb; ++bi) {
        const int blk = (home + bi) % nb;
        __syncthreads();                       // previous block completely finished
        if (tid == 0) s_left = a.n_tiles - min(a.n_tiles, *reinterpret_cast<volatile int*>(a.counter + blk));
        __syncthreads();
        {
            const bool has_home = a.home_split[blk + 1] > a.home_split[blk];
            const int need = (bi == 0 || !has_home) ? 1 : a.steal_min;
            if (s_left < need) continue;
        }
        // ---- block set-up: barriers, images, segment constants, bond-support table [slot][half][kernel] ----
        if (tid == 0) {
            for (int i = 0; i < 2; ++i) {
                tc::mbar_init(&bar_mma[i], 1); tc::mbar_init(&bar_cp[i], 1);
                tc::mbar_init(&bar_tfree[i], 1); tc::mbar_init(&bar_bfree[i], 1);
            }
            tc::fence_mbar_init();
        }
        const int nseg = a.tb.nseg[blk];
        bool has4 = false;
        if (!producer) {
            // kernel-block images -> tensor memory: thread = kernel row (TMEM lane), 4 warps per lane quadrant share the
            // Fk / 2 packed columns; the global image row is 16 B per 8-column chunk in the interleaved layout
            {
                const int row = (warp & 3) * 32 + (tid & 31);
                const unsigned char* gh = a.img + (size_t)blk * 2 * a.img_one;
                const unsigned char* gl = gh + a.img_one;
                const int c_end = min(a.Fk >> 1, ((warp >> 2) + 1) * 16);
                for (int c0 = (warp >> 2) * 16; c0 < c_end; c0 += 8) {       // 8 TMEM columns = 16 fp16 = two 16-byte chunks
                    const uint32_t o0 = tc::il_off(row, 2 * c0, a.Fk), o1 = tc::il_off(row, 2 * c0 + 8, a.Fk);
                    uint4 h0 = __ldg(reinterpret_cast<const uint4*>(gh + o0)), h1 = __ldg(reinterpret_cast<const uint4*>(gh + o1));
                    uint4 l0 = __ldg(reinterpret_cast<const uint4*>(gl + o0)), l1 = __ldg(reinterpret_cast<const uint4*>(gl + o1));
                    const uint32_t vh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                    const uint32_t vl[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
                    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
                    tc::tmem_st8(lane_addr + TF_A_HI + (uint32_t)c0, vh);
                    tc::tmem_st8(lane_addr + TF_A_LO + (uint32_t)c0, vl);
                }
                tc::tmem_st_wait();
            }
            int es_off = 0;
            for (int si = 0; si < nseg; ++si) {
                const TileSeg sg = a.tb.seg[blk][si];
                const int L = a.L[sg.d - 1];
                const PackedLayout pl(sg.d, L, a.Fp);
                const float* pk = a.packed[sg.d - 1];
                if (sg.d == 4) has4 = true;
                if (tid == 0) {
                    SegConst c;
                    c.ws = pk[pl.w + 0]; c.wc = pk[pl.w + 1]; c.we = pk[pl.w + 2]; c.W = pk[pl.w + 3];
                    c.rW = 1.0f / c.W;
                    c.d = sg.d; c.k0 = sg.k0; c.nk = sg.nk; c.rowbase = sg.rowbase; c.L = L;
                    c.es_off = es_off; c.rnk = 1.0f / (float)sg.nk;
                    c.supsign = reinterpret_cast<const int8_t*>(pk + pl.sign);
                    s_seg[si] = c;
                }
                for (int i = tid; i < sg.d * 2 * sg.nk; i += TF_THREADS) {
                    const int kl = i % sg.nk, sh = i / sg.nk;          // sh = slot * 2 + half
                    estab[es_off + i] =
                        *reinterpret_cast<const float4*>(pk + pl.es + (size_t)((sh >> 1) * L + sg.k0 + kl) * EP + (sh & 1) * 4);
                }
                es_off += sg.d * 2 * sg.nk;
            }
        }
        tc::fence_before_sync();               // the images were written with tcgen05.st, the producer's MMAs read them
        __syncthreads();
        tc::fence_after_sync();
        MK_PH(0);                                                          // block set-up (images, tables)

        if (producer) {
            if ((tid & 31) == 0) {
                // Event loop over two independent duties, each for the oldest tile that still needs it:
                //   copy  (tile seq_c): needs buffer set seq_c & 1 released by the consumers (tile seq_c - 2 finished)
                //   MMA   (tile seq_m): needs its copies landed and accumulator seq_m & 1 drained (dump of seq_m - 2 done)
                // so the copy of tile t+1 is in flight while the MMAs of tile t are issued and run.
                int seq_c = 0, seq_m = 0;
                bool last_copied = false;            // the end marker went out: no more tiles in this block's queue
                // the next tile of the queue and its bond-slot range are fetched one tile ahead: the atomic and the header
                // load (two L2 round trips) stay off the path between "buffer released" and "copy issued"
                int nxt = atomicAdd(a.counter + blk, 1);
                int2 nxt_e = nxt < a.n_tiles ? *reinterpret_cast<const int2*>(&a.meta[nxt].e0) : make_int2(0, 0);
                uint32_t idle = 0;                   // bounded: a lost completion traps instead of hanging the device
                while (true) {
                    if (++idle > (1u << 26)) __trap();
                    if (!last_copied && seq_c - seq_m < 2) {
                        const int b = seq_c & 1;
                        if (tc::mbar_try_wait(&bar_bfree[b], ((uint32_t)(seq_c >> 1) & 1u) ^ 1u)) {
                            const int tile = nxt;
                            s_tile[b] = tile < a.n_tiles ? tile : -1;
                            if (tile >= a.n_tiles) { mbar_arrive(&bar_cp[b]); last_copied = true; }
                            else {
                                tf_issue_copy(a, smem, &bufs[b], b, tile, nxt_e.x, nxt_e.y, &bar_cp[b]);
                                nxt = atomicAdd(a.counter + blk, 1);
                                nxt_e = nxt < a.n_tiles ? *reinterpret_cast<const int2*>(&a.meta[nxt].e0) : make_int2(0, 0);
                            }
                            ++seq_c;
                            idle = 0;
                            MK_PH(2);
                        }
                    }
                    const int pending = seq_c - seq_m - (last_copied ? 1 : 0);     // copied tiles whose MMAs are not issued yet
                    if (pending > 0) {
                        const int b = seq_m & 1;
                        const uint32_t par = (uint32_t)(seq_m >> 1) & 1u;
                        if (tc::mbar_try_wait(&bar_cp[b], par) && tc::mbar_try_wait(&bar_tfree[b], par ^ 1u)) {
                            tc::fence_after_sync();
                            MK_PH(4);
                            tf_issue_mma(a, smem, bufs[b].m.nn, tmem, b, &bar_mma[b]);
                            ++seq_m;
                            idle = 0;
                            MK_PH(5);
                        }
                    } else if (last_copied) {
                        break;
                    }
                }
            }
        } else {
            for (int seq = 0;; ++seq) {
                const int b = seq & 1;
                const uint32_t par = (uint32_t)(seq >> 1) & 1u;
                const TileBuf* tb = &bufs[b];
                tc::mbar_wait(&bar_cp[b], par);                        // metadata (or the end marker) visible
                MK_PH(1);
                if (s_tile[b] < 0) break;
                tc::mbar_wait(&bar_mma[b], par);                       // accumulator ready
                tc::fence_after_sync();
                MK_PH(2);
                tf_dump(dump, tmem, b, tb->m.nn);
                MK_PH(3);
                if (has4 && a.is_last) tf_dup_flags(a, tb, dupf);
                tc::fence_before_sync();
                consumer_sync();                                       // dump complete, accumulator drained
                MK_PH(4);
                if (tid == 0) mbar_arrive(&bar_tfree[b]);
                tf_epilogue<FORCED>(a, tb, dump, estab, dupf, s_seg, nseg, s_tile[b]);
                MK_PH(5);
                consumer_sync();                                       // dump and this tile's buffer are free again
                MK_PH(6);
                if (tid == 0) mbar_arrive(&bar_bfree[b]);
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    MK_PH(7);
#ifdef MK_PHASE_CLOCKS
    MK_PH_FLUSH(g_ph_fwd + (tid == 0 ? 0 : 16));
#endif
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ---- host side -----------------------------------------------------------------------------------------------------
static bool tile_plan_ok(const molkgnn_plan_t* plan) {
    return plan->n_tiles > 0 && plan->tile_start && plan->tile_meta && plan->ehat_node && plan->tile_max_nodes <= TNODES;
}

int launch_x_images(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                    const float* xnorm, void* ximg, float* pad_out, float* norm_out, cudaStream_t st) {
    XImgArgs a;
    a.x = x; a.xnorm = xnorm; a.ldx = ldx;
    a.F = layer->F; a.out = pad_out; a.norm_out = norm_out;
    a.Fp = layer->Fp; a.Fk = tile_fk(layer->Fp);
    a.tile_start = plan->tile_start;
    a.ximg = reinterpret_cast<unsigned char*>(ximg);
    a.x_one = tile_img_one(a.Fk);
    count_launches(1);
    ProfScope prof("x_images", st);
    k_x_images<<<plan->n_tiles, 512, 0, st>>>(a);
    MK_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// returns 1 if launched, 0 if the plan / layer is not eligible (caller falls back to the bucket-order kernels), <0 on error
int launch_conv_fwd_tile(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                         const void* ximg, int32_t is_last_layer, float* sc, int32_t sc_mode, int32_t ld_sc,
                         const int64_t scoff[4], uint8_t* argmax, uint8_t* argmax_free, const uint8_t* argmax_in,
                         uint8_t* argmax_tile, int32_t* counter, cudaStream_t st) {
    if (!ximg || !tile_plan_ok(plan) || !layer->tile_img || !tile_layer_ok(layer)) return 0;
    static int s_budget = 0, s_sms = 0;
    if (!s_budget) {
        s_budget = device_max_smem_optin();
        s_sms = device_num_sms();
        MK_REQUIRE(s_budget > 0 && s_sms > 0, "conv_fwd_tile: no CUDA device");
    }
    FwdTileArgs a;
    a.x = x; a.ldx = ldx;
    a.F = layer->F; a.Fp = layer->Fp; a.Fk = tile_fk(layer->Fp);
    a.meta = reinterpret_cast<const TileMetaG*>(plan->tile_meta);
    a.ehat_node = plan->ehat_node;
    a.ximg = reinterpret_cast<const unsigned char*>(ximg);
    a.n_tiles = plan->n_tiles;
    for (int d = 0; d < 4; ++d) {
        a.L[d] = layer->L[d]; a.koff[d] = layer->koff[d];
        a.packed[d] = layer->packed[d];
        a.scoff[d] = scoff[d];
    }
    if (!a.tb.build(layer->L)) return 0;
    a.img = reinterpret_cast<const unsigned char*>(layer->tile_img);
    a.img_one = tile_img_one(a.Fk);
    a.x_one = tile_img_one(a.Fk);
    a.is_last = is_last_layer;
    a.sc = sc; a.sc_mode = sc_mode; a.ld_sc = ld_sc;
    a.argmax = argmax; a.argmax_free = argmax_free; a.argmax_in = argmax_in;
    a.amT = argmax_tile; a.stride_am = tile_argmax_stride(plan, layer);
    a.counter = counter;
    static int s_steal = -1;
    if (s_steal < 0) { const char* e = getenv("MOLKGNN_FWD_STEAL_MIN"); s_steal = e ? std::max(1, atoi(e)) : TF_STEAL_MIN; }
    a.steal_min = s_steal;
    {
        // home CTAs per block in proportion to the block's estimated cost per tile visit (fixed part + its share of the
        // (node, kernel) pairs, weighted by the degree's permutation work); MOLKGNN_FWD_HOME="w0,w1,.." overrides
        const int grid_ = std::min(plan->n_tiles, s_sms);
        double w[TILE_MAXB], wsum = 0.0;
        static const double pair_cost[4] = {0.35, 0.6, 1.0, 1.6};      // relative epilogue cost of a degree-d pair
        for (int b = 0; b < a.tb.nb; ++b) {
            double pairs = 0.0;
            for (int si = 0; si < a.tb.nseg[b]; ++si) {
                const TileSeg sg = a.tb.seg[b][si];
                pairs += pair_cost[sg.d - 1] * sg.nk * (double)plan->n[sg.d - 1] / std::max(1, plan->n_tiles);
            }
            w[b] = 1400.0 + pairs;                                      // ~ cycles / 2: fixed visit cost + pair work
            wsum += w[b];
        }
        if (const char* e = getenv("MOLKGNN_FWD_HOME")) {
            wsum = 0.0;
            for (int b = 0; b < a.tb.nb; ++b) { w[b] = std::max(0.0, atof(e)); wsum += w[b]; const char* c = strchr(e, ','); e = c ? c + 1 : e; }
            if (wsum <= 0.0) { for (int b = 0; b < a.tb.nb; ++b) w[b] = 1.0; wsum = a.tb.nb; }
        }
        double acc = 0.0;
        a.home_split[0] = 0;
        for (int b = 0; b < a.tb.nb; ++b) {
            acc += w[b];
            a.home_split[b + 1] = b + 1 == a.tb.nb ? grid_ : (int)(grid_ * acc / wsum + 0.5);
        }
        for (int b = a.tb.nb + 1; b <= TILE_MAXB; ++b) a.home_split[b] = grid_;
    }
    int64_t off = 0;
    a.sm_img = 0;                                  // the kernel-block images live in tensor memory
    a.sm_x = (int)off; off += 2 * 2 * (int64_t)a.x_one;   // two node-image buffers
    a.sm_dump = (int)off; off += (int64_t)TNODES * 128 * 4;
    a.sm_buf = (int)off; off += 2 * (int64_t)sizeof(TileBuf);
    a.sm_es = (int)off; off += 128 * 32;        // <= 128 support rows x 8 floats per block
    a.sm_dup = (int)off; off += 128;
    if (off > s_budget - 1024) return 0;
    if (!g_fwd_counters_zeroed) MK_CHECK_CUDA(cudaMemsetAsync(counter, 0, TILE_MAXB * sizeof(int), st));
    static int64_t s_attr_dev[16] = {0};
    int64_t& s_attr = s_attr_dev[device_index()];      // function attributes are per device
    if (off > s_attr) {
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_conv_fwd_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off));
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_conv_fwd_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off));
        s_attr = off;
    }
    const int grid = std::min(plan->n_tiles, s_sms);
    count_launches(1);
    if (argmax_in || argmax_free) k_conv_fwd_tile<true><<<grid, TF_THREADS + 32, off, st>>>(a);     // parity harness / replay
    else k_conv_fwd_tile<false><<<grid, TF_THREADS + 32, off, st>>>(a);
    MK_CHECK_CUDA(cudaGetLastError());
    return 1;
}

}  // namespace mk

using namespace mk;

namespace mk {
bool wide_layer_ok(const molkgnn_layer_t* layer);
int launch_x_images_wide(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                         const float* xnorm, void* ximg, cudaStream_t st);
}

extern "C" int64_t molkgnn_tile_ximg_bytes(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer) {
    if (tile_plan_ok(plan) && wide_layer_ok(layer)) return 2 * (int64_t)(wide_fk(layer->Fp) / 32) * WIDE_STAGE;   // forward + backward layouts
    if (!tile_plan_ok(plan) || !tile_layer_ok(layer)) return 0;
    return 2 * (int64_t)tile_img_one(tile_fk(layer->Fp));
}

extern "C" int molkgnn_tile_ximg_build(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                                       const float* xnorm, void* ximg, void* stream_) {
    if (tile_plan_ok(plan) && wide_layer_ok(layer)) {
        MK_REQUIRE(ldx % 4 == 0 && ldx >= layer->Fp && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(ximg) & 127) == 0, "tile_ximg_build: bad alignment / ldx=%d", ldx);
        return launch_x_images_wide(plan, layer, x, ldx, xnorm, ximg, (cudaStream_t)stream_);
    }
    MK_REQUIRE(tile_plan_ok(plan) && tile_layer_ok(layer), "tile_ximg_build: plan or layer is not eligible for the tile kernels");
    MK_REQUIRE(ldx % 4 == 0 && ldx >= layer->Fp, "tile_ximg_build: ldx=%d must be a multiple of 4 and >= Fp=%d", ldx, layer->Fp);
    MK_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(ximg) & 127) == 0,
               "tile_ximg_build: x must be 16-byte and ximg 128-byte aligned");
    return launch_x_images(plan, layer, x, ldx, xnorm, ximg, nullptr, nullptr, (cudaStream_t)stream_);
}

// pad + norm + images of the raw layer-0 input in one pass (Fp <= 64): out [N, Fp] zero padded, norm [N], ximg as above
extern "C" int molkgnn_tile_ximg_build_raw(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x,
                                           int32_t ldx, float* out, float* norm, void* ximg, void* stream_) {
    MK_REQUIRE(tile_plan_ok(plan) && tile_layer_ok(layer) && layer->Fp <= 64,
               "tile_ximg_build_raw: needs a tiled plan, an eligible layer and Fp <= 64");
    MK_REQUIRE(out && norm && (reinterpret_cast<uintptr_t>(ximg) & 127) == 0, "tile_ximg_build_raw: bad arguments");
    return launch_x_images(plan, layer, x, ldx, nullptr, ximg, out, norm, (cudaStream_t)stream_);
}

#ifdef MK_PHASE_CLOCKS
// profiling build only: read (and clear) the accumulated phase clocks of k_conv_fwd_tile
extern "C" int molkgnn_debug_phase_clocks_fwd(unsigned long long* out32) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out32, mk::g_ph_fwd, sizeof(unsigned long long) * 32) != cudaSuccess) return -1;
    unsigned long long z[32] = {0};
    return cudaMemcpyToSymbol(mk::g_ph_fwd, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif
