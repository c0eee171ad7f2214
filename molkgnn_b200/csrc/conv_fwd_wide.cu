// Tile-major tensor-core forward for WIDE layers (sm_100a, tcgen05 + TMEM): kernel sets whose rows or whose feature width do not
// fit the resident-operand design of conv_fwd_tile.cu / stack_fwd_fused.cu (more than 8 kernel blocks of 128 rows, or more than
// 112 features: BASELINE configs[2], 40/80/120/200 kernels x 440 features = 15 blocks x 28 K steps).  Same contract as
// k_conv_fwd (conv_fwd.cu); replaces KernelConv.calculate_total_score (reference kernels.py:353-425) for such layers.
//
// Formulation as in the other tile kernels -- per molecule tile (<= 128 whole-molecule nodes) and kernel block (<= 128 rows of
// ONE degree) a dense GEMM on the tensor cores,  T[row, v] = khat[row, :] . xhat[v, :]  (M = 128 kernel rows, N = tile nodes,
// K = F), then the reference arithmetic per (node, kernel) pair on its d x d entries of T -- but here NEITHER operand is
// resident: a 128 x 448 (hi, lo) image is 229 KB.  Both operands are streamed from L2 through rings of 16 KB stages (two K
// steps = 32 features of a 128-row operand, [K step][hi | lo], written in that order by k_param_pack_wide / k_x_images_wide):
//   ring warp   per (tile, block pair, stage): the node-image stage (B ring) and the stage of each block of the pair (one A ring
//               per MMA warp) -- a pure function of the sequence number, so it runs ahead across pairs and tiles;
//   MMA warps   warp W issues block 2 g + W of pair g: per stage 2 K steps x 3 tcgen05.mma (hi*hi, lo*hi, hi*lo) into the block's
//               TMEM accumulator (4 x 128 columns: the pairs alternate between two sets), frees its A stage and -- both warps --
//               the shared B stage by tcgen05.commit;
//   consumers   16 warps: accumulator -> shared memory ([column][row]), one thread per (node, kernel) pair, scores / arg-max bytes
//               straight to the bucket-order arrays the backward and the neighbour sum read.
// fp32 accuracy: both operands are unscaled fp16 pairs v = hi + lo, three MMAs per K step into one fp32 accumulator.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "tc.cuh"
#include "tile.cuh"

namespace mk {

bool tile_layer_ok(const molkgnn_layer_t* layer);

constexpr int WF_CONS = 512;               // consumer threads
constexpr int WF_CWARPS = WF_CONS / 32;
constexpr int WF_RINGS = 3;                // ring warps: node-image stages, stages of the pair's first / second block
constexpr int WF_THREADS = WF_CONS + 32 * (WF_RINGS + 2);   // + ring warps + two MMA warps
constexpr int WF_MAXSLOT = 4;

bool wide_layer_ok(const molkgnn_layer_t* layer) {
    WideBlocks wb;
    return layer->K > 0 && !tile_layer_ok(layer) && wide_fk(layer->Fp) <= 512 && wb.build(layer->L);
}

int64_t wide_img_bytes(const molkgnn_layer_t* layer) {
    WideBlocks wb;
    if (!wide_layer_ok(layer) || !wb.build(layer->L)) return 0;
    const int Fk = wide_fk(layer->Fp);
    return wide_img_bwd_off(wb.nb, Fk, layer->L) + (int64_t)wb.nb * 8 * Fk * 64;      // forward stages, bond table, backward stages
}

// ---- operand images ----------------------------------------------------------------------------------------------------
// node images: [tile][stage kk][K step][hi | lo][16 row groups][2 chunks][8 rows][8 elements]; rows = tile-local node, rows
// beyond the tile's nodes are zero.  One CTA per (tile, stage): thread (v, c) converts 8 features of node v.
struct WideXArgs {
    const float* x; const float* xnorm; int ldx, Fp;
    const int* tile_start;
    unsigned char* ximg; int nk2;
    unsigned char* ximg_bwd; int Fk;      // K-step-major MN-major copy for the backward (tile.cuh), [tile][8 stages]
};

__global__ void __launch_bounds__(512) k_x_images_wide(const WideXArgs a) {
    const int tile = blockIdx.x, kk = blockIdx.y, tid = threadIdx.x;
    const int t0 = a.tile_start[tile], nn = a.tile_start[tile + 1] - t0;
    const int v = tid >> 2, c = tid & 3;
    const int col0 = kk * 32 + c * 8;
    float xv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) xv[u] = 0.f;
    if (v < nn && col0 < a.Fp) {
        const float rinv = 1.0f / fmaxf(a.xnorm[t0 + v], MOLKGNN_COS_EPS);
        const float* xr = a.x + (size_t)(t0 + v) * a.ldx + col0;
        const float4 q0 = ld4(xr);
        xv[0] = q0.x * rinv; xv[1] = q0.y * rinv; xv[2] = q0.z * rinv; xv[3] = q0.w * rinv;
        if (col0 + 4 < a.Fp) {
            const float4 q1 = ld4(xr + 4);
            xv[4] = q1.x * rinv; xv[5] = q1.y * rinv; xv[6] = q1.z * rinv; xv[7] = q1.w * rinv;
        }
    }
    __align__(16) __half2 hi[4];
    __align__(16) __half2 lo[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) tc::split_u2(xv[2 * u] * WIDE_OPSCALE, xv[2 * u + 1] * WIDE_OPSCALE, hi[u], lo[u]);
    unsigned char* dst = a.ximg + ((size_t)tile * a.nk2 + kk) * WIDE_STAGE + wide_stage_off(v, c);
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(dst + WIDE_STAGE / 4) = *reinterpret_cast<const uint4*>(lo);
    unsigned char* db = a.ximg_bwd + (size_t)tile * 8 * a.Fk * 64 + wide_bstage_off(v, kk * 4 + c, a.Fk);
    *reinterpret_cast<uint4*>(db) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(db + a.Fk * 32) = *reinterpret_cast<const uint4*>(lo);
}

int launch_x_images_wide(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                         const float* xnorm, void* ximg, cudaStream_t st) {
    MK_REQUIRE(xnorm, "x_images_wide: row norms are required");
    WideXArgs a;
    a.x = x; a.xnorm = xnorm; a.ldx = ldx; a.Fp = layer->Fp;
    a.tile_start = plan->tile_start;
    a.ximg = reinterpret_cast<unsigned char*>(ximg);
    a.Fk = wide_fk(layer->Fp);
    a.nk2 = a.Fk / 32;
    a.ximg_bwd = a.ximg + wide_ximg_bwd_off(plan->n_tiles, a.Fk);
    count_launches(1);
    ProfScope prof("x_images", st);
    k_x_images_wide<<<dim3(plan->n_tiles, a.nk2), 512, 0, st>>>(a);
    MK_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// kernel-block images: [block][stage kk][16 KB as above], rows = slot * nk + (k - k0) (slot d = centre row), then the
// bond-support table [block][slot][half][kernel] float4.  One thread per (block, row, 8-column chunk) / per table entry.
struct WidePackArgs {
    int Fp, Fk;
    int L[4];
    WideBlocks wb;
    const float* packed[4];
    unsigned char* img;
};

__global__ void __launch_bounds__(256) k_param_pack_wide(const __grid_constant__ WidePackArgs a) {
    const int nch = a.Fk / 8;
    const int it = blockIdx.x * 256 + threadIdx.x;
    if (it >= a.wb.nb * 128 * nch) {
        int e = it - a.wb.nb * 128 * nch;
        if (e >= tile_es_f4(a.L)) return;
        float4* es = reinterpret_cast<float4*>(a.img + wide_es_off(a.wb.nb, a.Fk));
        const int e_out = e;
        for (int blk = 0; blk < a.wb.nb; ++blk) {
            const int d = a.wb.d[blk], nk = a.wb.nk[blk];
            const int cnt = d * 2 * nk;
            if (e < cnt) {
                const int kl = e % nk, sh = e / nk;
                const int L = a.L[d - 1];
                const PackedLayout pl(d, L, a.Fp);
                es[e_out] = *reinterpret_cast<const float4*>(a.packed[d - 1] + pl.es + (size_t)((sh >> 1) * L + a.wb.k0[blk] + kl) * EP +
                                                             (sh & 1) * 4);
                return;
            }
            e -= cnt;
        }
        return;
    }
    const int c = it % nch;
    const int row = (it / nch) % 128;
    const int blk = it / (nch * 128);
    const int d = a.wb.d[blk], nk = a.wb.nk[blk];
    const float* src = nullptr;
    if (row < nk * (d + 1)) {
        const int L = a.L[d - 1];
        const PackedLayout pl(d, L, a.Fp);
        src = a.packed[d - 1] + pl.sup + ((size_t)(row / nk) * L + a.wb.k0[blk] + row % nk) * a.Fp;
    }
    __align__(16) __half2 hi[4];
    __align__(16) __half2 lo[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int col = 8 * c + 2 * t;
        const float v0 = (src && col < a.Fp) ? src[col] : 0.f;
        const float v1 = (src && col + 1 < a.Fp) ? src[col + 1] : 0.f;
        tc::split_u2(v0 * WIDE_OPSCALE, v1 * WIDE_OPSCALE, hi[t], lo[t]);
    }
    unsigned char* dst = a.img + ((size_t)blk * (a.Fk / 32) + (c >> 2)) * WIDE_STAGE + wide_stage_off(row, c & 3);
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(dst + WIDE_STAGE / 4) = *reinterpret_cast<const uint4*>(lo);
    unsigned char* db = a.img + wide_img_bwd_off(a.wb.nb, a.Fk, a.L) + (size_t)blk * 8 * a.Fk * 64 + wide_bstage_off(row, c, a.Fk);
    *reinterpret_cast<uint4*>(db) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(db + a.Fk * 32) = *reinterpret_cast<const uint4*>(lo);
}

int launch_param_pack_wide(const molkgnn_layer_t* layer, cudaStream_t st) {
    if (!layer->tile_img || !wide_layer_ok(layer)) return 0;
    MK_REQUIRE((reinterpret_cast<uintptr_t>(layer->tile_img) & 127) == 0, "param_pack: tile_img must be 128-byte aligned");
    WidePackArgs a;
    a.Fp = layer->Fp; a.Fk = wide_fk(layer->Fp);
    for (int d = 0; d < 4; ++d) { a.L[d] = layer->L[d]; a.packed[d] = layer->packed[d]; }
    a.wb.build(layer->L);
    a.img = reinterpret_cast<unsigned char*>(layer->tile_img);
    const int items = a.wb.nb * 128 * (a.Fk / 8) + tile_es_f4(layer->L);
    count_launches(1);
    k_param_pack_wide<<<(items + 255) / 256, 256, 0, st>>>(a);
    MK_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---- the convolution -----------------------------------------------------------------------------------------------------
struct WideFwdArgs {
    int F, Fp, Fk, nk2;
    int L[4];
    const float* packed[4];
    long long scoff[4];
    WideBlocks wb;
    int es_off[WIDE_MAXB];                 // float4 offset of the block inside the bond-support table
    const unsigned char* img; const float4* es;
    const unsigned char* ximg;
    const TileMetaG* meta; const float* ehat_node; int n_tiles;
    const int* order; int order_grid;
    const float* x; int ldx;               // fp32 rows of the layer input (chirality gate, kernels.py:310-317)
    int is_last;
    float* sc; uint8_t* argmax; uint8_t* argmax_free; const uint8_t* argmax_in;
    int nB, nA;
    int sm_dump, sm_meta, sm_eh, sm_dup, sm_ringB, sm_ringA;
};

__device__ __forceinline__ void wf_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(WF_CONS) : "memory"); }
__device__ __forceinline__ void wf_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

// accumulator buffer `buf` -> dump[column][row]: warp (quadrant q, column block cb) moves 32 rows x 32 columns
__device__ __forceinline__ void wf_dump(float* dump, uint32_t tmem, int buf, int nn) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = warp & 3, cb = warp >> 2;
    if (cb * 32 >= nn) return;
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TNODES + cb * 32);
    tc::tmem_ld16(taddr, v);
    tc::tmem_ld16(taddr + 16, v + 16);
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(v[i]));
    float* dst = dump + (size_t)(cb * 32) * 128 + q * 32 + lane;
#pragma unroll
    for (int i = 0; i < 32; ++i) dst[i * 128] = __uint_as_float(v[i]) * (1.0f / (WIDE_OPSCALE * WIDE_OPSCALE));   // exact
}

template <int D> __device__ __forceinline__ uint32_t wf_perm_code_rt(int p) {
    uint32_t c = 0;
#pragma unroll
    for (int q = 0; q < Perm<D>::P; ++q) if (q == p) c = perm_code<D>(q);
    return c;
}

struct WFDeg { float ws, wc, we, W, rW; const int8_t* supsign; };

// the (node, kernel) pairs of one block: the reference arithmetic on the d x d similarity tile of every pair (same operation
// order as conv_fwd_tile.cu / stack_fwd_fused.cu), results to the bucket-order arrays
template <int D, bool FORCED>
__device__ __forceinline__ void wf_pairs(const WideFwdArgs& a, const TileMetaG& m, const float* dump, const float* ehS,
                                         const unsigned char* dupf, const WFDeg& w, int blk) {
    constexpr int P = Perm<D>::P;
    const int nk = a.wb.nk[blk], k0 = a.wb.k0[blk], L = a.L[D - 1];
    const float rnk = 1.0f / (float)nk;
    const float4* estab = a.es + a.es_off[blk];
    const int np = m.cnt[D - 1] * nk;
    for (int p = (int)threadIdx.x; p < np; p += WF_CONS) {
        const int ni = (int)(((float)p + 0.5f) * rnk);
        const int kl = p - ni * nk, k = k0 + kl;
        const int nl_ = m.list[D - 1][ni];
        const uint32_t nw = m.nl[nl_];
        const int e0 = m.eslot[nl_];
        const float* col0 = dump + kl;
        float T[D][D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const float* c = col0 + (int)((nw >> (8 * j)) & 0xffu) * 128;
#pragma unroll
            for (int s = 0; s < D; ++s) T[j][s] = c[s * nk];
        }
        const float cdot = col0[nl_ * 128 + D * nk];
        const size_t cidx = (size_t)a.scoff[D - 1] + (size_t)m.posl[nl_] * L + k;
        const int forced = FORCED && a.argmax_in ? (a.argmax_in[cidx] & 0x7f) : -1;
        // mean over j for every permutation: sequential sum, then true division (kernels.py:194)
        float best = 0.f, used = 0.f;
        int bi = 0;
#pragma unroll
        for (int q = 0; q < P; ++q) {
            float s = T[0][Perm<D>::at(q, 0)];
#pragma unroll
            for (int j = 1; j < D; ++j) s += T[j][Perm<D>::at(q, j)];
            s = div_deg<D>(s);
            if (q == 0 || s > best) { best = s; bi = q; }   // first maximum wins (torch.max, kernels.py:373)
            if (FORCED && q == forced) used = s;
        }
        const int free_am = bi;
        if (FORCED && forced >= 0 && forced < P) { bi = forced; best = used; }
        // bond-attribute cosine at the chosen permutation (kernels.py:382-390)
        const uint32_t code = wf_perm_code_rt<D>(bi);
        float esum = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const int s = (code >> (2 * j)) & 3;
            const float4 e0v = *reinterpret_cast<const float4*>(ehS + (size_t)(e0 + j) * EP);
            const float4 e1v = *reinterpret_cast<const float4*>(ehS + (size_t)(e0 + j) * EP + 4);
            const float4 s0 = __ldg(estab + (s * 2 + 0) * nk + kl);
            const float4 s1 = __ldg(estab + (s * 2 + 1) * nk + kl);
            float dd = 0.f;
            dd = fmaf(e0v.x, s0.x, dd); dd = fmaf(e0v.y, s0.y, dd); dd = fmaf(e0v.z, s0.z, dd); dd = fmaf(e0v.w, s0.w, dd);
            dd = fmaf(e1v.x, s1.x, dd); dd = fmaf(e1v.y, s1.y, dd); dd = fmaf(e1v.z, s1.z, dd); dd = fmaf(e1v.w, s1.w, dd);
            esum = j == 0 ? dd : esum + dd;
        }
        const float E = div_deg<D>(esum);
        float sc = div_by((best * w.ws + cdot * w.wc) + E * w.we, w.W, w.rW);
        uint8_t am = (uint8_t)bi;
        if (D == 4 && a.is_last) {
            // chirality (kernels.py:279-350): +1 if any two neighbours are identical, else sign agreement
            int chi = 1;
            if (!dupf[nl_]) chi = (m.tsg[nl_] == w.supsign[k * 12 + bi]) ? 1 : -1;
            if (chi < 0) { sc = -sc; am |= 0x80; }
        }
        a.sc[cidx] = sc;
        a.argmax[cidx] = am;
        if (FORCED && a.argmax_free) a.argmax_free[cidx] = (uint8_t)free_am;
    }
}

// chirality gate of the degree-4 nodes of a tile: any two of the four neighbour feature rows bit-equal (torch.equal,
// kernels.py:310-317); one warp per node, raw fp32 rows of the layer input
__device__ __forceinline__ void wf_dup_flags(const float* rows, int ld, int F, const TileMetaG& m, unsigned char* dupf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n4 = m.cnt[3], t0 = m.t0;
    for (int i = warp; i < n4; i += WF_CWARPS) {
        const int nl_ = m.list[3][i];
        const uint32_t w = m.nl[nl_];
        unsigned neq = 0;
        for (int f = lane; f < F; f += 32) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = __ldg(rows + (size_t)(t0 + ((w >> (8 * j)) & 0xff)) * ld + f);
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

#ifdef MK_PHASE_CLOCKS
__device__ unsigned long long g_ph_wfwd[48];     // [0..15] consumer thread 0, [16..31] ring lane of block 2 g, [32..47] MMA lane (warp 0)
#endif

template <bool FORCED>
__global__ void __launch_bounds__(WF_THREADS, 1) k_conv_fwd_wide(const __grid_constant__ WideFwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar_meta[2], bar_eh, bar_mma[4], bar_tfree[4];
    __shared__ uint64_t bar_Bfull[WF_MAXSLOT], bar_Bfree[WF_MAXSLOT], bar_Afull[2][WF_MAXSLOT], bar_Afree[2][WF_MAXSLOT];
    __shared__ uint32_t tslot;
    __shared__ WFDeg s_deg[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    MK_PH_DECL(tid == 0 || tid == WF_CONS + 32 || tid == WF_CONS + 32 * WF_RINGS)
    if (tid == 0) {
        tc::mbar_init(&bar_meta[0], 1); tc::mbar_init(&bar_meta[1], 1); tc::mbar_init(&bar_eh, 1);
        for (int i = 0; i < 4; ++i) { tc::mbar_init(&bar_mma[i], 1); tc::mbar_init(&bar_tfree[i], 1); }
        for (int i = 0; i < WF_MAXSLOT; ++i) {
            tc::mbar_init(&bar_Bfull[i], 1); tc::mbar_init(&bar_Bfree[i], 2);      // both MMA warps read a node-image stage
            for (int w = 0; w < 2; ++w) { tc::mbar_init(&bar_Afull[w][i], 1); tc::mbar_init(&bar_Afree[w][i], 1); }
        }
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(&tslot, 512);
    if (tid < 4 && a.L[tid] > 0) {
        const PackedLayout pl(tid + 1, a.L[tid], a.Fp);
        const float* pk = a.packed[tid];
        WFDeg w;
        w.ws = pk[pl.w + 0]; w.wc = pk[pl.w + 1]; w.we = pk[pl.w + 2]; w.W = pk[pl.w + 3];
        w.rW = 1.0f / w.W;
        w.supsign = reinterpret_cast<const int8_t*>(pk + pl.sign);
        s_deg[tid] = w;
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tslot;
    const TileWalk walk(a.order, a.order_grid, a.n_tiles);
    const int nb = a.wb.nb, ng = (nb + 1) >> 1, nk2 = a.nk2;
    unsigned char* ringB = smem + a.sm_ringB;
    unsigned char* ringA = smem + a.sm_ringA;        // warp W's ring at + W * nA stages

    if (warp >= WF_CWARPS && warp < WF_CWARPS + WF_RINGS) {
        // ================= ring warps: one per ring (issuing a bulk copy costs its thread ~300 cycles when many are in flight:
        // one thread for all three rings was the bottleneck of the kernel) =================
        const int R = warp - WF_CWARPS;                   // 0: node-image stages, 1 / 2: stages of block 2 g / 2 g + 1
        if (lane == 0) {
            uint32_t q = 0;
            for (int wk = 0; wk < walk.cnt; ++wk) {
                const unsigned char* xsrc = a.ximg + (size_t)walk.tile(wk) * nk2 * WIDE_STAGE;
                for (int g = 0; g < ng; ++g) {
                    const int blk = 2 * g + (R - 1);
                    if (R > 0 && blk >= nb) continue;
                    for (int kk = 0; kk < nk2; ++kk, ++q) {
                        if (R == 0) {
                            const uint32_t slot = q % (uint32_t)a.nB, use = q / (uint32_t)a.nB;
                            MK_PH(0);
                            tc::mbar_wait(&bar_Bfree[slot], (use & 1u) ^ 1u);
                            MK_PH(1);                                     // ring: waiting for a free stage
                            mbar_expect_tx(&bar_Bfull[slot], WIDE_STAGE);
                            bulk_g2s(ringB + (size_t)slot * WIDE_STAGE, xsrc + (size_t)kk * WIDE_STAGE, WIDE_STAGE, &bar_Bfull[slot]);
                        } else {
                            const int w = R - 1;
                            const uint32_t slot = q % (uint32_t)a.nA, use = q / (uint32_t)a.nA;
                            MK_PH(0);
                            tc::mbar_wait(&bar_Afree[w][slot], (use & 1u) ^ 1u);
                            MK_PH(1);
                            mbar_expect_tx(&bar_Afull[w][slot], WIDE_STAGE);
                            bulk_g2s(ringA + (size_t)(w * a.nA + (int)slot) * WIDE_STAGE,
                                     a.img + ((size_t)blk * nk2 + kk) * WIDE_STAGE, WIDE_STAGE, &bar_Afull[w][slot]);
                        }
                    }
                }
            }
        }
    } else if (warp >= WF_CWARPS + WF_RINGS) {
        // ================= MMA warps: warp W issues block 2 g + W of every pair g =================
        const int W = warp - (WF_CWARPS + WF_RINGS);
        if (lane == 0) {
            uint32_t qB = 0, qA = 0, use_t[2] = {0u, 0u};
            for (int wk = 0; wk < walk.cnt; ++wk) {
                const int b = wk & 1;
                tc::mbar_wait(&bar_meta[b], (uint32_t)(wk >> 1) & 1u);
                const int nn = reinterpret_cast<const TileMetaG*>(smem + a.sm_meta + (size_t)b * sizeof(TileMetaG))->nn;
                const uint32_t idesc = tc::idesc_f16(128, max(16, (nn + 15) & ~15), 0, 0);
                for (int g = 0; g < ng; ++g) {
                    const int blk = 2 * g + W;
                    const bool has = blk < nb;
                    const int tb = g & 1, buf = 2 * tb + W;
                    if (has) {
                        MK_PH(0);
                        tc::mbar_wait(&bar_tfree[buf], (use_t[tb] & 1u) ^ 1u);
                        MK_PH(1);                                         // MMA: waiting for the accumulator to be drained
                        ++use_t[tb];
                        tc::fence_after_sync();
                    }
                    const uint32_t d = tmem + (uint32_t)(buf * TNODES);
                    for (int kk = 0; kk < nk2; ++kk, ++qB) {
                        const uint32_t sB = qB % (uint32_t)a.nB, uB = qB / (uint32_t)a.nB;
                        MK_PH(0);
                        tc::mbar_wait(&bar_Bfull[sB], uB & 1u);
                        MK_PH(2);                                         // MMA: waiting for a node-image stage
                        if (has) {
                            const uint32_t sA = qA % (uint32_t)a.nA, uA = qA / (uint32_t)a.nA;
                            tc::mbar_wait(&bar_Afull[W][sA], uA & 1u);
                            MK_PH(3);                                     // MMA: waiting for a kernel-block stage
                            tc::fence_after_sync();
                            const uint32_t aBase = tc::smem_u32(ringA + (size_t)(W * a.nA + (int)sA) * WIDE_STAGE);
                            const uint32_t bBase = tc::smem_u32(ringB + (size_t)sB * WIDE_STAGE);
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const uint32_t o = (uint32_t)h * (WIDE_STAGE / 2);
                                const uint64_t dKh = tc::smem_desc(aBase + o, 128u, 256u), dKl = tc::smem_desc(aBase + o + WIDE_STAGE / 4, 128u, 256u);
                                const uint64_t dBh = tc::smem_desc(bBase + o, 128u, 256u), dBl = tc::smem_desc(bBase + o + WIDE_STAGE / 4, 128u, 256u);
                                tc::umma_f16(d, dKh, dBh, idesc, (kk > 0 || h > 0) ? 1u : 0u);
                                tc::umma_f16(d, dKl, dBh, idesc, 1u);
                                tc::umma_f16(d, dKh, dBl, idesc, 1u);
                            }
                            tc::umma_commit(&bar_Afree[W][sA]);
                            ++qA;
                            MK_PH(4);                                     // MMA: issue
                        }
                        tc::umma_commit(&bar_Bfree[sB]);                  // arrives when this warp's MMAs on the stage are done
                    }
                    if (has) tc::umma_commit(&bar_mma[buf]);
                }
            }
        }
    } else {
        // ================= consumers =================
        float* dump = reinterpret_cast<float*>(smem + a.sm_dump);
        float* ehS = reinterpret_cast<float*>(smem + a.sm_eh);
        unsigned char* dupf = smem + a.sm_dup;
        uint32_t cm[2] = {0u, 0u};
        auto issue_meta = [&](int wk) {
            uint64_t* bar = &bar_meta[wk & 1];
            mbar_expect_tx(bar, (uint32_t)sizeof(TileMetaG));
            bulk_g2s(smem + a.sm_meta + (size_t)(wk & 1) * sizeof(TileMetaG), a.meta + walk.tile(wk), (uint32_t)sizeof(TileMetaG), bar);
        };
        auto issue_eh = [&](int wk) {                 // bond rows of tile wk (its metadata must have landed)
            const TileMetaG* mm = reinterpret_cast<const TileMetaG*>(smem + a.sm_meta + (size_t)(wk & 1) * sizeof(TileMetaG));
            const uint32_t eb = (uint32_t)mm->ne * EP * 4u;
            mbar_expect_tx(&bar_eh, eb);
            if (eb) bulk_g2s(ehS, a.ehat_node + (size_t)mm->e0 * EP, eb, &bar_eh);
        };
        if (tid == 0 && walk.cnt > 0) {
            issue_meta(0);
            tc::mbar_wait(&bar_meta[0], 0u);
            issue_eh(0);
        }
        MK_PH(0);
        for (int wk = 0; wk < walk.cnt; ++wk) {
            const int b = wk & 1;
            if (warp == 0) {
                tc::mbar_wait(&bar_meta[b], (uint32_t)(wk >> 1) & 1u);
                tc::mbar_wait(&bar_eh, (uint32_t)wk & 1u);
            }
            wf_consumer_sync();                                   // every warp has left the previous tile
            const TileMetaG& m = *reinterpret_cast<const TileMetaG*>(smem + a.sm_meta + (size_t)b * sizeof(TileMetaG));
            const int nn = m.nn;
            if (tid == 0 && wk + 1 < walk.cnt) issue_meta(wk + 1);
            if (a.is_last && a.L[3] > 0) wf_dup_flags(a.x, a.ldx, a.F, m, dupf);
            MK_PH(1);
            for (int blk = 0; blk < nb; ++blk) {
                const int tb = (blk >> 1) & 1, buf = 2 * tb + (blk & 1);
                // ONE warp polls the mbarrier, the others sleep in the hardware barrier
                if (warp == 0) tc::mbar_wait(&bar_mma[buf], (cm[tb] >> (blk & 1)) & 1u);
                wf_consumer_sync();
                cm[tb] ^= 1u << (blk & 1);
                tc::fence_after_sync();
                MK_PH(2);                                         // waiting for the block's MMAs
                wf_dump(dump, tmem, buf, nn);
                tc::fence_before_sync();
                wf_consumer_sync();                               // dump complete, accumulator drained
                if (tid == 0) wf_arrive(&bar_tfree[buf]);
                MK_PH(3);
                const int d = a.wb.d[blk];
                switch (d) {
                    case 1: wf_pairs<1, FORCED>(a, m, dump, ehS, dupf, s_deg[0], blk); break;
                    case 2: wf_pairs<2, FORCED>(a, m, dump, ehS, dupf, s_deg[1], blk); break;
                    case 3: wf_pairs<3, FORCED>(a, m, dump, ehS, dupf, s_deg[2], blk); break;
                    default: wf_pairs<4, FORCED>(a, m, dump, ehS, dupf, s_deg[3], blk); break;
                }
                MK_PH(4);
                wf_consumer_sync();                               // pairs done: dump (and after the last block the bond rows) free
                MK_PH(5);
            }
            if (tid == 0 && wk + 1 < walk.cnt) {
                tc::mbar_wait(&bar_meta[b ^ 1], (uint32_t)((wk + 1) >> 1) & 1u);
                issue_eh(wk + 1);
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
#ifdef MK_PHASE_CLOCKS
    MK_PH_FLUSH(g_ph_wfwd + (tid == 0 ? 0 : tid == WF_CONS + 32 ? 16 : 32));
#endif
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// returns 1 if launched, 0 if the plan / layer is not eligible (the caller falls back to the bucket-order kernels), < 0 on error
int launch_conv_fwd_wide(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                         const void* ximg, int32_t is_last_layer, float* sc, int32_t sc_mode, const int64_t scoff[4],
                         uint8_t* argmax, uint8_t* argmax_free, const uint8_t* argmax_in, cudaStream_t st) {
    if (!ximg || sc_mode != 0 || !sc || !argmax || !layer->tile_img || !wide_layer_ok(layer)) return 0;
    if (!(plan->n_tiles > 0 && plan->tile_start && plan->tile_meta && plan->ehat_node && plan->tile_max_nodes <= TNODES)) return 0;
    static int s_budget = 0, s_sms = 0;
    if (!s_budget) {
        s_budget = device_max_smem_optin();
        s_sms = device_num_sms();
        MK_REQUIRE(s_budget > 0 && s_sms > 0, "conv_fwd_wide: no CUDA device");
    }
    WideFwdArgs a;
    memset(&a, 0, sizeof(a));
    a.F = layer->F; a.Fp = layer->Fp; a.Fk = wide_fk(layer->Fp); a.nk2 = a.Fk / 32;
    if (!a.wb.build(layer->L)) return 0;
    int eo = 0;
    for (int b = 0; b < a.wb.nb; ++b) { a.es_off[b] = eo; eo += a.wb.d[b] * 2 * a.wb.nk[b]; }
    for (int d = 0; d < 4; ++d) { a.L[d] = layer->L[d]; a.packed[d] = layer->packed[d]; a.scoff[d] = scoff[d]; }
    a.img = reinterpret_cast<const unsigned char*>(layer->tile_img);
    a.es = reinterpret_cast<const float4*>(a.img + wide_es_off(a.wb.nb, a.Fk));
    a.ximg = reinterpret_cast<const unsigned char*>(ximg);
    a.meta = reinterpret_cast<const TileMetaG*>(plan->tile_meta);
    a.ehat_node = plan->ehat_node;
    a.n_tiles = plan->n_tiles;
    const int grid = std::min(plan->n_tiles, s_sms);
    a.order = plan->tile_grid == grid ? plan->tile_order : nullptr;
    a.order_grid = a.order ? grid : 0;
    a.x = x; a.ldx = ldx;
    a.is_last = is_last_layer;
    a.sc = sc; a.argmax = argmax; a.argmax_free = argmax_free; a.argmax_in = argmax_in;
    int64_t off = 0;
    auto take = [&](int64_t bytes) { const int64_t o = off; off += (bytes + 127) / 128 * 128; return (int)o; };
    a.sm_dump = take((int64_t)TNODES * 128 * 4);
    a.sm_meta = take(2 * (int64_t)sizeof(TileMetaG));
    a.sm_eh = take((int64_t)TILE_ESLOTS * EP * 4);
    a.sm_dup = take(128);
    const int64_t room = ((int64_t)s_budget - 2048 - off) / WIDE_STAGE;      // static shared memory: barriers
    if (room < 6) return 0;
    a.nA = room >= 9 ? 3 : 2;
    a.nB = (int)std::min<int64_t>(WF_MAXSLOT, room - 2 * a.nA);
    a.sm_ringB = take((int64_t)a.nB * WIDE_STAGE);
    a.sm_ringA = take(2 * (int64_t)a.nA * WIDE_STAGE);
    static int64_t s_attr_dev[16] = {0};
    int64_t& s_attr = s_attr_dev[device_index()];      // function attributes are per device
    if (off > s_attr) {
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_conv_fwd_wide<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off));
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_conv_fwd_wide<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off));
        s_attr = off;
    }
    count_launches(1);
    if (argmax_in || argmax_free) k_conv_fwd_wide<true><<<grid, WF_THREADS, off, st>>>(a);      // parity harness / replay
    else k_conv_fwd_wide<false><<<grid, WF_THREADS, off, st>>>(a);
    MK_CHECK_CUDA(cudaGetLastError());
    return 1;
}

}  // namespace mk

#ifdef MK_PHASE_CLOCKS
extern "C" int molkgnn_debug_phase_clocks_wfwd(unsigned long long* out48) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out48, mk::g_ph_wfwd, sizeof(unsigned long long) * 48) != cudaSuccess) return -1;
    unsigned long long z[48] = {0};
    return cudaMemcpyToSymbol(mk::g_ph_wfwd, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif
