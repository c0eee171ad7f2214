// Layer-fused, tile-major forward of the whole conv stack (sm_100a, tcgen05 + TMEM): ONE launch runs, for every molecule
// tile, all layers of MolGCN.forward (reference KernelLayer.py:107-120: `sim_sc = layer(...)`, `h = propagate(...)`) with the
// activations of the tile resident in shared memory from the raw input rows to the final h.
//
// A tile is a run of <= 128 consecutive nodes holding whole molecules (tile.cuh), so the conv of a layer (one dense GEMM per
// kernel block on the tensor cores, T = khat . xhat^T, then the reference arithmetic per (node, kernel) pair -- kernels.py:353-425,
// exactly as in conv_fwd_tile.cu) AND the neighbour sum that follows it (KernelLayer.py:119) only touch tile-local rows: the
// next layer's normalised fp16 (hi, lo) operand image is rebuilt in place in shared memory.  HBM sees the raw x rows, the
// per-tile metadata and bond rows, the final h -- and, in training, what the backward needs: per layer the operand image
// (4 bytes per feature = the fp32 activation it replaces), the row norms and the arg-max bytes.  No score matrix, no
// intermediate activations, no per-layer launches (7 launches of the per-layer path become 1).
//
// CTA = 16 consumer warps + 2 producer warps, persistent over its tiles (TileWalk):
//   ring warp  streams the kernel-block images of every (tile, layer, block), one K step (8 KB: hi | lo) per stage, through a
//              ring of bulk copies -- a pure function of the sequence number, so it runs ahead across blocks, layers and tiles;
//   MMA warp   per (tile, layer): waits for the tile's operand image, then per block and K step 3 tcgen05.mma (hi*hi, lo*hi,
//              hi*lo) into the block's own TMEM accumulator (4 x 128 columns), releasing ring stages by tcgen05.commit;
//   consumers  per block: accumulator -> shared memory ([column][row]), one thread per (node, kernel) pair, scores into a
//              compact shared-memory tile; per layer: neighbour sum from that tile, norm, next operand image (+ its bulk
//              store to HBM for the backward).  Thread 0 also issues the bulk copies of the next layer's bond-support table
//              and of the next tile's metadata / bond rows / raw x rows at the points where their buffers fall free.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "tc.cuh"
#include "tile.cuh"

namespace mk {

bool tile_layer_ok(const molkgnn_layer_t* layer);
int tile_argmax_stride(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer);

constexpr int SF_MAXL = 4;                 // layers the fused kernel takes (kernel-parameter space)
constexpr int SF_CONS = 512;               // consumer threads
constexpr int SF_CWARPS = SF_CONS / 32;
constexpr int SF_THREADS = SF_CONS + 96;   // + ring warp + two MMA warps
constexpr int SF_MAXSTAGES = 8;

struct SFLayer {
    int F, Fp, Fk, K, Kp;
    int L[4], koff[4];
    const float* packed[4];
    const unsigned char* img_ks;           // K-step-major kernel-block images (tile.cuh)
    const float4* es_img; int es_f4;       // bond-support table of the layer
    TileBlocks tb;
    int x_one;                             // bytes of one operand image (hi or lo) of this layer's INPUT
    unsigned char* ximg; float* hnorm;     // out (training): image / row norms of this layer's input, nullable
    uint8_t* amT; int stride_am;           // out (training): tile-ordered arg-max bytes, nullable
    float* sc; uint8_t* argmax; uint8_t* argmax_free; const uint8_t* argmax_in;   // parity harness / replay, nullable
    long long scoff[4];
};

struct StackFwdArgs {
    int nl;
    SFLayer ly[SF_MAXL];
    const float* x; int ldx; int x_stage;  // raw input rows; x_stage: rows of a tile are bulk-copied to shared memory ...
    int xs_off;                            // ... behind the layer-0 image inside the operand-image buffer (free between tiles)
    const TileMetaG* meta; const float* ehat_node; int n_tiles;
    const int* order; int order_grid;
    float* hgate; int ld_hgate;            // fp32 input rows of the LAST layer (chirality gate, kernels.py:310-317); nl == 1: = x
    const float* hgate_r;
    float* h_out; int ldh;
    int nstages;
    int sm_x, sm_dump, sm_ring, sm_meta, sm_eh, sm_es, sm_sc, sm_dup;
};
static_assert(sizeof(StackFwdArgs) <= 4000, "StackFwdArgs must fit the kernel-parameter space");

struct SFSeg {                             // per (layer, block, segment) constants, shared memory
    float ws, wc, we, W, rW, rnk;
    int d, k0, nk, rowbase, L, es_off;
    const int8_t* supsign;
};

__device__ __forceinline__ void sf_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(SF_CONS) : "memory"); }
__device__ __forceinline__ void sf_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
// shared -> global bulk store (async proxy), tracked by the issuing thread's bulk group
__device__ __forceinline__ void sf_bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(tc::smem_u32(src_smem)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void sf_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void sf_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// accumulator of block `blk` -> dump[column][row]: warp (quadrant q, column block cb) moves 32 rows x 32 columns
__device__ __forceinline__ void sf_dump(float* dump, uint32_t tmem, int blk, int nn) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = warp & 3, cb = warp >> 2;
    if (cb * 32 >= nn) return;
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(blk * TNODES + cb * 32);
    tc::tmem_ld16(taddr, v);
    tc::tmem_ld16(taddr + 16, v + 16);
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(v[i]));
    float* dst = dump + (size_t)(cb * 32) * 128 + q * 32 + lane;
#pragma unroll
    for (int i = 0; i < 32; ++i) dst[i * 128] = __uint_as_float(v[i]);
}

// Transposed variant (TS kernel: D[node, kernel row], lane = node): thread = node v moves 32 consecutive kernel rows of its
// lane into dumpT[v][row], with the 16-byte chunk index XOR-swizzled by (v & 7) so that the float4 stores of a quarter warp
// (8 nodes, 512 B apart) and the pair epilogue's reads (consecutive rows of one node) are both bank-conflict free.
__device__ __forceinline__ int sf_dt_idx(int node, int r) { return node * 128 + ((((r >> 2) ^ (node & 7)) << 2) | (r & 3)); }
__device__ __forceinline__ void sf_dump_t(float* dump, uint32_t tmem, int buf, int nn, int rows) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = warp & 3, cb = warp >> 2;
    if (cb * 32 >= rows || q * 32 >= nn) return;
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 128 + cb * 32);
    tc::tmem_ld16(taddr, v);
    tc::tmem_ld16(taddr + 16, v + 16);
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(v[i]));
    const int node = q * 32 + lane;
    float* base = dump + node * 128;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(base + (((cb * 8 + i) ^ (node & 7)) << 2)) =
            make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
}

template <int D> __device__ __forceinline__ uint32_t sf_perm_code_rt(int p) {
    uint32_t c = 0;
#pragma unroll
    for (int q = 0; q < Perm<D>::P; ++q) if (q == p) c = perm_code<D>(q);
    return c;
}

struct SFPairOut { float sc; size_t cidx; int tidx; uint8_t am, free; };

// one (node, kernel) pair: the reference arithmetic on its d x d similarity tile (same operation order as conv_fwd_tile.cu)
template <int D, bool FORCED, bool TS>
__device__ __forceinline__ SFPairOut sf_pair(const SFLayer& ly, const TileMetaG& m, const float* dump, const float* ehS,
                                             const float4* estab, const unsigned char* dupf, const SFSeg& sg, int doff,
                                             bool is_last, int nl_, int kl) {
    constexpr int P = Perm<D>::P;
    const uint32_t nw = m.nl[nl_];
    const int e0 = m.eslot[nl_];
    const int k = sg.k0 + kl;
    const float* col0 = dump + sg.rowbase + kl;
    float T[D][D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const int nj = (int)((nw >> (8 * j)) & 0xffu);
        const float* c = col0 + nj * 128;
#pragma unroll
        for (int s = 0; s < D; ++s) T[j][s] = TS ? dump[sf_dt_idx(nj, sg.rowbase + s * sg.nk + kl)] : c[s * sg.nk];
    }
    const float cdot = TS ? dump[sf_dt_idx(nl_, sg.rowbase + D * sg.nk + kl)] : col0[nl_ * 128 + D * sg.nk];
    SFPairOut o;
    o.cidx = FORCED ? (size_t)ly.scoff[D - 1] + (size_t)m.posl[nl_] * sg.L + k : 0;
    o.tidx = doff + m.lidx[nl_] * sg.L + k;             // tile order: degree blocks, node-of-degree major, kernel minor
    const int forced = FORCED && ly.argmax_in ? (ly.argmax_in[o.cidx] & 0x7f) : -1;
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
    const uint32_t code = sf_perm_code_rt<D>(bi);
    float esum = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const int s = (code >> (2 * j)) & 3;
        const float4 e0v = *reinterpret_cast<const float4*>(ehS + (size_t)(e0 + j) * EP);
        const float4 e1v = *reinterpret_cast<const float4*>(ehS + (size_t)(e0 + j) * EP + 4);
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
    if (D == 4 && is_last) {
        // chirality (kernels.py:279-350): +1 if any two neighbours are identical, else sign agreement
        int chi = 1;
        if (!dupf[nl_]) chi = (m.tsg[nl_] == sg.supsign[k * 12 + bi]) ? 1 : -1;
        if (chi < 0) { sc = -sc; am |= 0x80; }
    }
    o.sc = sc; o.am = am;
    return o;
}

template <bool FORCED>
__device__ __forceinline__ void sf_store(const SFLayer& ly, float* scS, uint8_t* amT_tile, const SFPairOut& o) {
    scS[o.tidx] = o.sc;
    if (amT_tile) amT_tile[o.tidx] = o.am;
    if (FORCED) {
        if (ly.argmax_free) ly.argmax_free[o.cidx] = o.free;
        if (ly.argmax) ly.argmax[o.cidx] = o.am;
        if (ly.sc) ly.sc[o.cidx] = o.sc;
    }
}

template <int D, bool FORCED, bool TS>
__device__ __forceinline__ void sf_pairs_of_thread(const SFLayer& ly, const TileMetaG& m, const float* dump, const float* ehS,
                                                   const float4* estab, const unsigned char* dupf, const SFSeg& sg, int doff,
                                                   bool is_last, float* scS, uint8_t* amT_tile) {
    const int np = m.cnt[D - 1] * sg.nk;
    // the segment's pairs (node-major, then kernel), two per thread and iteration so that their dependency chains interleave
    for (int p = (int)threadIdx.x; p < np; p += 2 * SF_CONS) {
        const int p2 = p + SF_CONS;
        const int ni = (int)(((float)p + 0.5f) * sg.rnk);
        const SFPairOut o1 = sf_pair<D, FORCED, TS>(ly, m, dump, ehS, estab, dupf, sg, doff, is_last, m.list[D - 1][ni], p - ni * sg.nk);
        SFPairOut o2;
        const bool has2 = p2 < np;
        if (has2) {
            const int ni2 = (int)(((float)p2 + 0.5f) * sg.rnk);
            o2 = sf_pair<D, FORCED, TS>(ly, m, dump, ehS, estab, dupf, sg, doff, is_last, m.list[D - 1][ni2], p2 - ni2 * sg.nk);
        }
        sf_store<FORCED>(ly, scS, amT_tile, o1);
        if (has2) sf_store<FORCED>(ly, scS, amT_tile, o2);
    }
}

// chirality gate of the degree-4 nodes of a tile: any two of the four neighbour feature rows bit-equal (torch.equal,
// kernels.py:310-317); one warp per node, raw fp32 rows of the last layer's input from global memory
__device__ __forceinline__ void sf_dup_flags(const float* rows, int ld, int F, const TileMetaG& m, unsigned char* dupf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n4 = m.cnt[3], t0 = m.t0;
    for (int i = warp; i < n4; i += SF_CWARPS) {
        const int nl_ = m.list[3][i];
        const uint32_t w = m.nl[nl_];
        unsigned neq = 0;
        for (int f = lane; f < F; f += 32) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = __ldcg(rows + (size_t)(t0 + ((w >> (8 * j)) & 0xff)) * ld + f);
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
__device__ unsigned long long g_ph_sfwd[48];     // [0..15] consumer thread 0, [16..31] ring lane, [32..47] MMA lane
#endif

// TS = true (MOLKGNN_FWD_TS=1): the NODE image is the MMA's A operand, copied once per (tile, layer) from shared to tensor memory by
// tcgen05.cp, the kernel-block K steps of the ring are the B operand, the accumulator is D[node, kernel row] (3 buffers of 128
// columns).  Per MMA the tensor core then reads 4 KB of shared memory instead of 8 KB (A and B of an SS-mode 128 x 128 x 16 MMA
// are exactly the SM's 128 B/cycle, so every other shared-memory access -- dump, pairs, ring fill -- used to slow the MMAs).
// Measured: no faster (0.420 vs 0.393 ms per step at 4096 molecules) -- the MMAs are bound by the ~120 cycles their issuing
// thread needs per instruction, not by operand bandwidth.
// TS = false (default): both operands from shared memory, D[kernel row, node] (4 buffers), TWO issuing warps.
constexpr uint32_t SF_A_HI = 384u, SF_A_LO = 448u;     // TMEM columns of the node operand (Fk / 2 <= 56 each)

template <bool FORCED, bool TS>
__global__ void __launch_bounds__(SF_THREADS, 1) k_stack_fwd_fused(const __grid_constant__ StackFwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar_meta[2], bar_eh, bar_es, bar_x, bar_mma[4], bar_tfree[4], bar_rfull[SF_MAXSTAGES], bar_rfree[SF_MAXSTAGES];
    __shared__ uint32_t tslot;
    __shared__ SFSeg s_seg[SF_MAXL][4][TILE_MAXSEG];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    MK_PH_DECL(tid == 0 || tid == SF_CONS || tid == SF_CONS + 32)
    if (tid == 0) {
        tc::mbar_init(&bar_meta[0], 1); tc::mbar_init(&bar_meta[1], 1);
        tc::mbar_init(&bar_eh, 1); tc::mbar_init(&bar_es, 1); tc::mbar_init(&bar_x, 1);
        for (int i = 0; i < 4; ++i) { tc::mbar_init(&bar_mma[i], 1); tc::mbar_init(&bar_tfree[i], 1); }
        for (int i = 0; i < SF_MAXSTAGES; ++i) { tc::mbar_init(&bar_rfull[i], 1); tc::mbar_init(&bar_rfree[i], 1); }
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(&tslot, 512);
    // per (layer, block, segment) constants
    for (int i = tid; i < a.nl * 4 * TILE_MAXSEG; i += SF_THREADS) {
        const int l = i / (4 * TILE_MAXSEG), blk = (i / TILE_MAXSEG) % 4, si = i % TILE_MAXSEG;
        const SFLayer& ly = a.ly[l];
        if (blk < ly.tb.nb && si < ly.tb.nseg[blk]) {
            const TileSeg sg = ly.tb.seg[blk][si];
            const int L = ly.L[sg.d - 1];
            const PackedLayout pl(sg.d, L, ly.Fp);
            const float* pk = ly.packed[sg.d - 1];
            SFSeg c;
            c.ws = pk[pl.w + 0]; c.wc = pk[pl.w + 1]; c.we = pk[pl.w + 2]; c.W = pk[pl.w + 3];
            c.rW = 1.0f / c.W; c.rnk = 1.0f / (float)sg.nk;
            c.d = sg.d; c.k0 = sg.k0; c.nk = sg.nk; c.rowbase = sg.rowbase; c.L = L;
            int es_off = 0;                       // float4 offset of the segment inside the layer's bond-support table
            for (int b2 = 0; b2 <= blk; ++b2)
                for (int s2 = 0; s2 < (b2 == blk ? si : ly.tb.nseg[b2]); ++s2) es_off += ly.tb.seg[b2][s2].d * 2 * ly.tb.seg[b2][s2].nk;
            c.es_off = es_off;
            c.supsign = reinterpret_cast<const int8_t*>(pk + pl.sign);
            s_seg[l][blk][si] = c;
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tslot;
    const TileWalk walk(a.order, a.order_grid, a.n_tiles);
    const int NS = a.nstages;
    unsigned char* Xs = smem + a.sm_x;
    unsigned char* ring = smem + a.sm_ring;

    if (warp == SF_CWARPS) {
        // ================= ring warp: kernel-block images, one K step per stage =================
        if (lane == 0) {
            uint32_t q = 0;
            for (int wk = 0; wk < walk.cnt; ++wk)
                for (int l = 0; l < a.nl; ++l) {
                    const SFLayer& ly = a.ly[l];
                    const int nks = ly.Fk >> 4, nst = ly.tb.nb * nks;
                    for (int s = 0; s < nst; ++s, ++q) {
                        // stage order inside a layer: sequential (TS), or the K steps of a PAIR of blocks interleaved (the two
                        // MMA warps work on the two blocks of a pair side by side)
                        int src_stage = s;
                        if (!TS) {
                            const int p2 = s / (2 * nks), r = s - p2 * 2 * nks;
                            const bool two = 2 * p2 + 1 < ly.tb.nb;
                            src_stage = two ? (2 * p2 + (r & 1)) * nks + (r >> 1) : 2 * p2 * nks + r;
                        }
                        const uint32_t slot = q % (uint32_t)NS, use = q / (uint32_t)NS;
                        MK_PH(0);
                        tc::mbar_wait(&bar_rfree[slot], (use & 1u) ^ 1u);
                        MK_PH(1);                                         // ring: waiting for a free stage
                        mbar_expect_tx(&bar_rfull[slot], TILE_KS_BYTES);
                        bulk_g2s(ring + (size_t)slot * TILE_KS_BYTES, ly.img_ks + (size_t)src_stage * TILE_KS_BYTES, TILE_KS_BYTES, &bar_rfull[slot]);
                    }
                }
        }
    } else if (warp > SF_CWARPS) {
        // ================= MMA warps: SS mode -- warp W issues the blocks 2 p + W (one tcgen05.mma costs its issuing thread
        // ~120 cycles, twice the tensor time of a 128 x 128 x 16 MMA: two issuers keep the pipe busy); TS mode -- warp 0 alone ====
        const int W = warp - (SF_CWARPS + 1);
        if (lane == 0 && (!TS || W == 0)) {
            uint32_t q = 0, xseq = 0, use_t[4] = {0u, 0u, 0u, 0u};
            for (int wk = 0; wk < walk.cnt; ++wk) {
                const int b = wk & 1;
                tc::mbar_wait(&bar_meta[b], (uint32_t)(wk >> 1) & 1u);
                const int nn = reinterpret_cast<const TileMetaG*>(smem + a.sm_meta + (size_t)b * sizeof(TileMetaG))->nn;
                const int N = max(16, (nn + 15) & ~15);
                const uint32_t idesc = tc::idesc_f16(128, N, 0, 0);
                for (int l = 0; l < a.nl; ++l) {
                    const SFLayer& ly = a.ly[l];
                    const int nks = ly.Fk >> 4;
                    const uint32_t sbo = (uint32_t)(ly.Fk >> 3) * 128u;
                    const uint32_t xhi = tc::smem_u32(Xs), xlo = xhi + (uint32_t)ly.x_one;
                    MK_PH(0);
                    tc::mbar_wait(&bar_x, xseq & 1u);
                    MK_PH(1);                                             // MMA: waiting for the layer's operand image
                    ++xseq;
                    tc::fence_after_sync();
                    if (TS) {
                        // node image -> tensor memory (the MMAs of the previous layer are complete: the consumers waited for
                        // all of them before they built this image)
                        for (int ks = 0; ks < nks; ++ks) {
                            tc::tmem_cp_128x256b(tmem + SF_A_HI + (uint32_t)ks * 8u, tc::smem_desc(xhi + (uint32_t)ks * 256u, 128u, sbo));
                            tc::tmem_cp_128x256b(tmem + SF_A_LO + (uint32_t)ks * 8u, tc::smem_desc(xlo + (uint32_t)ks * 256u, 128u, sbo));
                        }
                    }
                    for (int blk = TS ? 0 : W; blk < ly.tb.nb; blk += TS ? 1 : 2) {
                        const int buf = TS ? blk % 3 : blk;
                        tc::mbar_wait(&bar_tfree[buf], (use_t[buf] & 1u) ^ 1u);
                        MK_PH(2);                                         // MMA: waiting for the accumulator to be drained
                        ++use_t[buf];
                        tc::fence_after_sync();
                        const uint32_t d = tmem + (uint32_t)(buf * TNODES);
                        const uint32_t idesc_b = tc::idesc_f16(128, max(16, (ly.tb.rows[blk] + 15) & ~15), 0, 0);
                        for (int ks = 0; ks < nks; ++ks) {
                            // ring stage of (blk, ks): sequential (TS) / interleaved inside the block pair (SS)
                            const bool two = (blk | 1) < ly.tb.nb;
                            const uint32_t qs = q + (uint32_t)(TS ? blk * nks + ks : (blk >> 1) * 2 * nks + (two ? 2 * ks + (blk & 1) : ks));
                            const uint32_t slot = qs % (uint32_t)NS, use = qs / (uint32_t)NS;
                            MK_PH(4);
                            tc::mbar_wait(&bar_rfull[slot], use & 1u);
                            MK_PH(3);                                     // MMA: waiting for a ring stage to land
                            const uint32_t aH = tc::smem_u32(ring + (size_t)slot * TILE_KS_BYTES), aL = aH + TILE_KS_BYTES / 2;
                            const uint64_t dKh = tc::smem_desc(aH, 128u, 256u), dKl = tc::smem_desc(aL, 128u, 256u);
                            if (TS) {
                                // same three products, same order: khat_hi . xhat_hi, khat_lo . xhat_hi, khat_hi . xhat_lo
                                const uint32_t xH = tmem + SF_A_HI + (uint32_t)ks * 8u, xL = tmem + SF_A_LO + (uint32_t)ks * 8u;
                                tc::umma_f16_ts(d, xH, dKh, idesc_b, ks > 0 ? 1u : 0u);
                                tc::umma_f16_ts(d, xH, dKl, idesc_b, 1u);
                                tc::umma_f16_ts(d, xL, dKh, idesc_b, 1u);
                            } else {
                                const uint32_t o = (uint32_t)ks * 256u;
                                const uint64_t dBh = tc::smem_desc(xhi + o, 128u, sbo), dBl = tc::smem_desc(xlo + o, 128u, sbo);
                                tc::umma_f16(d, dKh, dBh, idesc, ks > 0 ? 1u : 0u);
                                tc::umma_f16(d, dKl, dBh, idesc, 1u);
                                tc::umma_f16(d, dKh, dBl, idesc, 1u);
                            }
                            tc::umma_commit(&bar_rfree[slot]);          // the stage is free once these MMAs have read it
                        }
                        tc::umma_commit(&bar_mma[buf]);
                    }
                    q += (uint32_t)(ly.tb.nb * nks);
                }
            }
        }
    } else {
        // ================= consumers =================
        float* dump = reinterpret_cast<float*>(smem + a.sm_dump);
        float* ehS = reinterpret_cast<float*>(smem + a.sm_eh);
        float4* estab = reinterpret_cast<float4*>(smem + a.sm_es);
        float* scS = reinterpret_cast<float*>(smem + a.sm_sc);
        unsigned char* dupf = smem + a.sm_dup;
        uint32_t cm[4] = {0u, 0u, 0u, 0u}, esseq = 0;
        // thread 0 issues the bulk copies whose buffers the consumers themselves release
        auto issue_meta = [&](int wk) {
            uint64_t* bar = &bar_meta[wk & 1];
            mbar_expect_tx(bar, (uint32_t)sizeof(TileMetaG));
            bulk_g2s(smem + a.sm_meta + (size_t)(wk & 1) * sizeof(TileMetaG), a.meta + walk.tile(wk), (uint32_t)sizeof(TileMetaG), bar);
        };
        auto issue_es = [&](int l) {
            const SFLayer& ly = a.ly[l];
            mbar_expect_tx(&bar_es, (uint32_t)ly.es_f4 * 16u);
            bulk_g2s(estab, ly.es_img, (uint32_t)ly.es_f4 * 16u, &bar_es);
        };
        auto issue_tile_data = [&](int wk) {          // bond rows + raw x rows of tile wk (its metadata must have landed)
            const TileMetaG* mm = reinterpret_cast<const TileMetaG*>(smem + a.sm_meta + (size_t)(wk & 1) * sizeof(TileMetaG));
            const uint32_t eb = (uint32_t)mm->ne * EP * 4u;
            const uint32_t xb = a.x_stage ? (uint32_t)mm->nn * (uint32_t)a.ldx * 4u : 0u;
            mbar_expect_tx(&bar_eh, eb + xb);
            if (eb) bulk_g2s(ehS, a.ehat_node + (size_t)mm->e0 * EP, eb, &bar_eh);
            if (xb) bulk_g2s(Xs + a.xs_off, a.x + (size_t)mm->t0 * a.ldx, xb, &bar_eh);
        };
        if (tid == 0 && walk.cnt > 0) {
            issue_meta(0);
            issue_es(0);
            tc::mbar_wait(&bar_meta[0], 0u);
            issue_tile_data(0);
        }
        MK_PH(0);
        for (int wk = 0; wk < walk.cnt; ++wk) {
            const int b = wk & 1;
            const int tile = walk.tile(wk);
            tc::mbar_wait(&bar_meta[b], (uint32_t)(wk >> 1) & 1u);
            tc::mbar_wait(&bar_eh, (uint32_t)wk & 1u);
            const TileMetaG& m = *reinterpret_cast<const TileMetaG*>(smem + a.sm_meta + (size_t)b * sizeof(TileMetaG));
            const int nn = m.nn, t0 = m.t0;
            const int rend = min(TNODES, (nn + 15) & ~15);
            MK_PH(1);
            // ---- operand image of layer 0 from the raw x rows: a warp per row, norm with k_pad_norm's summation order ----
            {
                const SFLayer& ly = a.ly[0];
                unsigned char* Xhi = Xs;
                unsigned char* Xlo = Xs + ly.x_one;
                const float* xs = a.x_stage ? reinterpret_cast<const float*>(Xs + a.xs_off) : a.x + (size_t)t0 * a.ldx;
                // Warp w owns the 8-row group w of the image; lane = (row of the group, 8-column chunk).  A lane then stores whole
                // 16-byte core-matrix rows (rows 16 B apart, chunks 128 B apart: the 32 lanes of a store cover all banks) -- a
                // warp-per-row split would hit 4 banks with every store (the chunks of ONE row are 128 B apart).
                const int v = warp * 8 + (lane & 7), cq = lane >> 3;
                if (warp * 8 < rend) {
                    constexpr int NP = 4;                          // passes of 4 chunks: up to 128 columns
                    float xv[NP][8];
                    float ss = 0.f;
#pragma unroll
                    for (int i = 0; i < NP; ++i) {
                        const int c8 = (cq + 4 * i) * 8;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            xv[i][u] = (v < nn && c8 + u < ly.F) ? xs[(size_t)v * a.ldx + c8 + u] : 0.f;
                            ss += xv[i][u] * xv[i][u];
                        }
                    }
                    ss += __shfl_xor_sync(0xffffffffu, ss, 8);
                    ss += __shfl_xor_sync(0xffffffffu, ss, 16);
                    const float nrm = sqrtf(ss);
                    const float rinv = 1.0f / fmaxf(nrm, MOLKGNN_COS_EPS);
                    if (v < nn && cq == 0 && ly.hnorm) ly.hnorm[t0 + v] = nrm;
#pragma unroll
                    for (int i = 0; i < NP; ++i) {
                        const int c8 = (cq + 4 * i) * 8;
                        if (c8 < ly.Fk) {
                            __align__(16) __half2 hi[4];
                            __align__(16) __half2 lo[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) tc::split_u2(xv[i][2 * u] * rinv, xv[i][2 * u + 1] * rinv, hi[u], lo[u]);
                            const uint32_t off = tc::il_off(v, c8, ly.Fk);
                            *reinterpret_cast<uint4*>(Xhi + off) = *reinterpret_cast<const uint4*>(hi);
                            *reinterpret_cast<uint4*>(Xlo + off) = *reinterpret_cast<const uint4*>(lo);
                        }
                    }
                }
            }
            tc::fence_async_smem();
            sf_consumer_sync();                                   // (S0) image complete; every warp has left the previous tile
            if (tid == 0) {
                sf_arrive(&bar_x);
                if (a.ly[0].ximg) sf_bulk_s2g(a.ly[0].ximg + (size_t)tile * 2 * a.ly[0].x_one, Xs, 2u * (uint32_t)a.ly[0].x_one);
                if (wk + 1 < walk.cnt) issue_meta(wk + 1);        // into the other metadata buffer (free since S0)
            }
            MK_PH(2);
            for (int l = 0; l < a.nl; ++l) {
                const SFLayer& ly = a.ly[l];
                const bool is_last = l == a.nl - 1;
                int doff[4];
                doff[0] = 0;
#pragma unroll
                for (int d = 1; d < 4; ++d) doff[d] = doff[d - 1] + m.cnt[d - 1] * ly.L[d - 1];
                uint8_t* amT_tile = ly.amT ? ly.amT + (size_t)tile * ly.stride_am : nullptr;
                tc::mbar_wait(&bar_es, esseq & 1u);
                ++esseq;
                for (int blk = 0; blk < ly.tb.nb; ++blk) {
                    const int buf = TS ? blk % 3 : blk;
                    // ONE warp polls the mbarrier, the other 15 sleep in the hardware barrier: 512 threads spinning on
                    // try_wait take issue slots from the ring / MMA warps that they are waiting for
                    if (warp == 0) tc::mbar_wait(&bar_mma[buf], cm[buf] & 1u);
                    sf_consumer_sync();
                    ++cm[buf];
                    tc::fence_after_sync();
                    MK_PH(3);
                    if (TS) sf_dump_t(dump, tmem, buf, nn, ly.tb.rows[blk]);
                    else sf_dump(dump, tmem, blk, nn);
                    if (blk == 0 && is_last && ly.L[3] > 0) sf_dup_flags(a.hgate_r, a.ld_hgate, ly.F, m, dupf);
                    if (tid == 0) sf_bulk_wait_read();            // the image's bulk store has read shared memory (long ago)
                    tc::fence_before_sync();
                    sf_consumer_sync();                           // (S1) dump complete, accumulator drained
                    if (tid == 0) sf_arrive(&bar_tfree[buf]);
                    MK_PH(4);
                    const int nseg = ly.tb.nseg[blk];
                    for (int si = 0; si < nseg; ++si) {
                        const SFSeg sg = s_seg[l][blk][si];
                        switch (sg.d) {
                            case 1: sf_pairs_of_thread<1, FORCED, TS>(ly, m, dump, ehS, estab, dupf, sg, doff[0], is_last, scS, amT_tile); break;
                            case 2: sf_pairs_of_thread<2, FORCED, TS>(ly, m, dump, ehS, estab, dupf, sg, doff[1], is_last, scS, amT_tile); break;
                            case 3: sf_pairs_of_thread<3, FORCED, TS>(ly, m, dump, ehS, estab, dupf, sg, doff[2], is_last, scS, amT_tile); break;
                            default: sf_pairs_of_thread<4, FORCED, TS>(ly, m, dump, ehS, estab, dupf, sg, doff[3], is_last, scS, amT_tile); break;
                        }
                    }
                    MK_PH(5);
                    sf_consumer_sync();                           // (S2) scores of the block written; dump free
                    MK_PH(6);
                }
                if (tid == 0) {
                    // the bond-support table, and after the last layer the bond rows / x staging area, are free now
                    if (!is_last) issue_es(l + 1);
                    else if (wk + 1 < walk.cnt) {
                        issue_es(0);
                        tc::mbar_wait(&bar_meta[b ^ 1], (uint32_t)((wk + 1) >> 1) & 1u);
                        issue_tile_data(wk + 1);
                    }
                }
                // ---- neighbour sum h[v] = sum over in-edges of sc[source] (KernelLayer.py:119), norm, next operand image ----
                // The compact scores are first expanded into a dense [node][column] block (the accumulator dump area is free
                // now): every output row is then <= 4 conflict-free float4 row reads in edge order -- the same fp32 sums,
                // element by element, as k_propagate_tile (activations.cu).
                {
                    const int Kp = ly.Kp;
                    const int DS = Kp + 4;                         // row stride of the dense block: 8 different rows -> 8 different bank groups
                    float* dense = dump;
                    MK_PH(9);                                     // thread 0: bulk-copy issues for the next layer / tile
                    for (int i = tid; i < nn * (DS >> 2); i += SF_CONS) reinterpret_cast<float4*>(dense)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    sf_consumer_sync();
                    MK_PH(10);                                    // zero fill of the dense score block
#pragma unroll
                    for (int d = 1; d <= 4; ++d) {
                        const int L = ly.L[d - 1];
                        if (L == 0) continue;
                        const int np = m.cnt[d - 1] * L;
                        const float rL = 1.0f / (float)L;
                        const int ko = ly.koff[d - 1];
                        const float* src = scS + doff[d - 1];
#pragma unroll 4
                        for (int p = tid; p < np; p += SF_CONS) {
                            const int i = (int)(((float)p + 0.5f) * rL);
                            dense[(int)m.list[d - 1][i] * DS + ko + (p - i * L)] = src[p];
                        }
                    }
                    sf_consumer_sync();
                    MK_PH(11);                                    // expansion of the compact scores
                    const SFLayer& nx = a.ly[is_last ? l : l + 1];
                    unsigned char* Xhi = Xs;
                    unsigned char* Xlo = Xs + nx.x_one;
                    const bool gate_rows = !is_last && l + 1 == a.nl - 1 && a.hgate;
                    // Warp w owns the 8-row group w; lane = (row of the group, 8-column chunk), 4 passes of 4 chunks.  Element by
                    // element the same in-edge-order sums as k_propagate_tile; the image is stored as whole 16-byte core-matrix
                    // rows (conflict free, see the layer-0 image above).
                    const int v = warp * 8 + (lane & 7), cq = lane >> 3;
                    if (warp * 8 < rend) {
                        constexpr int NP = 4;
                        float acc[NP][8];
                        float ss = 0.f;
                        const int cnt = v < nn ? min((int)m.incnt[v], 4) : 0;
                        const uint32_t w = v < nn ? m.inl[v] : 0u;
#pragma unroll
                        for (int i = 0; i < NP; ++i) {
                            const int c8 = (cq + 4 * i) * 8;
#pragma unroll
                            for (int u = 0; u < 8; ++u) acc[i][u] = 0.f;
                            if (c8 < Kp) {
#pragma unroll
                                for (int t = 0; t < 4; ++t) {      // edge order
                                    if (t < cnt) {
                                        const float* row = dense + (int)((w >> (8 * t)) & 0xffu) * DS + c8;
                                        const float4 s0 = *reinterpret_cast<const float4*>(row);
                                        acc[i][0] += s0.x; acc[i][1] += s0.y; acc[i][2] += s0.z; acc[i][3] += s0.w;
                                        if (c8 + 4 < Kp) {
                                            const float4 s1 = *reinterpret_cast<const float4*>(row + 4);
                                            acc[i][4] += s1.x; acc[i][5] += s1.y; acc[i][6] += s1.z; acc[i][7] += s1.w;
                                        }
                                    }
                                }
                            }
#pragma unroll
                            for (int u = 0; u < 8; ++u) { if (c8 + u >= ly.K) acc[i][u] = 0.f; ss += acc[i][u] * acc[i][u]; }
                        }
                        const int gi = t0 + v;
                        if (is_last) {
                            if (v < nn) {
#pragma unroll
                                for (int i = 0; i < NP; ++i) {
                                    const int c8 = (cq + 4 * i) * 8;
                                    if (c8 < a.ldh) st4(a.h_out + (size_t)gi * a.ldh + c8, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
                                    if (c8 + 4 < a.ldh) st4(a.h_out + (size_t)gi * a.ldh + c8 + 4, make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]));
                                }
                            }
                        } else {
                            ss += __shfl_xor_sync(0xffffffffu, ss, 8);
                            ss += __shfl_xor_sync(0xffffffffu, ss, 16);
                            const float nrm = sqrtf(ss);
                            const float rinv = 1.0f / fmaxf(nrm, MOLKGNN_COS_EPS);
                            if (v < nn && cq == 0 && nx.hnorm) nx.hnorm[gi] = nrm;
#pragma unroll
                            for (int i = 0; i < NP; ++i) {
                                const int c8 = (cq + 4 * i) * 8;
                                if (gate_rows && v < nn) {
                                    if (c8 < a.ld_hgate) st4(a.hgate + (size_t)gi * a.ld_hgate + c8, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
                                    if (c8 + 4 < a.ld_hgate) st4(a.hgate + (size_t)gi * a.ld_hgate + c8 + 4, make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]));
                                }
                                if (c8 < nx.Fk) {                  // rows nn .. rend-1 (the MMA's N extent) receive zeros
                                    __align__(16) __half2 hi[4];
                                    __align__(16) __half2 lo[4];
#pragma unroll
                                    for (int u = 0; u < 4; ++u) tc::split_u2(acc[i][2 * u] * rinv, acc[i][2 * u + 1] * rinv, hi[u], lo[u]);
                                    const uint32_t off = tc::il_off(v, c8, nx.Fk);
                                    *reinterpret_cast<uint4*>(Xhi + off) = *reinterpret_cast<const uint4*>(hi);
                                    *reinterpret_cast<uint4*>(Xlo + off) = *reinterpret_cast<const uint4*>(lo);
                                }
                            }
                        }
                    }
                    MK_PH(12);                                    // neighbour sums, norms, next image
                    if (!is_last) {
                        tc::fence_async_smem();
                        sf_consumer_sync();                       // (S3) next layer's image complete
                        if (tid == 0) {
                            sf_arrive(&bar_x);
                            if (nx.ximg) sf_bulk_s2g(nx.ximg + (size_t)tile * 2 * nx.x_one, Xs, 2u * (uint32_t)nx.x_one);
                        }
                    }
                }
                MK_PH(7);
            }
        }
        if (tid == 0) sf_bulk_wait_all();
    }
    tc::fence_before_sync();
    __syncthreads();
    MK_PH(8);
#ifdef MK_PHASE_CLOCKS
    MK_PH_FLUSH(g_ph_sfwd + (tid == 0 ? 0 : tid == SF_CONS ? 16 : 32));
#endif
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ---- host side -----------------------------------------------------------------------------------------------------
// One launch for the whole stack.  Returns 1 if launched, 0 if the plan / stack is not eligible (the caller runs the per-layer
// path), < 0 on error.  `ximg[l]`, `hnorm[l]`, `amT[l]` (nullable: inference) receive what the tile backward reads;
// sc / argmax / argmax_free / argmax_in per layer serve the parity harness.
int launch_stack_fwd_fused(const molkgnn_plan_t* plan, const molkgnn_layer_t* layers, int nl, const float* x, int32_t ldx,
                           void* const* ximg, float* const* hnorm, uint8_t* const* amT, float* hgate, int32_t ld_hgate,
                           float* h_out, int32_t ldh, float* const* sc, uint8_t* const* argmax, uint8_t* const* argmax_free,
                           const uint8_t* const* argmax_in, const int64_t (*scoff)[4], cudaStream_t st) {
    if (nl < 1 || nl > SF_MAXL) return 0;
    if (!(plan->n_tiles > 0 && plan->tile_start && plan->tile_meta && plan->ehat_node && plan->tile_max_nodes <= TNODES)) return 0;
    static int s_budget = 0, s_sms = 0;
    if (!s_budget) {
        s_budget = device_max_smem_optin();
        s_sms = device_num_sms();
        MK_REQUIRE(s_budget > 0 && s_sms > 0, "stack_fwd_fused: no CUDA device");
    }
    StackFwdArgs a;
    memset(&a, 0, sizeof(a));
    a.nl = nl;
    int64_t x_max = 0, es_max = 0, sc_max = 0;
    bool forced = false;
    for (int l = 0; l < nl; ++l) {
        const molkgnn_layer_t& ly = layers[l];
        SFLayer& s = a.ly[l];
        if (!ly.tile_img || !tile_layer_ok(&ly)) return 0;
        if (!s.tb.build(ly.L) || s.tb.nb > 4) return 0;
        if (l > 0 && ly.F != layers[l - 1].K) return 0;
        s.F = ly.F; s.Fp = ly.Fp; s.Fk = tile_fk(ly.Fp); s.K = ly.K; s.Kp = (ly.K + 3) / 4 * 4;
        if (s.Kp > 128 || s.Fp > 128) return 0;
        for (int d = 0; d < 4; ++d) { s.L[d] = ly.L[d]; s.koff[d] = ly.koff[d]; s.packed[d] = ly.packed[d]; s.scoff[d] = scoff ? scoff[l][d] : 0; }
        const unsigned char* img = reinterpret_cast<const unsigned char*>(ly.tile_img);
        s.img_ks = img + tile_img_ks_off(s.tb.nb, s.Fk);
        s.es_img = reinterpret_cast<const float4*>(img + tile_es_off(s.tb.nb, s.Fk));
        s.es_f4 = tile_es_f4(ly.L);
        s.x_one = tile_img_one(s.Fk);
        s.ximg = ximg ? reinterpret_cast<unsigned char*>(ximg[l]) : nullptr;
        s.hnorm = hnorm ? hnorm[l] : nullptr;
        s.amT = amT ? amT[l] : nullptr;
        s.stride_am = tile_argmax_stride(plan, &ly);
        s.sc = sc ? sc[l] : nullptr;
        s.argmax = argmax ? argmax[l] : nullptr;
        s.argmax_free = argmax_free ? argmax_free[l] : nullptr;
        s.argmax_in = argmax_in ? argmax_in[l] : nullptr;
        if (s.sc || s.argmax || s.argmax_free || s.argmax_in) forced = true;
        x_max = std::max<int64_t>(x_max, 2 * (int64_t)s.x_one);
        es_max = std::max<int64_t>(es_max, (int64_t)s.es_f4 * 16);
        sc_max = std::max<int64_t>(sc_max, (int64_t)s.stride_am * 4);      // stride_am >= pairs of the fullest tile
        if (s.ximg && (reinterpret_cast<uintptr_t>(s.ximg) & 127)) return 0;
    }
    a.x = x; a.ldx = ldx;
    // raw x rows of the NEXT tile are staged behind the layer-0 image in the operand-image buffer (no MMA reads it between the
    // last layer's MMAs and the next tile's layer-0 image)
    a.xs_off = (2 * a.ly[0].x_one + 127) / 128 * 128;
    a.x_stage = (ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && a.xs_off + (int64_t)TNODES * ldx * 4 <= 64 * 1024) ? 1 : 0;
    if (a.x_stage) x_max = std::max<int64_t>(x_max, a.xs_off + (int64_t)TNODES * ldx * 4);
    a.meta = reinterpret_cast<const TileMetaG*>(plan->tile_meta);
    a.ehat_node = plan->ehat_node;
    a.n_tiles = plan->n_tiles;
    const int grid = std::min(plan->n_tiles, s_sms);
    a.order = plan->tile_grid == grid ? plan->tile_order : nullptr;
    a.order_grid = a.order ? grid : 0;
    if (nl == 1) { a.hgate = nullptr; a.hgate_r = x; a.ld_hgate = ldx; }
    else {
        MK_REQUIRE(hgate && ld_hgate % 4 == 0 && ld_hgate >= layers[nl - 1].F, "stack_fwd_fused: hgate buffer missing");
        a.hgate = hgate; a.hgate_r = hgate; a.ld_hgate = ld_hgate;
    }
    a.h_out = h_out; a.ldh = ldh;
    MK_REQUIRE(ldh % 4 == 0 && ldh >= layers[nl - 1].K && ldh <= 128, "stack_fwd_fused: ldh=%d", ldh);
    int64_t off = 0;
    auto take = [&](int64_t bytes) { const int64_t o = off; off += (bytes + 127) / 128 * 128; return (int)o; };
    a.sm_x = take(x_max);
    a.sm_dump = take((int64_t)TNODES * 128 * 4);
    a.sm_meta = take(2 * (int64_t)sizeof(TileMetaG));
    a.sm_eh = take((int64_t)TILE_ESLOTS * EP * 4);
    a.sm_es = take(es_max);
    a.sm_sc = take(sc_max);
    a.sm_dup = take(128);
    a.sm_ring = take(0);
    const int64_t room = (int64_t)s_budget - 5120 - off;          // static shared memory: barriers + segment constants
    a.nstages = (int)std::min<int64_t>(SF_MAXSTAGES, room / TILE_KS_BYTES);
    if (a.nstages < 3) return 0;
    off += (int64_t)a.nstages * TILE_KS_BYTES;
    static int64_t s_attr_dev[16] = {0};
    int64_t& s_attr = s_attr_dev[device_index()];      // function attributes are per device
    if (off > s_attr) {
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_stack_fwd_fused<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off));
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_stack_fwd_fused<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off));
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_stack_fwd_fused<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off));
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_stack_fwd_fused<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off));
        s_attr = off;
    }
    static int s_ts = -1;
    if (s_ts < 0) { const char* e = getenv("MOLKGNN_FWD_TS"); s_ts = (e && e[0] == '1') ? 1 : 0; }
    count_launches(1);
    ProfScope prof("stack_fwd_fused", st);
    if (s_ts) {
        if (forced) k_stack_fwd_fused<true, true><<<grid, SF_THREADS, off, st>>>(a);      // parity harness / replay
        else k_stack_fwd_fused<false, true><<<grid, SF_THREADS, off, st>>>(a);
    } else {
        if (forced) k_stack_fwd_fused<true, false><<<grid, SF_THREADS, off, st>>>(a);
        else k_stack_fwd_fused<false, false><<<grid, SF_THREADS, off, st>>>(a);
    }
    MK_CHECK_CUDA(cudaGetLastError());
    return 1;
}

}  // namespace mk

#ifdef MK_PHASE_CLOCKS
extern "C" int molkgnn_debug_phase_clocks_sfwd(unsigned long long* out48) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out48, mk::g_ph_sfwd, sizeof(unsigned long long) * 48) != cudaSuccess) return -1;
    unsigned long long z[48] = {0};
    return cudaMemcpyToSymbol(mk::g_ph_sfwd, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif
