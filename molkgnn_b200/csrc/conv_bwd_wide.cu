// Tensor-core backward for WIDE layers (sm_100a, tcgen05 + TMEM): the layers conv_fwd_wide.cu runs forward (more than 8 kernel
// blocks or more than 112 features; BASELINE configs[2]: 15 blocks x 448 feature columns).  The reference relies on autograd over
// kernels.py:353-425 and KernelLayer.py:119; gradients are routed through the SAVED arg-max permutation.
//
// Same formulation as conv_bwd_tile.cu -- per (molecule tile, kernel block) the sparse coefficient block Wt[row, v] is scattered
// into shared memory as an fp16 (hi, lo) operand and feeds two GEMMs,
//        dxh[v, f]   += Wt[:, v]^T . khat_b[:, f]        (gradient w.r.t. the normalised input rows)
//        G_b[row, f] += Wt[row, :] . xhat[:, f]          (kernel-parameter gradients)
// -- but with 448 feature columns ONE of the two accumulators fills tensor memory, so the two GEMMs run as two launches of the
// same kernel with opposite loop orders:
//   phase X  tile-major: a persistent CTA walks its tiles, per tile all blocks; dxh[128 nodes, Fk] stays in TMEM over the blocks
//            of the tile; the block's khat rows stream in as 16-row K steps; epilogue: raw dxh (transposed through shared memory
//            for coalesced rows) to grad_x, the normalisation Jacobian follows in k_wide_jacobian;
//   phase G  block-major: a CTA owns ONE kernel block and a share of the tiles; G_b[128 rows, Fk] stays in TMEM over all its
//            tiles; the tile's xhat rows stream in as 16-node K steps; epilogue: one partial copy per CTA of a block
//            (k_param_finalize, params.cu, reduces the copies in fixed order).
// A "unit" is one (tile, block): bulk copy of [tile metadata | coefficients | arg-max codes] (k_coef_tile in blocked tile order),
// scatter with one thread per (node, kernel) pair (collision chains as in conv_bwd_tile.cu: deterministic, no atomics), then
// 8 K steps x 3 MMAs (hi*hi, lo*hi, hi*lo) per 224-column half, the halves issued by two warps.  The streamed operand is an
// MN-major image in K-step-major order (k_x_images_wide / k_param_pack_wide write it next to the forward's images).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "tc.cuh"
#include "tile.cuh"

namespace mk {

bool wide_layer_ok(const molkgnn_layer_t* layer);
int64_t launch_coef_wide(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* grad, int32_t ldg, int32_t grad_mode,
                         const uint8_t* argmax, const int64_t scoff[4], float* coefT, uint8_t* amT, int stride, int stride_am,
                         float* bondP, float* amax, int grid, bool do_launch, cudaStream_t st);

constexpr int WB_CONS = 512;
constexpr int WB_CWARPS = WB_CONS / 32;
constexpr int WB_THREADS = WB_CONS + 96;      // + ring warp + two MMA warps
constexpr int WB_WT_ONE = 16 * 16 * 128;      // one fp16 image of the 128 x 128 coefficient block
constexpr int WB_MAXSTAGES = 6;

struct WideBwdArgs {
    int phase;                         // 0 = X (dxh, tile-major), 1 = G (kernel gradients, block-major)
    int F, Fp, Fk, nh, Nh;             // nh column halves of Nh columns (the last one takes the rest)
    int L[4];
    const float* packed[4];
    WideBlocks wb;
    const TileMetaG* meta; const int* tile_start; int n_tiles;
    const int* order; int order_grid;
    const unsigned char* bimg;         // X: kernel-block images [block][8 K steps][stage]; G: node images [tile][8 K steps][stage]
    int stage_bytes;                   // Fk * 64: [hi | lo] x 2 row groups x Fk/8 chunks x 128 B
    const float* coefT; const uint8_t* amT; int stride, stride_am;
    const float* amax; int namax;
    float* gx; int ldgx;               // X: raw dxh
    int cta_begin[WIDE_MAXB + 1];      // G: CTAs [cta_begin[b], cta_begin[b + 1]) work on block b
    float* partials; long long part_off[4]; int FW;
    int nstages, sm_wt, sm_unit, unit_bytes, ub_a, ub_am, sm_ring;
    int flush;                         // units per accumulation chunk (see k_conv_bwd_wide)
};

__device__ __forceinline__ void wb_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(WB_CONS) : "memory"); }
__device__ __forceinline__ void wb_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wbt_store_hl(unsigned char* wt, int row, int col, __half hi, __half lo) {
    const uint32_t off = tc::il_off(row, col, 128);
    *reinterpret_cast<__half*>(wt + off) = hi;
    *reinterpret_cast<__half*>(wt + WB_WT_ONE + off) = lo;
}
__device__ __forceinline__ void wbt_store(unsigned char* wt, int row, int col, float v) {
    const __half hi = __float2half_rn(v);
    wbt_store_hl(wt, row, col, hi, __float2half_rn(v - __half2float(hi)));
}
__device__ __forceinline__ void wbt_add(unsigned char* wt, int row, int col, float v) {
    const uint32_t off = tc::il_off(row, col, 128);
    __half* ph = reinterpret_cast<__half*>(wt + off);
    __half* pl = reinterpret_cast<__half*>(wt + WB_WT_ONE + off);
    v += __half2float(*ph) + __half2float(*pl);
    const __half hi = __float2half_rn(v);
    *ph = hi;
    *pl = __float2half_rn(v - __half2float(hi));
}

// the unit (tile, block) of a CTA's sequence
struct WBUnit { int tile, blk; };
struct WBSeq {
    int phase, nb, cnt_tiles, nunits;
    int blk_fixed, rank, cpb;
    const TileWalk* walk;
    __device__ __forceinline__ WBUnit at(int u) const {
        WBUnit r;
        if (phase == 0) { const int wk = u / nb; r.tile = walk->tile(wk); r.blk = u - wk * nb; }
        else { r.tile = rank + u * cpb; r.blk = blk_fixed; }
        return r;
    }
};

// Accumulation chunks (see k_conv_bwd_wide): position of a unit = block index inside the tile (phase X) / unit index (phase G);
// the chunks of column half h are shifted by h * flush / 2 positions, so that the two halves are never flushed together
// (while one half leaves tensor memory, the other half's MMA warp keeps the tensor cores busy).
__device__ __forceinline__ bool wb_fresh(int flush, int pos, int h) { return pos == 0 || (pos + h * (flush >> 1)) % flush == 0; }
__device__ __forceinline__ bool wb_chunk_end(int flush, int pos, int last, int h) { return pos == last || (pos + 1 + h * (flush >> 1)) % flush == 0; }
__device__ __forceinline__ bool wb_chunk_add(int flush, int pos, int h) { return (pos + h * (flush >> 1)) / flush > 0; }

// 32 x 32 fp32 transpose through a per-warp shared-memory tile in 16-byte units: lane r writes its 32 columns as 8 float4 (the
// chunk index XOR-swizzled by r & 7: conflict free per quarter warp), then lane l reads the float4 (row 4 i + l / 8, chunk l & 7):
// a warp instruction covers 4 rows x 128 contiguous bytes of the result.
__device__ __forceinline__ void wb_tile_put(float* tb, int lane, const uint32_t* u, float scale) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(tb + lane * 32 + ((j ^ (lane & 7)) << 2)) =
            make_float4(__uint_as_float(u[4 * j]) * scale, __uint_as_float(u[4 * j + 1]) * scale, __uint_as_float(u[4 * j + 2]) * scale,
                        __uint_as_float(u[4 * j + 3]) * scale);
}
__device__ __forceinline__ float4 wb_tile_get4(const float* tb, int row, int j) {
    return *reinterpret_cast<const float4*>(tb + row * 32 + ((j ^ (row & 7)) << 2));
}
// fire-and-forget 16-byte reduction (red.global.add.v4.f32, sm_90+): the L2 does the four round-to-nearest adds
__device__ __forceinline__ void wb_red4(float* dst, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

#ifdef MK_PHASE_CLOCKS
__device__ unsigned long long g_ph_wbwd[2][48];   // [phase][0..15 consumer thread 0 | 16..31 ring lane | 32..47 MMA lane (warp 0)]
#endif

__global__ void __launch_bounds__(WB_THREADS, 1) k_conv_bwd_wide(const __grid_constant__ WideBwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar_cp[2], bar_wt[2], bar_mma[2], bar_fl[2], bar_full[WB_MAXSTAGES], bar_free[WB_MAXSTAGES];
    __shared__ uint32_t tslot;
    __shared__ unsigned char s_lut[4][12];               // packed permutation codes (2 bits per j) per degree
    __shared__ float s_alpha[4], s_beta[4], s_gmax;
    __shared__ int s_rowoff[128];                         // phase G: float offset of the block's row r inside a partial copy
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NI = a.nh;                                  // issuing warps = column halves
    MK_PH_DECL(tid == 0 || tid == WB_CONS || tid == WB_CONS + 32)
    if (tid == 0) {
        tc::mbar_init(&bar_cp[0], 1); tc::mbar_init(&bar_cp[1], 1);
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&bar_wt[i], 1); tc::mbar_init(&bar_mma[i], (uint32_t)NI); tc::mbar_init(&bar_fl[i], 1); }
        for (int i = 0; i < WB_MAXSTAGES; ++i) { tc::mbar_init(&bar_full[i], 1); tc::mbar_init(&bar_free[i], (uint32_t)NI); }
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(&tslot, 512);
    if (tid < 48) {
        s_lut[tid / 12][tid % 12] = c_perm_code[tid / 12][tid % 12];
    }
    if (tid >= 64 && tid < 68 && a.L[tid - 64] > 0) {
        const int d = tid - 63;
        const PackedLayout pl(d, a.L[d - 1], a.Fp);
        const float* pk = a.packed[d - 1];
        s_alpha[d - 1] = pk[pl.w + 0] / pk[pl.w + 3] / (float)d;      // w_s / (d W)
        s_beta[d - 1] = pk[pl.w + 1] / pk[pl.w + 3];                  // w_c / W
    }
    if (warp == 3) {                                      // max |coef| over the per-CTA values of k_coef_tile
        float gm = 0.f;
        for (int i = lane; i < a.namax; i += 32) gm = fmaxf(gm, __ldg(a.amax + i));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
        if (lane == 0) s_gmax = gm;
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tslot;
    // power-of-two scale: |alpha * chi * g| / scale <= 2^10
    float scale, rscale;
    {
        const float gm = fmaxf(s_gmax, 1e-30f);
        int e;
        frexpf(gm, &e);
        scale = ldexpf(1.0f, e - 10) * (1.0f / WIDE_OPSCALE);      // the streamed operand carries WIDE_OPSCALE (tile.cuh)
        rscale = ldexpf(1.0f, 10 - e);
    }
    // ---- this CTA's unit sequence ----
    const TileWalk walk(a.order, a.order_grid, a.n_tiles);
    WBSeq seq;
    seq.phase = a.phase; seq.nb = a.wb.nb; seq.walk = &walk;
    seq.blk_fixed = 0; seq.rank = 0; seq.cpb = 1;
    if (a.phase == 0) { seq.cnt_tiles = walk.cnt; seq.nunits = walk.cnt * a.wb.nb; }
    else {
        int b = 0;
        while (b + 1 < a.wb.nb && (int)blockIdx.x >= a.cta_begin[b + 1]) ++b;
        seq.blk_fixed = b; seq.rank = (int)blockIdx.x - a.cta_begin[b]; seq.cpb = a.cta_begin[b + 1] - a.cta_begin[b];
        seq.cnt_tiles = seq.rank < a.n_tiles ? (a.n_tiles - seq.rank + seq.cpb - 1) / seq.cpb : 0;
        seq.nunits = seq.cnt_tiles;
    }
    if (a.phase == 1 && tid < 128) {
        const int blk = seq.blk_fixed, nk = a.wb.nk[blk];
        const int slot = tid / nk;
        s_rowoff[tid] = (slot * a.L[a.wb.d[blk] - 1] + a.wb.k0[blk] + tid - slot * nk) * a.FW;
    }
    __syncthreads();
    const int NS = a.nstages;
    unsigned char* wt = smem + a.sm_wt;
    unsigned char* ring = smem + a.sm_ring;
    const int chunk_bytes = (a.Fk >> 3) * 128;            // one 8-row group of a stage half

    if (warp == WB_CWARPS) {
        // ================= ring warp: K steps of the streamed operand =================
        if (lane == 0) {
            uint32_t q = 0;
            for (int u = 0; u < seq.nunits; ++u) {
                const WBUnit un = seq.at(u);
                int nks;
                const unsigned char* src;
                if (a.phase == 0) {
                    nks = ((a.wb.nk[un.blk] * (a.wb.d[un.blk] + 1) + 15) >> 4);
                    src = a.bimg + (size_t)un.blk * 8 * a.stage_bytes;
                } else {
                    const int nn = __ldg(a.tile_start + un.tile + 1) - __ldg(a.tile_start + un.tile);
                    nks = max(1, (nn + 15) >> 4);
                    src = a.bimg + (size_t)un.tile * 8 * a.stage_bytes;
                }
                for (int ks = 0; ks < nks; ++ks, ++q) {
                    const uint32_t slot = q % (uint32_t)NS, use = q / (uint32_t)NS;
                    MK_PH(0);
                    tc::mbar_wait(&bar_free[slot], (use & 1u) ^ 1u);
                    MK_PH(1);                                         // ring: waiting for a free stage
                    mbar_expect_tx(&bar_full[slot], (uint32_t)a.stage_bytes);
                    bulk_g2s(ring + (size_t)slot * a.stage_bytes, src + (size_t)ks * a.stage_bytes, (uint32_t)a.stage_bytes, &bar_full[slot]);
                }
            }
        }
    } else if (warp > WB_CWARPS) {
        // ================= MMA warps: warp W issues the column half W =================
        const int W = warp - (WB_CWARPS + 1);
        if (lane == 0 && W < NI) {
            uint32_t q = 0, nfl = 0;
            const int ncol = W == NI - 1 ? a.Fk - W * a.Nh : a.Nh;
            const uint32_t dcol = tmem + (uint32_t)(W * a.Nh);
            const uint32_t boff = (uint32_t)(W * (a.Nh >> 3)) * 128u;      // first 8-column chunk of this half inside a row group
            const uint32_t idesc = a.phase == 0 ? tc::idesc_f16(128, ncol, 1, 1)      // A = Wt MN-major (M = node, K = row)
                                                : tc::idesc_f16(128, ncol, 0, 1);     // A = Wt K-major (M = row, K = node)
            for (int u = 0; u < seq.nunits; ++u) {
                const WBUnit un = seq.at(u);
                MK_PH(0);
                tc::mbar_wait(&bar_wt[u & 1], (uint32_t)(u >> 1) & 1u);
                MK_PH(1);                                             // MMA: waiting for the coefficient block
                tc::fence_after_sync();
                const uint32_t whi = tc::smem_u32(wt + (size_t)(u & 1) * 2 * WB_WT_ONE), wlo = whi + WB_WT_ONE;
                int nks;
                bool fresh;
                if (a.phase == 0) { nks = ((a.wb.nk[un.blk] * (a.wb.d[un.blk] + 1) + 15) >> 4); fresh = wb_fresh(a.flush, un.blk, W); }
                else {
                    const int nn = __ldg(a.tile_start + un.tile + 1) - __ldg(a.tile_start + un.tile);
                    nks = max(1, (nn + 15) >> 4);
                    fresh = wb_fresh(a.flush, u, W);
                }
                if (fresh && u > 0) {                                 // the previous chunk of this half has left tensor memory
                    tc::mbar_wait(&bar_fl[W], nfl & 1u);
                    ++nfl;
                    tc::fence_after_sync();
                }
                for (int ks = 0; ks < nks; ++ks, ++q) {
                    const uint32_t slot = q % (uint32_t)NS, use = q / (uint32_t)NS;
                    MK_PH(0);
                    tc::mbar_wait(&bar_full[slot], use & 1u);
                    MK_PH(2);                                         // MMA: waiting for a stage
                    tc::fence_after_sync();
                    const uint32_t bhi = tc::smem_u32(ring + (size_t)slot * a.stage_bytes) + boff, blo = bhi + (uint32_t)(a.stage_bytes >> 1);
                    // streamed operand, MN-major: K groups (8 rows) chunk_bytes apart, 8-column chunks 128 B apart
                    const uint64_t dBh = tc::smem_desc(bhi, (uint32_t)chunk_bytes, 128u), dBl = tc::smem_desc(blo, (uint32_t)chunk_bytes, 128u);
                    uint64_t dAh, dAl;
                    if (a.phase == 0) {          // Wt MN-major: +2 row groups (2 * 2048 B) per K step
                        dAh = tc::smem_desc(whi + ks * 4096u, 2048u, 128u); dAl = tc::smem_desc(wlo + ks * 4096u, 2048u, 128u);
                    } else {                     // Wt K-major: +256 B per K step (two 8-column chunks)
                        dAh = tc::smem_desc(whi + ks * 256u, 128u, 2048u); dAl = tc::smem_desc(wlo + ks * 256u, 128u, 2048u);
                    }
                    tc::umma_f16(dcol, dAh, dBh, idesc, (fresh && ks == 0) ? 0u : 1u);
                    tc::umma_f16(dcol, dAl, dBh, idesc, 1u);
                    tc::umma_f16(dcol, dAh, dBl, idesc, 1u);
                    tc::umma_commit(&bar_free[slot]);
                }
                tc::umma_commit(&bar_mma[u & 1]);
                MK_PH(3);                                             // MMA: issue
            }
        }
    } else {
        // ================= consumers =================
        const int q = warp & 3, cpart = warp >> 2;            // TMEM lane quadrant / column part of this warp
        // thread 0: tile metadata of unit u -> unit buffer u & 1.  The unit's coefficients / arg-max codes (one contiguous run of
        // the blocked tile order, L2 resident: k_coef_tile just wrote them) are read straight from global memory by the scatter --
        // staging them in shared memory cost the room of a third ring stage, and it is the ring depth that bounds the kernel
        auto issue_unit = [&](int u) {
            const WBUnit un = seq.at(u);
            uint64_t* bar = &bar_cp[u & 1];
            mbar_expect_tx(bar, (uint32_t)sizeof(TileMetaG));
            bulk_g2s(smem + a.sm_unit + (size_t)(u & 1) * a.unit_bytes, a.meta + un.tile, (uint32_t)sizeof(TileMetaG), bar);
        };
        // Accumulation chunks.  The tensor core TRUNCATES every accumulate (measured: the raw sums come out smaller in magnitude by
        // ~2^-24 per tcgen05.mma that added to them), so a chain of 360 (phase X: 15 blocks x 24) or 2000+ (phase G: all tiles of
        // the CTA) MMAs into one accumulator misses 1e-5.  The accumulator is therefore FLUSHED every `flush` units -- at most 96
        // MMAs per chain, as in conv_bwd_tile.cu -- into the fp32 result in global memory (round-to-nearest adds; the rows are
        // L2 resident between the flushes of a tile / of a CTA's partial copy).
        // end_of_chunk(u): unit u is the last of its chunk; flush_unit(u): accumulator -> global (first chunk: store, later: add).
        auto upos = [&](int u) { return a.phase == 0 ? u % a.wb.nb : u; };
        const int last_pos = a.phase == 0 ? a.wb.nb - 1 : seq.nunits - 1;
        auto flush_unit = [&](int u, int h) {
            float* tb = reinterpret_cast<float*>(wt + (size_t)(u & 1) * 2 * WB_WT_ONE) + warp * 1024;   // the unit's Wt buffer is free: transposition buffer
            const WBUnit pu = seq.at(u);
            const bool add = wb_chunk_add(a.flush, upos(u), h);
            const int cbeg = h * (a.Nh >> 5), cend = (h == NI - 1 ? a.Fk : (h + 1) * a.Nh) >> 5;      // 32-column chunks of the half
            int t0p = 0, nnp = 0, nk = 1, k0 = 0, L = 0, rows = 0;
            float* part = nullptr;
            if (a.phase == 0) { t0p = __ldg(a.tile_start + pu.tile); nnp = __ldg(a.tile_start + pu.tile + 1) - t0p; }
            else {
                const int d = a.wb.d[pu.blk];
                nk = a.wb.nk[pu.blk]; k0 = a.wb.k0[pu.blk]; L = a.L[d - 1]; rows = nk * (d + 1);
                part = a.partials + a.part_off[d - 1] + (size_t)seq.rank * (d + 1) * L * a.FW;
            }
            for (int c = cbeg + cpart; c < cend; c += 4) {
                uint32_t v[32];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
                tc::tmem_ld16(taddr, v);
                tc::tmem_ld16(taddr + 16, v + 16);
                tc::tmem_ld_wait();
                wb_tile_put(tb, lane, v, scale);
                __syncwarp();
                // Later chunks ADD with fire-and-forget reductions (the L2 does the round-to-nearest adds, no load latency on the SM);
                // every element has ONE writer thread and the reductions of a thread to one address arrive in program order, so
                // the result is deterministic.
                const int j = lane & 7, f = c * 32 + 4 * j;
                if (f < a.Fp) {
                    const int rmax = a.phase == 0 ? min(32, nnp - q * 32) : min(32, rows - q * 32);
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int r = it * 4 + (lane >> 3);
                        if (r < rmax) {
                            const float4 val = wb_tile_get4(tb, r, j);
                            float* dst = a.phase == 0 ? a.gx + (size_t)(t0p + q * 32 + r) * a.ldgx + f : part + s_rowoff[q * 32 + r] + f;
                            if (add) wb_red4(dst, val);
                            else __stcg(reinterpret_cast<float4*>(dst), val);
                        }
                    }
                }
                __syncwarp();
            }
            tc::fence_before_sync();
            wb_consumer_sync();
        };
        if (tid == 0 && seq.nunits > 0) issue_unit(0);
        MK_PH(0);
        for (int u = 0; u < seq.nunits; ++u) {
            const WBUnit un = seq.at(u);
            const int blk = un.blk, d = a.wb.d[blk], nk = a.wb.nk[blk], k0 = a.wb.k0[blk];
            unsigned char* ub = smem + a.sm_unit + (size_t)(u & 1) * a.unit_bytes;
            const TileMetaG& m = *reinterpret_cast<const TileMetaG*>(ub);
            // Two Wt buffers: unit u is scattered while the tensor cores work on unit u - 1.  Buffer u & 1 is free once the MMAs
            // of unit u - 2 are done; this unit's copies have landed (one warp polls, the others sleep in the hardware barrier)
            unsigned char* wtb = wt + (size_t)(u & 1) * 2 * WB_WT_ONE;
            if (warp == 0) {
                if (u >= 2) tc::mbar_wait(&bar_mma[u & 1], (uint32_t)((u - 2) >> 1) & 1u);
                tc::mbar_wait(&bar_cp[u & 1], (uint32_t)(u >> 1) & 1u);
            }
            wb_consumer_sync();
            MK_PH(1);                                     // waiting for the MMAs of unit u - 2 / this unit's copies
            if (tid == 0 && u + 1 < seq.nunits) issue_unit(u + 1);      // the other unit buffer is free since the sync above
            // ---- clear Wt ----
            for (int i = tid * 16; i < 2 * WB_WT_ONE; i += WB_CONS * 16) *reinterpret_cast<uint4*>(wtb + i) = make_uint4(0, 0, 0, 0);
            wb_consumer_sync();
            MK_PH(2);                                     // Wt clear
            // ---- rank 0: one thread per (node, kernel) pair -- centre entry, collision-free support entries ----
            {
                int base = m.cnt[d - 1] * k0;                 // first pair of the block in the tile's blocked order
                for (int dd = 1; dd < d; ++dd) base += m.cnt[dd - 1] * a.L[dd - 1];
                const float* a_s = a.coefT + (size_t)un.tile * a.stride + base;
                const unsigned char* am_s = a.amT + (size_t)un.tile * a.stride_am + base;
                const float alpha = s_alpha[d - 1], beta = s_beta[d - 1];
                const float rnk = 1.0f / (float)nk;
                const int np = m.cnt[d - 1] * nk;
                const unsigned char* lut = s_lut[d - 1];
                for (int p = tid; p < np; p += WB_CONS) {
                    const int ni = (int)(((float)p + 0.5f) * rnk);
                    const int kl = p - ni * nk;
                    const int nl_ = m.list[d - 1][ni];
                    const float av = __ldg(a_s + p) * rscale;
                    const uint32_t code = lut[__ldg(am_s + p) & 0x7f];
                    const uint32_t nw = m.nl[nl_];
                    const uint32_t cr = m.cr[nl_];
                    wbt_store(wtb, d * nk + kl, nl_, av * beta);
                    const float as = av * alpha;              // the same value goes to all d support entries: split it once
                    const __half ah = __float2half_rn(as);
                    const __half al = __float2half_rn(as - __half2float(ah));
                    for (int j = 0; j < d; ++j)
                        if (((cr >> (2 * j)) & 3u) == 0u)
                            wbt_store_hl(wtb, (int)((code >> (2 * j)) & 3u) * nk + kl, (int)((nw >> (8 * j)) & 0xffu), ah, al);
                }
                MK_PH(3);                                 // rank-0 scatter
                // ---- collision chains: one thread per (chain, kernel) adds the followers in in-edge order after ONE barrier ----
                const int c0 = m.choff[d - 1], nch = m.choff[d] - c0;
                if (nch > 0) {
                    wb_consumer_sync();
                    for (int p = tid; p < nch * nk; p += WB_CONS) {
                        const int ci = (int)(((float)p + 0.5f) * rnk);
                        const int kl = p - ci * nk;
                        const uint32_t ch = m.chains[c0 + ci];
                        const int nf = (int)((ch >> 27) & 3u) + 1;
                        for (int f = 0; f < nf; ++f) {
                            const int ent = (int)((ch >> (9 * f)) & 0x1ffu);
                            const int nl_ = ent >> 2, j = ent & 3;
                            const int pi = m.lidx[nl_] * nk + kl;
                            const int s = (lut[__ldg(am_s + pi) & 0x7f] >> (2 * j)) & 3;
                            wbt_add(wtb, s * nk + kl, (int)((m.nl[nl_] >> (8 * j)) & 0xffu), (__ldg(a_s + pi) * rscale) * alpha);
                        }
                    }
                }
                MK_PH(4);                                 // collision chains
            }
            tc::fence_async_smem();
            wb_consumer_sync();
            MK_PH(5);
            if (tid == 0) wb_arrive(&bar_wt[u & 1]);      // the MMA warps may start this unit (a half whose chunk ended: after its flush)
            // ---- the previous unit closed an accumulation chunk of a column half: its MMAs must be done, then that half of the
            // accumulator leaves tensor memory; the other half's MMA warp is already working on this unit ----
            if (u > 0) {
                bool e[2];
                e[0] = wb_chunk_end(a.flush, upos(u - 1), last_pos, 0);
                e[1] = NI > 1 && wb_chunk_end(a.flush, upos(u - 1), last_pos, 1);
                if (e[0] || e[1]) {
                    if (warp == 0) tc::mbar_wait(&bar_mma[(u - 1) & 1], (uint32_t)((u - 1) >> 1) & 1u);
                    wb_consumer_sync();
                    tc::fence_after_sync();
                    MK_PH(7);                             // waiting for the chunk's last MMAs
                    for (int h = 0; h < NI; ++h)
                        if (e[h]) {
                            flush_unit(u - 1, h);
                            if (tid == 0) wb_arrive(&bar_fl[h]);
                        }
                    MK_PH(6);                             // flush
                }
            }
        }
        // ---- tail: the last unit's MMAs, then the accumulator that is still in tensor memory ----
        if (seq.nunits > 0) {
            if (warp == 0) tc::mbar_wait(&bar_mma[(seq.nunits - 1) & 1], (uint32_t)((seq.nunits - 1) >> 1) & 1u);
            wb_consumer_sync();
            tc::fence_after_sync();
        }
        MK_PH(1);
        if (seq.nunits > 0)
            for (int h = 0; h < NI; ++h) flush_unit(seq.nunits - 1, h);
        if (a.phase == 1) {
            // one partial copy per CTA of the block.  A CTA without tiles never ran an MMA: its sums are zero.  Bond columns: zeros
            // for the copies >= 1 (copy 0 receives the reduced bond sums from k_wide_bond_reduce)
            const int blk = seq.blk_fixed, d = a.wb.d[blk], nk = a.wb.nk[blk], k0 = a.wb.k0[blk], L = a.L[d - 1];
            const int rows = nk * (d + 1);
            float* part = a.partials + a.part_off[d - 1] + (size_t)seq.rank * (d + 1) * L * a.FW;
            if (seq.nunits == 0) {
                for (int i = tid; i < rows * a.Fp; i += WB_CONS) {
                    const int row = i / a.Fp, f = i - row * a.Fp;
                    const int slot = row / nk, k = k0 + row - slot * nk;
                    part[((size_t)slot * L + k) * a.FW + f] = 0.f;
                }
            }
            if (seq.rank > 0) {
                for (int i = tid; i < rows * EP; i += WB_CONS) {
                    const int row = i / EP, e = i - row * EP;
                    const int slot = row / nk, k = k0 + row - slot * nk;
                    part[((size_t)slot * L + k) * a.FW + a.Fp + e] = 0.f;
                }
            }
        }
        MK_PH(7);                                         // final epilogue
    }
    tc::fence_before_sync();
    __syncthreads();
#ifdef MK_PHASE_CLOCKS
    MK_PH_FLUSH(g_ph_wbwd[a.phase] + (tid == 0 ? 0 : tid == WB_CONS ? 16 : 32));
#endif
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// bond-attribute sums: the `nc` per-CTA copies of k_coef_tile (bondP: [degree][CTA][(d + 1) L rows][8]) reduced in CTA order into
// the bond columns of partial copy 0
struct BondReduceArgs {
    const float* bondP; int nc;
    int L[4];
    float* partials; long long part_off[4]; int FW, Fp;
};
__global__ void __launch_bounds__(256) k_wide_bond_reduce(const BondReduceArgs a) {
    int i = blockIdx.x * 256 + threadIdx.x;
    long long boff = 0;
    for (int d = 1; d <= 4; ++d) {
        const int rows_x = (d + 1) * a.L[d - 1];
        const int n = rows_x * EP;
        if (i < n) {
            const int row = i / EP, e = i - row * EP;
            float s = 0.f;
            for (int c = 0; c < a.nc; ++c) s += a.bondP[boff + ((size_t)c * rows_x + row) * EP + e];
            a.partials[a.part_off[d - 1] + (size_t)row * a.FW + a.Fp + e] = s;
            return;
        }
        i -= n;
        boff += (long long)a.nc * rows_x * EP;
    }
}

// chain rule through xhat = x / max(|x|, eps), in place on the raw dxh rows:  gx = (g - (xhat . g) xhat) / |x|;  a warp per row
struct JacArgs { float* gx; int ldgx; const float* x; int ldx; const float* xnorm; int N, F, Fp; };
__global__ void __launch_bounds__(256) k_wide_jacobian(const JacArgs a) {
    const int lane = threadIdx.x & 31;
    const int v = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (v >= a.N) return;
    const float nrm = a.xnorm[v];
    const float rden = 1.0f / fmaxf(nrm, MOLKGNN_COS_EPS);
    const bool clamped = !(nrm > MOLKGNN_COS_EPS);
    float* g = a.gx + (size_t)v * a.ldgx;
    const float* xr = a.x + (size_t)v * a.ldx;
    float gv[16], xh[16];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int f = lane + 32 * i;
        gv[i] = 0.f; xh[i] = 0.f;
        if (f < a.Fp) { gv[i] = g[f]; xh[i] = xr[f] * rden; dot = fmaf(gv[i], xh[i], dot); }
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int f = lane + 32 * i;
        if (f < a.ldgx) {
            float o = clamped ? gv[i] * rden : (gv[i] - dot * xh[i]) * rden;
            if (f >= a.F) o = 0.f;
            g[f] = o;
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------
static void wide_coef_strides(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, int* stride, int* stride_am) {
    int64_t st = 0;
    for (int d = 0; d < 4; ++d) st += (int64_t)plan->tile_max_deg[d] * layer->L[d];
    *stride = (int)((st + 3) / 4 * 4);
    *stride_am = (*stride + 15) / 16 * 16;
}

static bool wide_plan_ok(const molkgnn_plan_t* plan) {
    return plan->n_tiles > 0 && plan->tile_start && plan->tile_meta && plan->ehat_node && plan->tile_max_nodes <= TNODES;
}

bool wide_bwd_ok(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer) {
    if (!wide_plan_ok(plan) || !layer->tile_img || !wide_layer_ok(layer)) return false;
    int rows = 0;
    for (int d = 0; d < 4; ++d) rows += layer->L[d] * (d + 1);
    return rows <= 4 * 512 && wide_fk(layer->Fp) <= 512;
}

// CTAs per block of every degree for phase G: in proportion to the cost of a (tile, block) unit, the same for the blocks of a degree
static void wide_bwd_partition(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, int sms, const WideBlocks& wb, int cpb[4]) {
    int nblk[4] = {0, 0, 0, 0};
    double w[4] = {0, 0, 0, 0};
    for (int b = 0; b < wb.nb; ++b) ++nblk[wb.d[b] - 1];
    for (int d = 0; d < 4; ++d) {
        cpb[d] = 0;
        if (!nblk[d]) continue;
        const double pairs = (double)plan->n[d] / std::max(1, plan->n_tiles) * ((double)layer->L[d] / nblk[d]);
        static double s_w0 = -1.0, s_w1 = 0.0;             // MOLKGNN_WIDE_GW="fixed,per_pair" overrides the cost model
        if (s_w0 < 0.0) {
            s_w0 = 14000.0; s_w1 = 3.0;      // measured: 2.97 ms for the 5 layers of configs[2]; (7000, 6) gave 4.1 ms
            if (const char* e = getenv("MOLKGNN_WIDE_GW")) { s_w0 = std::max(1.0, atof(e)); const char* c = strchr(e, ','); if (c) s_w1 = atof(c + 1); }
        }
        w[d] = s_w0 + s_w1 * pairs;                        // ~cycles per unit (phase clocks): clear + MMAs + flush share + barriers, scatter / chains per pair
        cpb[d] = 1;
    }
    auto used = [&] { int u = 0; for (int d = 0; d < 4; ++d) u += nblk[d] * cpb[d]; return u; };
    while (true) {
        int best = -1;
        double bl = 0.0;
        for (int d = 0; d < 4; ++d)
            if (nblk[d] && used() + nblk[d] <= sms && w[d] / cpb[d] > bl) { bl = w[d] / cpb[d]; best = d; }
        if (best < 0) break;
        ++cpb[best];
    }
}

int64_t wide_bwd_partial_floats(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer) {
    if (!wide_bwd_ok(plan, layer)) return 0;
    WideBlocks wb;
    wb.build(layer->L);
    int cpb[4];
    wide_bwd_partition(plan, layer, device_num_sms(), wb, cpb);
    int64_t tot = 0, rows = 0;
    for (int d = 0; d < 4; ++d) {
        tot += (int64_t)cpb[d] * (d + 2) * layer->L[d] * (layer->Fp + EP);
        rows += (int64_t)(d + 2) * layer->L[d];
    }
    return tot + 2 * rows + 16 + 512;
}

// floats of the `coef` scratch: blocked tile-ordered coefficients, arg-max codes, per-CTA bond sums
int64_t wide_bwd_coef_floats(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer) {
    if (!wide_bwd_ok(plan, layer)) return 0;
    int stride, stride_am;
    wide_coef_strides(plan, layer, &stride, &stride_am);
    const int grid = std::max(1, std::min(plan->n_tiles, device_num_sms()));
    int64_t rows = 0;
    for (int d = 0; d < 4; ++d) rows += (int64_t)(d + 2) * layer->L[d];
    return (int64_t)plan->n_tiles * stride + 16 + ((int64_t)plan->n_tiles * stride_am + 3) / 4 + 16 + (int64_t)grid * rows * EP + 16;
}

// returns 1 if launched (or would launch: do_launch false), 0 if not eligible, < 0 on error.  part_off / ncta / part_total describe
// the partial copies for k_param_finalize.
int launch_conv_bwd_wide(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx, const float* xnorm,
                         const void* ximg, const float* grad, int32_t ldg, int32_t grad_mode, const uint8_t* argmax,
                         const int64_t scoff[4], float* coef, float* partials, float* grad_x, int32_t ldgx, int64_t part_off[4],
                         int ncta[4], int64_t* part_total, bool do_launch, cudaStream_t st) {
    if (!ximg || !coef || !wide_bwd_ok(plan, layer)) return 0;
    static int s_budget = 0, s_sms = 0;
    if (!s_budget) {
        s_budget = device_max_smem_optin();
        s_sms = device_num_sms();
        MK_REQUIRE(s_budget > 0 && s_sms > 0, "conv_bwd_wide: no CUDA device");
    }
    WideBwdArgs a;
    memset(&a, 0, sizeof(a));
    if (!a.wb.build(layer->L)) return 0;
    a.F = layer->F; a.Fp = layer->Fp; a.Fk = wide_fk(layer->Fp);
    a.nh = (a.Fk + 255) / 256;
    if (a.nh > 2) return 0;
    a.Nh = a.nh == 1 ? a.Fk : ((a.Fk / 2 + 31) / 32 * 32);      // whole 32-column flush chunks per half
    a.stage_bytes = a.Fk * 64;
    a.FW = layer->Fp + EP;
    int cpb[4];
    wide_bwd_partition(plan, layer, s_sms, a.wb, cpb);
    int64_t po = 0, rows_all = 0;
    for (int d = 0; d < 4; ++d) {
        a.L[d] = layer->L[d]; a.packed[d] = layer->packed[d];
        a.part_off[d] = part_off[d] = po;
        ncta[d] = layer->L[d] > 0 ? cpb[d] : 0;
        po += (int64_t)ncta[d] * (d + 2) * layer->L[d] * a.FW;
        rows_all += (int64_t)(d + 2) * layer->L[d];
    }
    *part_total = po;
    int stride, stride_am;
    wide_coef_strides(plan, layer, &stride, &stride_am);
    const int cgrid = std::max(1, std::min(plan->n_tiles, s_sms));
    float* amax = partials + po + 2 * rows_all + 16;       // [cgrid <= 512] floats behind k_param_finalize's Q scratch
    MK_REQUIRE(cgrid <= 512, "conv_bwd_wide: %d CTAs (the per-CTA max |coef| array holds 512)", cgrid);
    float* coefT = coef;
    uint8_t* amT = reinterpret_cast<uint8_t*>(coef + (size_t)plan->n_tiles * stride + 16);
    float* bondP = coef + (size_t)plan->n_tiles * stride + 16 + ((size_t)plan->n_tiles * stride_am + 3) / 4 + 16;
    bondP = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(bondP) + 15) & ~(uintptr_t)15);
    if (launch_coef_wide(plan, layer, grad, ldg, grad_mode, argmax, scoff, coefT, amT, stride, stride_am, bondP, amax, cgrid, false, st) <= 0)
        return 0;
    // shared memory: two Wt buffers, two tile-metadata buffers, ring of K-step stages
    int64_t off = 0;
    auto take = [&](int64_t bytes) { const int64_t o = off; off += (bytes + 127) / 128 * 128; return (int)o; };
    a.sm_wt = take(2 * 2 * (int64_t)WB_WT_ONE);          // two coefficient-block buffers
    a.ub_a = a.ub_am = 0;
    a.unit_bytes = (int)((sizeof(TileMetaG) + 127) / 128 * 128);
    a.sm_unit = take(2 * (int64_t)a.unit_bytes);
    a.sm_ring = take(0);
    const int64_t room = (int64_t)s_budget - 2048 - off;
    a.nstages = (int)std::min<int64_t>(WB_MAXSTAGES, room / a.stage_bytes);
    if (a.nstages < 2) return 0;
    off += (int64_t)a.nstages * a.stage_bytes;
    if (!do_launch) return 1;
    MK_REQUIRE((reinterpret_cast<uintptr_t>(coef) & 15) == 0, "conv_bwd_wide: coef must be 16-byte aligned");
    // ---- coefficients (blocked tile order), max |coef|, bond sums ----
    if (launch_coef_wide(plan, layer, grad, ldg, grad_mode, argmax, scoff, coefT, amT, stride, stride_am, bondP, amax, cgrid, true, st) <= 0)
        return -1;
    static int64_t s_attr_dev[16] = {0};
    int64_t& s_attr = s_attr_dev[device_index()];
    if (off > s_attr) {
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_conv_bwd_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off));
        s_attr = off;
    }
    a.meta = reinterpret_cast<const TileMetaG*>(plan->tile_meta);
    a.tile_start = plan->tile_start;
    a.n_tiles = plan->n_tiles;
    a.coefT = coefT; a.amT = amT; a.stride = stride; a.stride_am = stride_am;
    a.amax = amax; a.namax = cgrid;
    a.partials = partials;
    const unsigned char* kimg = reinterpret_cast<const unsigned char*>(layer->tile_img);
    const unsigned char* xim = reinterpret_cast<const unsigned char*>(ximg);
    static int s_flx = -1, s_flg = -1;
    if (s_flx < 0) {
        const char* e = getenv("MOLKGNN_WIDE_FLUSH_X");
        s_flx = e ? std::max(1, atoi(e)) : 4;
        e = getenv("MOLKGNN_WIDE_FLUSH_G");
        s_flg = e ? std::max(1, atoi(e)) : 4;
    }
    // ---- phase G: kernel-parameter gradients ----
    {
        WideBwdArgs g = a;
        g.phase = 1;
        g.flush = s_flg;
        g.bimg = xim + wide_ximg_bwd_off(plan->n_tiles, a.Fk);
        int cb = 0;
        for (int b = 0; b < a.wb.nb; ++b) { g.cta_begin[b] = cb; cb += cpb[a.wb.d[b] - 1]; }
        for (int b = a.wb.nb; b <= WIDE_MAXB; ++b) g.cta_begin[b] = cb;
        count_launches(1);
        ProfScope prof("bwd_w", st);
        k_conv_bwd_wide<<<cb, WB_THREADS, off, st>>>(g);
        MK_CHECK_CUDA(cudaGetLastError());
        BondReduceArgs r;
        r.bondP = bondP; r.nc = cgrid;
        for (int d = 0; d < 4; ++d) { r.L[d] = layer->L[d]; r.part_off[d] = part_off[d]; }
        r.partials = partials; r.FW = a.FW; r.Fp = layer->Fp;
        count_launches(1);
        k_wide_bond_reduce<<<(int)((rows_all * EP + 255) / 256), 256, 0, st>>>(r);
        MK_CHECK_CUDA(cudaGetLastError());
    }
    // ---- phase X: input gradient ----
    if (grad_x) {
        MK_REQUIRE(ldgx >= layer->Fp && ldgx <= 512, "conv_bwd_wide: ldgx=%d", ldgx);
        WideBwdArgs xg = a;
        xg.phase = 0;
        xg.flush = s_flx;
        xg.bimg = kimg + wide_img_bwd_off(a.wb.nb, a.Fk, layer->L);
        xg.gx = grad_x; xg.ldgx = ldgx;
        const int grid = std::min(plan->n_tiles, s_sms);
        xg.order = plan->tile_grid == grid ? plan->tile_order : nullptr;
        xg.order_grid = xg.order ? grid : 0;
        count_launches(2);
        ProfScope prof("bwd_x", st);
        k_conv_bwd_wide<<<grid, WB_THREADS, off, st>>>(xg);
        MK_CHECK_CUDA(cudaGetLastError());
        JacArgs j;
        j.gx = grad_x; j.ldgx = ldgx; j.x = x; j.ldx = ldx; j.xnorm = xnorm; j.N = plan->N; j.F = layer->F; j.Fp = layer->Fp;
        k_wide_jacobian<<<(plan->N + 7) / 8, 256, 0, st>>>(j);
        MK_CHECK_CUDA(cudaGetLastError());
    }
    return 1;
}

}  // namespace mk

#ifdef MK_PHASE_CLOCKS
extern "C" int molkgnn_debug_phase_clocks_wbwd(unsigned long long* out96) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out96, mk::g_ph_wbwd, sizeof(unsigned long long) * 96) != cudaSuccess) return -1;
    unsigned long long z[96] = {0};
    return cudaMemcpyToSymbol(mk::g_ph_wbwd, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif
