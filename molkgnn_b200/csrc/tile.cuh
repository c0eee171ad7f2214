// Shared infrastructure of the molecule-tile kernels (conv_fwd_tile.cu, conv_bwd_tile.cu, bucket.cu, params.cu).
//
// A TILE is a run of <= 128 consecutive nodes that holds whole molecules (no edge crosses its boundary): one tcgen05 N
// block / 128 TMEM columns.  The kernel rows of a layer are packed into BLOCKS of <= 128 rows (one UMMA M block = the 128
// TMEM lanes); a block holds segments, a segment holds kernels k0..k0+nk-1 of one degree d in slot-major order:
//        row = rowbase + slot * nk + (k - k0),   slot 0..d-1 = support row s, slot d = centre row.
// The tile kernels run block-major: a persistent CTA keeps ONE block's fp16 images resident in shared memory and pulls
// tiles from a queue.
#pragma once
#include "common.cuh"
#include "tc.cuh"

namespace mk {

constexpr int TNODES = MOLKGNN_TILE_NODES;      // 128
constexpr int TILE_ESLOTS = 4 * TNODES;         // bond slots of a tile (node order)
constexpr int TILE_MAXB = 8;                    // blocks per layer
constexpr int TILE_MAXSEG = 4;                  // segments per block (one per degree)

struct TileSeg { int d, k0, nk, rowbase; };

struct TileBlocks {
    int nb;
    int nseg[TILE_MAXB];
    int rows[TILE_MAXB];
    TileSeg seg[TILE_MAXB][TILE_MAXSEG];
    // A degree's kernels are split evenly over the fewest blocks that hold them; a later (smaller) degree shares the
    // last block only if it fits there completely.  Base model 10/20/30/50: [d4 k0-24] [d4 k25-49] [d3] [d2, d1].
    __host__ __device__ bool build(const int* L) {
        nb = 0;
        for (int b = 0; b < TILE_MAXB; ++b) { nseg[b] = 0; rows[b] = 0; }
        for (int d = 4; d >= 1; --d) {
            const int Ld = L[d - 1];
            if (Ld <= 0) continue;
            const int per = 128 / (d + 1);                       // kernels per empty block
            if (nb > 0 && rows[nb - 1] + Ld * (d + 1) <= 128 && nseg[nb - 1] < TILE_MAXSEG) {
                TileSeg& sg = seg[nb - 1][nseg[nb - 1]++];
                sg.d = d; sg.k0 = 0; sg.nk = Ld; sg.rowbase = rows[nb - 1];
                rows[nb - 1] += Ld * (d + 1);
                continue;
            }
            const int need = (Ld + per - 1) / per;
            if (nb + need > TILE_MAXB) return false;
            int k0 = 0;
            for (int i = 0; i < need; ++i) {
                const int nk = (Ld - k0 + (need - i) - 1) / (need - i);
                TileSeg& sg = seg[nb][0];
                sg.d = d; sg.k0 = k0; sg.nk = nk; sg.rowbase = 0;
                nseg[nb] = 1; rows[nb] = nk * (d + 1);
                ++nb;
                k0 += nk;
            }
        }
        return nb > 0;
    }
};

__host__ __device__ inline int tile_fk(int Fp) { return (Fp + 15) / 16 * 16; }
// one fp16 image (hi or lo) of a 128-row block / of a node tile: [16 row groups][Fk/8 chunks][8 rows][8 elements]
__host__ __device__ inline int tile_img_one(int Fk) { return 16 * (Fk / 8) * 128; }
// K-step-major copy of the kernel-block images for the layer-fused forward (stack_fwd_fused.cu): one K step (16 features) of
// one block is 8 KB contiguous -- [hi: 16 row groups x 2 chunks x 128 B][lo: same] -- so that a CTA streams a block through
// a ring of small stages (K-major operand: LBO = 128, SBO = 256).  It lives behind the block-major images in layer->tile_img,
// followed by the bond-support table of the layer: for every block, every segment: [slot][half][kernel] float4.
constexpr int TILE_KS_BYTES = 8192;
__host__ __device__ inline int64_t tile_img_ks_off(int nb, int Fk) { return (int64_t)nb * 2 * tile_img_one(Fk); }
__host__ __device__ inline int64_t tile_es_off(int nb, int Fk) { return 2 * tile_img_ks_off(nb, Fk); }
__host__ __device__ inline int tile_es_f4(const int* L) { return 2 * (L[0] + 2 * L[1] + 3 * L[2] + 4 * L[3]); }
// ---- wide layers (conv_fwd_wide.cu, conv_bwd_wide.cu): more than TILE_MAXB blocks or more than 112 features -----------------
// Blocks of ONE degree each (rows = slot * nk + (k - k0), slot d = centre row), degree 4 first.  Both operands are streamed in
// STAGES of 16 KB = two K steps (32 features) of a 128-row operand: [K step][hi | lo][16 row groups][2 chunks][8 rows][8 elements]
// (K-major UMMA operand per K step: LBO = 128, SBO = 256).
constexpr int WIDE_MAXB = 16;
constexpr int WIDE_STAGE = 16384;
// Both operands of the wide kernels are multiplied by 2^4 before the (hi, lo) split: the entries of a normalised 440-feature row
// are ~0.05, whose fp16 remainder is a subnormal (absolute error 2^-25 = 2^-20.7 relative); scaled, the split keeps 4 more bits.
// The accumulators then carry 2^8 (forward: both operands scaled) or 2^4 (backward: one), removed exactly in the epilogues.
constexpr float WIDE_OPSCALE = 16.0f;
struct WideBlocks {
    int nb;
    int d[WIDE_MAXB], k0[WIDE_MAXB], nk[WIDE_MAXB];
    __host__ __device__ bool build(const int* L) {
        nb = 0;
        for (int b = 0; b < WIDE_MAXB; ++b) { d[b] = 1; k0[b] = 0; nk[b] = 0; }
        for (int dd = 4; dd >= 1; --dd) {
            const int Ld = L[dd - 1];
            if (Ld <= 0) continue;
            const int per = 128 / (dd + 1);
            const int need = (Ld + per - 1) / per;
            if (nb + need > WIDE_MAXB) return false;
            int k = 0;
            for (int i = 0; i < need; ++i) {
                const int n = (Ld - k + (need - i) - 1) / (need - i);
                d[nb] = dd; k0[nb] = k; nk[nb] = n;
                ++nb;
                k += n;
            }
        }
        return nb > 0;
    }
};
__host__ __device__ inline int wide_fk(int Fp) { return (Fp + 31) / 32 * 32; }
// byte offset of (row, 8-column chunk c = 0..3) inside a stage; the lo half lies WIDE_STAGE / 4 behind the hi half
__host__ __device__ inline uint32_t wide_stage_off(int row, int c) {
    return (uint32_t)(c >> 1) * (WIDE_STAGE / 2) + (uint32_t)(row >> 3) * 256u + (uint32_t)(c & 1) * 128u + (uint32_t)(row & 7) * 16u;
}
// kernel-block images [block][stage], then the bond-support table
__host__ __device__ inline int64_t wide_es_off(int nb, int Fk) { return (int64_t)nb * (Fk / 32) * WIDE_STAGE; }
// Backward images (conv_bwd_wide.cu) behind them: the same 128-row operands as MN-major UMMA operands in K-step-major order --
// stage ks = rows 16 ks .. 16 ks + 15: [hi | lo][2 row groups][Fk / 8 chunks][8 rows][8 elements] = Fk * 64 bytes, 8 stages per
// operand (LBO = Fk / 8 * 128: next row group, SBO = 128: next 8-column chunk).
__host__ __device__ inline int64_t wide_img_bwd_off(int nb, int Fk, const int* L) {
    return wide_es_off(nb, Fk) + ((int64_t)tile_es_f4(L) * 16 + 127) / 128 * 128;
}
__host__ __device__ inline int64_t wide_ximg_bwd_off(int n_tiles, int Fk) { return (int64_t)n_tiles * (Fk / 32) * WIDE_STAGE; }
// byte offset of (row, 8-column chunk c8) inside a 128-row backward operand; the lo half lies Fk * 32 behind the hi half
__host__ __device__ inline uint32_t wide_bstage_off(int row, int c8, int Fk) {
    return (uint32_t)(row >> 4) * (uint32_t)(Fk * 64) + (uint32_t)((row >> 3) & 1) * (uint32_t)(Fk >> 3) * 128u + (uint32_t)c8 * 128u +
           (uint32_t)(row & 7) * 16u;
}

// Images in global memory: kernel blocks [block][hi | lo]; node tiles [tile][hi | lo].  v = hi + lo, BOTH halves unscaled
// (lo is usually an fp16 subnormal, which tcgen05 honours): one fp32 accumulator receives hi*hi + lo*hi + hi*lo.

// Per-tile metadata, built once per batch by k_tile_meta (bucket.cu) and fetched with one bulk copy per tile visit.
struct __align__(16) TileMetaG {
    int t0, nn;                       // first node, number of nodes
    int e0, ne;                       // first bond slot in ehat_node, number of slots
    int cnt[4];                       // nodes per degree
    uint32_t nl[TNODES];              // 4 local neighbour ids, 8 bits each (neighbour order = edge order)
    int posl[TNODES];                 // bucket row R of the node
    unsigned short eslot[TNODES];     // local slot of the node's first bond
    unsigned char degl[TNODES];
    unsigned char list[4][TNODES];    // local ids of the degree-d nodes, ascending
    signed char tsg[TNODES];          // degree 4: sign of the neighbour triple product (kernels.py:336)
    uint32_t inl[TNODES];             // 4 local in-neighbour ids (sources of the in-edges, edge order), 8 bits each
    unsigned char inj[TNODES][4];     // position of this node in that source's neighbour list
    unsigned char incnt[TNODES];
    unsigned char lidx[TNODES];       // index of the node inside list[deg - 1]
    unsigned char cr[TNODES];         // collision rank of the node's neighbour slots, 2 bits per j (see elist)
    // Collision chains of the backward scatter.  The neighbour slots (node n, j) of degree-d nodes that share a TARGET
    // nei(n, j) = v write the same column of the coefficient block; their rank (cr) is the number of earlier in-edges of v whose
    // source also has degree d.  Rank 0 is a plain store by the slot's own (node, kernel) thread; the FOLLOWERS (ranks 1..3) of
    // one (target, degree) form a chain that ONE thread per kernel adds in in-edge order after a barrier -- deterministic, no
    // atomics, one barrier.  chain word: follower f (0..2) in bits [9 f, 9 f + 9) as (node << 2 | j), count - 1 in bits [27, 29).
    uint32_t chains[TILE_ESLOTS / 2];
    int choff[5];                     // chains of degree d: [choff[d-1], choff[d])
    int pad_[15];
};
static_assert(sizeof(TileMetaG) % 16 == 0, "TileMetaG must be a multiple of 16 bytes");

// ---- balanced tile schedule of the tile-major kernels (k_coef_tile, k_conv_bwd_tile) ------------------------------------
// With n tiles on G persistent CTAs (n = q G + r) CTAs 0..r-1 walk q + 1 tiles and the others q.  Without an order array
// CTA b walks tiles b, b + G, ... (round robin: per-CTA node totals differ by ~6 %, and the r CTAs with one tile more decide
// the kernel time).  With plan->tile_order (k_tile_order, bucket.cu: tiles sorted by node count, the r long CTAs take the
// smallest (q + 1) r tiles, the others the rest, boustrophedon inside each group) CTA b walks order[beg .. beg + cnt).
struct TileWalk {
    int beg, cnt, b, G;
    const int* order;
    __device__ __forceinline__ TileWalk(const int* order_, int order_grid, int n_tiles) {
        b = (int)blockIdx.x; G = (int)gridDim.x;
        const int q = n_tiles / G, r = n_tiles - q * G;
        cnt = q + (b < r ? 1 : 0);
        beg = b < r ? b * (q + 1) : r * (q + 1) + (b - r) * q;
        order = (order_ && order_grid == G) ? order_ : nullptr;
    }
    __device__ __forceinline__ int tile(int k) const { return order ? __ldg(order + beg + k) : b + k * G; }
};

// ---- bulk async copy global -> shared with mbarrier completion ---------------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
// size % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc::smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}

}  // namespace mk
