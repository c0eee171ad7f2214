// Degree bucketing of a collated PyG batch on the GPU: one CSR-to-degree-bucket pass.
// Replaces ToXAndPAndEdgeAttrForDeg (reference wrapper.py:559-672) + PyG collation of its 20 attributes.
//
// Integer work, HBM/latency bound.  Four small launches:
//   k_edge_slots   every edge claims a slot in its source's out-list and its target's in-list (<=4 each)
//   k_node_prepare per node: sort the <=4 claimed edge ids (restores edge order, wrapper.py:567-572), degree class,
//                  per-block class histogram
//   k_scan_blocks  exclusive scan of the block histograms -> stable bucket positions (ascending node id,
//                  wrapper.py:599-600) and the bucket sizes
//   k_assign       writes sel / pos / nei / nei_eid / ehat / tsign / in-lists
#include <cstring>
#include "common.cuh"
#include "tile.cuh"

namespace mk {

constexpr int BT = 256;  // threads per block in the node kernels

__global__ void k_edge_slots(const int64_t* __restrict__ ei, int E, int N, int* out_cnt, int* out_eid, int* in_cnt,
                             int* in_eid, int* far, int* err) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    long long u = ei[e], v = ei[(size_t)E + e];
    if (u < 0 || u >= N || v < 0 || v >= N) { atomicOr(err, 1); return; }
    // furthest node reached from the lower end of the edge: a tile boundary may not fall inside (lo, hi]
    {
        const int lo = (int)(u < v ? u : v), hi = (int)(u < v ? v : u);
        if (hi > lo) atomicMax(&far[lo], hi);
    }
    int so = atomicAdd(&out_cnt[u], 1);
    if (so < 4) out_eid[4 * u + so] = e; else atomicOr(err, 2);
    int si = atomicAdd(&in_cnt[v], 1);
    if (si < 4) in_eid[4 * v + si] = e; else atomicOr(err, 4);
}

__device__ __forceinline__ void sort4(int* a, int n) {
    // tiny insertion sort, n <= 4
    for (int i = 1; i < n; ++i) {
        int key = a[i], j = i - 1;
        while (j >= 0 && a[j] > key) { a[j + 1] = a[j]; --j; }
        a[j + 1] = key;
    }
}

__global__ void k_node_prepare(int N, const int* __restrict__ out_cnt, int* out_eid, const int* __restrict__ in_cnt,
                               int* in_eid, int* deg, int* blk_counts, int* err) {
    int v = blockIdx.x * BT + threadIdx.x;
    int cls = -1;
    if (v < N) {
        int d = out_cnt[v];
        deg[v] = d;
        if (d >= 1 && d <= 4) {
            cls = d - 1;
            int a[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) a[j] = j < d ? out_eid[4 * v + j] : 0x7fffffff;
            sort4(a, d);
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j < d) out_eid[4 * v + j] = a[j];
        } else {
            cls = 4;
            atomicOr(err, 8);
        }
        int di = min(in_cnt[v], 4);
        int b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = j < di ? in_eid[4 * v + j] : 0x7fffffff;
        sort4(b, di);
#pragma unroll
        for (int j = 0; j < 4; ++j) if (j < di) in_eid[4 * v + j] = b[j];
    }
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        int cnt = __syncthreads_count(cls == c);
        if (threadIdx.x == 0) blk_counts[blockIdx.x * 5 + c] = cnt;
    }
}

// totals[0..3] = n_d, totals[4] = bad nodes, totals[5..8] = boff, totals[9..12] = eoff.  One warp per class scans the
// block histograms 32 blocks at a time.
__global__ void __launch_bounds__(160) k_scan_blocks(int nblk, const int* __restrict__ blk_counts, int* blk_off, int* totals) {
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int run = 0;
    for (int b0 = 0; b0 < nblk; b0 += 32) {
        const int b = b0 + lane;
        const int v = b < nblk ? blk_counts[b * 5 + c] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (b < nblk) blk_off[b * 5 + c] = run + incl - v;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) totals[c] = run;
    __syncthreads();
    if (threadIdx.x == 0) {
        int bo = 0, eo = 0;
        for (int d = 0; d < 4; ++d) {
            totals[5 + d] = bo;
            totals[9 + d] = eo;
            bo += totals[d];
            eo += totals[d] * (d + 1);
        }
    }
}

// Optional source of the raw per-degree data tensors a reference batch carries (nei_edge_attr_deg*, p_focal_deg4, nei_p_deg4,
// kernels.py:628-645): the reference convolves THOSE (kernels.py:679), never the edge_attr / p handed to MolGCN.forward --
// MolKGNNNet passes a batch-normalised edge_attr there (MolKGNNNet.py:116-119).  Bucket row r, slot j = row r * d + j.
struct RefRows { const float* nea[4]; const float* pf4; const float* np4; };

__global__ void k_assign(int N, int E, const int64_t* __restrict__ ei, const int* __restrict__ deg,
                         const int* __restrict__ out_eid, const int* __restrict__ in_cnt,
                         const int* __restrict__ in_eid, const int* __restrict__ blk_off,
                         const int* __restrict__ totals, const float* __restrict__ p, int p_dim,
                         const float* __restrict__ edge_attr, int Fe, int* pos, int* sel, int* nei, int* nei_eid,
                         float* ehat, int8_t* tsign, int* in_src, int* in_j, const RefRows ref) {
    __shared__ int warp_cnt[BT / 32][4];
    int v = blockIdx.x * BT + threadIdx.x;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int d = v < N ? deg[v] : 0;
    int cls = (d >= 1 && d <= 4) ? d - 1 : -1;
    int rank = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        unsigned m = __ballot_sync(0xffffffffu, cls == c);
        if (cls == c) rank = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) warp_cnt[wid][c] = __popc(m);
    }
    __syncthreads();
    if (cls < 0) return;
    int r = blk_off[blockIdx.x * 5 + cls] + rank;
    for (int w = 0; w < wid; ++w) r += warp_cnt[w][cls];
    const int boff = totals[5 + cls], eoff = totals[9 + cls];
    pos[v] = r;
    sel[boff + r] = v;
    int nb[4];
    for (int j = 0; j < d; ++j) {
        int e = out_eid[4 * v + j];
        int u = (int)ei[(size_t)E + e];
        nb[j] = u;
        size_t row = (size_t)eoff + (size_t)r * d + j;
        nei[row] = u;
        nei_eid[row] = e;
        // bond attributes of the undirected bond: row 2*(eid/2) (wrapper.py:586-591), normalised for the cosine
        const float* ea = ref.nea[cls] ? ref.nea[cls] + ((size_t)r * d + j) * Fe : edge_attr + (size_t)(2 * (e / 2)) * Fe;
        float t[EP];
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < EP; ++c) { t[c] = c < Fe ? ea[c] : 0.f; ss += t[c] * t[c]; }
        float nrm = fmaxf(sqrtf(ss), MOLKGNN_COS_EPS);
#pragma unroll
        for (int c = 0; c < EP; ++c) ehat[row * EP + c] = t[c] / nrm;
    }
    if (d == 4 && p_dim == 3) {
        // sign of p2 . (p0 x p1) on neighbour coordinates calibrated by the focal atom (kernels.py:336,356)
        float q[3][3];
        for (int j = 0; j < 3; ++j)
            for (int c = 0; c < 3; ++c)
                q[j][c] = (ref.pf4 && ref.np4) ? __fsub_rn(ref.np4[((size_t)r * 4 + j) * 3 + c], ref.pf4[(size_t)r * 3 + c])
                                               : __fsub_rn(p[(size_t)nb[j] * 3 + c], p[(size_t)v * 3 + c]);
        float cx = __fsub_rn(__fmul_rn(q[0][1], q[1][2]), __fmul_rn(q[0][2], q[1][1]));
        float cy = __fsub_rn(__fmul_rn(q[0][2], q[1][0]), __fmul_rn(q[0][0], q[1][2]));
        float cz = __fsub_rn(__fmul_rn(q[0][0], q[1][1]), __fmul_rn(q[0][1], q[1][0]));
        float dt = __fadd_rn(__fadd_rn(__fmul_rn(q[2][0], cx), __fmul_rn(q[2][1], cy)), __fmul_rn(q[2][2], cz));
        tsign[r] = dt > 0.f ? 1 : (dt < 0.f ? -1 : 0);
    }
    int di = min(in_cnt[v], 4);
    for (int t = 0; t < 4; ++t) {
        int u = -1, jj = 0;
        if (t < di) {
            int e = in_eid[4 * v + t];
            u = (int)ei[e];
            const int du = min(deg[u], 4);   // slots >= deg[u] are uninitialised scratch
            for (int k = 0; k < du; ++k) if (out_eid[4 * u + k] == e) jj = k;
        }
        in_src[4 * v + t] = u;
        in_j[4 * v + t] = jj;
    }
}

// ---- molecule tiles ------------------------------------------------------------------------------------------------
// tinfo: [0] widest gap between consecutive valid cuts (= largest molecule), [1] flags (1 = an edge spans >= 128 nodes),
//        [2] n_tiles, [3] largest tile, [4..7] largest number of degree-1..4 nodes in one tile
constexpr int TILE_CAP = MOLKGNN_TILE_NODES;
constexpr int TILE_MIN_STRIDE = 32;

// cutpos[c] = c if no edge crosses the boundary in front of node c (c = 0..N), else -1
__global__ void __launch_bounds__(256) k_valid_cuts(int N, const int* __restrict__ far, int* cutpos, int* tinfo) {
    __shared__ int f[256 + TILE_CAP];
    const int B = blockIdx.x * 256;
    for (int i = threadIdx.x; i < 256 + TILE_CAP; i += 256) {
        const int u = B - TILE_CAP + i;
        int r = -1;
        if (u >= 0 && u < N) {
            r = max(far[u], u);
            if (r - u >= TILE_CAP) atomicOr(&tinfo[1], 1);
        }
        f[i] = r;
    }
    __syncthreads();
    const int c = B + threadIdx.x;
    if (c > N) return;
    int m = -1;   // edges that could cross start at one of the TILE_CAP - 1 nodes in front of c (longer spans are flagged)
#pragma unroll 8
    for (int i = 0; i < TILE_CAP; ++i) m = max(m, f[threadIdx.x + i]);
    cutpos[c] = (m < c) ? c : -1;
}

__global__ void __launch_bounds__(256) k_cut_gaps(int N, const int* __restrict__ cutpos, int* tinfo) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    int gap = 0;
    if (c > 0 && c <= N && cutpos[c] >= 0) {
        int q = c - 1;
        while (q > 0 && cutpos[q] < 0 && c - q <= 2 * TILE_CAP) --q;   // bounded: anything wider disables tiling anyway
        gap = c - q;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gap = max(gap, __shfl_xor_sync(0xffffffffu, gap, o));
    if ((threadIdx.x & 31) == 0 && gap > 0) atomicMax(&tinfo[0], gap);
}

// tile i starts at the last valid cut <= i * S with S = TILE_CAP + 1 - (largest molecule): every tile then holds whole
// molecules and at most TILE_CAP nodes.  One warp per tile.
__global__ void __launch_bounds__(128) k_tile_starts(int N, int cap, const int* __restrict__ cutpos,
                                                     const int* __restrict__ deg, int* tile_start, int* tinfo) {
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int gap = tinfo[0];
    const int S = TILE_CAP + 1 - gap;
    if (S < TILE_MIN_STRIDE || (tinfo[1] & 1)) {
        if (i == 0 && lane == 0) tinfo[2] = 0;
        return;
    }
    const int T = (N + S - 1) / S;
    if (i == 0 && lane == 0) tinfo[2] = T;
    if (i > T || i >= cap) return;
    int a = 0, b = 0;
    if (lane == 0) {
        a = (int)min((long long)i * S, (long long)N);
        while (a > 0 && cutpos[a] < 0) --a;
        tile_start[i] = a;
    } else if (lane == 1 && i < T) {
        b = (int)min((long long)(i + 1) * S, (long long)N);
        while (b > 0 && cutpos[b] < 0) --b;
    }
    a = __shfl_sync(0xffffffffu, a, 0);
    b = __shfl_sync(0xffffffffu, b, 1);
    if (i == T) return;
    int cnt[4] = {0, 0, 0, 0};
    for (int v = a + lane; v < b; v += 32) {
        const int d = deg[v];
#pragma unroll
        for (int c = 0; c < 4; ++c) cnt[c] += (d == c + 1);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt[c] += __shfl_xor_sync(0xffffffffu, cnt[c], o);
    if (lane == 0) {
        atomicMax(&tinfo[3], b - a);
#pragma unroll
        for (int d = 0; d < 4; ++d) atomicMax(&tinfo[4 + d], cnt[d]);
    }
}

// Greedy tiling: every tile starts where the previous one ended and takes as many whole molecules as fit TILE_CAP nodes
// (the stride scheme above leaves a quarter of every tile empty).  The walk is sequential, so ONE CTA does it on a bitmap
// of the valid cuts held in shared memory (N + 1 bits): warp 0 hops tile by tile (the last set bit of a 128-bit window),
// then all warps gather the per-tile maxima.  Used whenever the bitmap fits shared memory.
// one greedy hop on the cut bitmap: the last valid cut in (c, min(c + TILE_CAP, N)], or -1
__device__ __forceinline__ int greedy_hop(const uint32_t* bits, int c, int N) {
    const int hi = min(c + TILE_CAP, N);
    // Fast path: a 64-bit window ending at hi (bit 63 <-> position hi) holds at least the 32 positions below hi, enough
    // whenever the last molecule before hi has <= 32 atoms; both words load independently.
    const int w1 = hi >> 5, sh = 31 - (hi & 31);
    const unsigned long long win = (((unsigned long long)bits[w1] << 32) | (w1 > 0 ? bits[w1 - 1] : 0u)) << sh;
    if (win != 0ull) {
        const int b = hi - __clzll(win);
        if (b > c) return b;
    }
    int w = w1;                                            // long molecule: scan the words downwards from hi's word
    uint32_t word = bits[w];
    if ((hi & 31) != 31) word &= (2u << (hi & 31)) - 1u;
    const int wlo = (c + 1) >> 5;
    while (true) {
        if (w == wlo) word &= ~((1u << ((c + 1) & 31)) - 1u);
        if (word) return 32 * w + 31 - __clz(word);
        if (w == wlo) return -1;
        word = bits[--w];
    }
}

// The walk is a chain of dependent hops (~160 cycles each), so it is cut into up to 32 node segments walked by one warp
// each.  The chain enters segment p at its first tile start >= p*S, one of the valid cuts in [p*S, p*S + TILE_CAP): every
// lane of warp p walks from one of those candidates to the segment's end and notes where it leaves and how many tiles it
// made; one thread then threads the true entry through the segments (32 table look-ups), and the warps re-walk from their
// true entries writing tile_start at their prefix offsets.  Exact: the result is the sequential greedy tiling.
constexpr int GW = 32;                                     // segments = warps
__global__ void __launch_bounds__(1024) k_tile_starts_greedy(int N, int cap, const int* __restrict__ cutpos,
                                                             int* tile_start, int* tinfo) {
    extern __shared__ uint32_t bits[];
    __shared__ int s_cand[GW][32], s_exit[GW][32], s_cnt[GW][32], s_ncand[GW];
    __shared__ int s_entry[GW + 1], s_off[GW + 1], s_fail;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwords = (N + 32) / 32;
    constexpr int U = 8;                                  // words per warp in flight
    for (int w0 = warp * U; w0 < nwords; w0 += 32 * U) {
        int v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = 32 * (w0 + u) + lane;
            v[u] = (w0 + u < nwords && c <= N) ? cutpos[c] : -1;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t word = __ballot_sync(0xffffffffu, v[u] >= 0);
            if (lane == 0 && w0 + u < nwords) bits[w0 + u] = word;
        }
    }
    if (tid == 0) s_fail = 0;
    __syncthreads();
    const int gap = tinfo[0];
    // the greedy walk only needs every molecule to fit a tile (any two consecutive greedy tiles hold > TILE_CAP nodes
    // together, so there are at most N / 64 + 1 of them whatever the molecule sizes -- inside the N / 32 + 4 capacity)
    if (gap > TILE_CAP || (tinfo[1] & 1)) { if (tid == 0) tinfo[2] = 0; return; }
    const int P = max(1, min(GW, N / 2048));              // segments of >= 2048 nodes
    const int S = (N + P - 1) / P;
    const int seg_lo = warp * S, seg_hi = min(N, (warp + 1) * S);
    // ---- candidates of this segment and their (exit, tiles) ----
    if (warp < P) {
        int cand = -1;
        if (warp == 0) {
            if (lane == 0) cand = 0;
        } else {
            // the valid cuts in [seg_lo, seg_lo + TILE_CAP), in order, one per lane
            int found = 0;
            for (int q0 = 0; q0 < TILE_CAP; q0 += 32) {
                const int q = seg_lo + q0 + lane;
                const bool ok = q <= N && ((bits[q >> 5] >> (q & 31)) & 1u);
                const uint32_t m = __ballot_sync(0xffffffffu, ok);
                const int rank = found + __popc(m & ((1u << lane) - 1u));
                if (ok && rank < 32) s_cand[warp][rank] = q;
                found += __popc(m);
            }
            __syncwarp();
            if (lane < min(found, 32)) cand = s_cand[warp][lane];
            __syncwarp();
            if (found > 32 && lane == 0) s_fail = 1;       // more candidates than lanes (tiny molecules): sequential fallback
            if (lane == 0) s_ncand[warp] = min(found, 32);
        }
        if (warp == 0 && lane == 0) s_ncand[0] = 1;
        int c = cand, n = 0;
        if (c >= 0) {
            while (c < seg_hi) {
                const int b = greedy_hop(bits, c, N);
                if (b < 0) { s_fail = 1; break; }
                c = b;
                ++n;
            }
        }
        s_cand[warp][lane] = cand; s_exit[warp][lane] = c; s_cnt[warp][lane] = n;
    }
    __syncthreads();
    // ---- thread the true entry through the segments ----
    if (tid == 0) {
        int e = 0, off = 0;
        bool ok = !s_fail;
        for (int p = 0; p < P && ok; ++p) {
            s_entry[p] = e; s_off[p] = off;
            if (e >= min(N, (p + 1) * S)) continue;        // the chain jumps over this segment's start window entirely
            int l = -1;
            for (int i = 0; i < s_ncand[p]; ++i) if (s_cand[p][i] == e) l = i;
            if (l < 0) { ok = false; break; }
            off += s_cnt[p][l];
            e = s_exit[p][l];
        }
        if (ok && (e != N || off > cap - 1)) ok = false;
        s_entry[P] = N; s_off[P] = off;
        if (!ok) s_fail = 1;
    }
    __syncthreads();
    if (s_fail) {                                          // exact sequential walk
        if (tid != 0) return;
        int c = 0, t = 0;
        while (c < N && t < cap - 1) {
            tile_start[t] = c;
            const int b = greedy_hop(bits, c, N);
            if (b <= c) { t = 0; c = N; break; }           // cannot happen when gap <= TILE_CAP; disables tiling if it does
            c = b;
            ++t;
        }
        if (c < N) t = 0;                                  // ran out of tile slots
        if (t > 0) tile_start[t] = N;
        tinfo[2] = t;
        return;
    }
    // ---- write the tile starts: warp p re-walks its part of the chain ----
    if (warp < P && lane == 0) {
        int c = s_entry[warp], t = s_off[warp];
        while (c < seg_hi) {
            tile_start[t++] = c;
            c = greedy_hop(bits, c, N);
        }
    }
    if (tid == 0) { tile_start[s_off[P]] = N; tinfo[2] = s_off[P]; }
}

// largest tile and largest per-degree node count of a tile (one warp per tile)
__global__ void __launch_bounds__(128) k_tile_stats(int cap, const int* __restrict__ tile_start, const int* __restrict__ deg,
                                                    int* tinfo) {
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int T = tinfo[2];
    if (i >= T || i >= cap) return;
    const int a = tile_start[i], b = tile_start[i + 1];
    int cnt[4] = {0, 0, 0, 0};
    for (int v = a + lane; v < b; v += 32) {
        const int d = deg[v];
#pragma unroll
        for (int c = 0; c < 4; ++c) cnt[c] += (d == c + 1);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt[c] += __shfl_xor_sync(0xffffffffu, cnt[c], o);
    if (lane == 0) {
        atomicMax(&tinfo[3], b - a);
#pragma unroll
        for (int d = 0; d < 4; ++d) atomicMax(&tinfo[4 + d], cnt[d]);
    }
}

// per-tile metadata record + bond rows in node order (one block of 128 threads per tile, thread = local node)
__global__ void __launch_bounds__(TNODES) k_tile_meta(int N, const int* __restrict__ tile_start, const int* __restrict__ deg,
                                                      const int* __restrict__ pos, const int* __restrict__ nei,
                                                      const float* __restrict__ ehat, const int8_t* __restrict__ tsign,
                                                      const int* __restrict__ in_cnt, const int* __restrict__ in_src,
                                                      const int* __restrict__ in_j, const int* __restrict__ blk_off,
                                                      const int* __restrict__ totals, TileMetaG* meta, float* ehat_node,
                                                      int* node_tile) {
    __shared__ int wsum[4][6];
    __shared__ int s_e0;
    __shared__ int s_ccnt[4], s_coff[5], s_cfill[4];
    const int tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t0 = tile_start[tile], t1 = tile_start[tile + 1];
    const int nn = t1 - t0;
    TileMetaG* m = meta + tile;
    if (tid < 4) { s_ccnt[tid] = 0; s_cfill[tid] = 0; }
    // first bond slot of the tile = sum of the degrees of all nodes in front of it: whole 256-node blocks from the
    // class histogram of the bucket pass, the rest by a block reduction
    const int b0 = t0 / BT;
    int part = 0;
    for (int v = b0 * BT + tid; v < t0; v += TNODES) part += deg[v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) wsum[warp][5] = part;
    int d = 0, R = 0, base = 0;
    if (tid < nn) {
        d = deg[t0 + tid];
        R = pos[t0 + tid];
        base = totals[9 + d - 1] + R * d;
    }
    int rank[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const unsigned mk_ = __ballot_sync(0xffffffffu, d == c + 1);
        rank[c] = __popc(mk_ & ((1u << lane) - 1u));
        if (lane == 0) wsum[warp][c] = __popc(mk_);
    }
    int incl = d;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp][4] = incl;
    __syncthreads();
    if (tid == 0) {
        int e0 = wsum[0][5] + wsum[1][5] + wsum[2][5] + wsum[3][5];
        for (int c = 0; c < 4; ++c) e0 += (c + 1) * blk_off[b0 * 5 + c];
        s_e0 = e0;
        m->t0 = t0; m->nn = nn; m->e0 = e0;
        m->ne = wsum[0][4] + wsum[1][4] + wsum[2][4] + wsum[3][4];
        for (int c = 0; c < 4; ++c) m->cnt[c] = wsum[0][c] + wsum[1][c] + wsum[2][c] + wsum[3][c];
    }
    __syncthreads();
    int soff = incl - d;
    int loff = d > 0 ? rank[d - 1] : 0;
    for (int w = 0; w < warp; ++w) {
        soff += wsum[w][4];
        if (d > 0) loff += wsum[w][d - 1];
    }
    uint32_t w_nl = 0, w_in = 0;
    unsigned char ij[4] = {0, 0, 0, 0};
    int ic = 0;
    int crank[4] = {0, 0, 0, 0};
    if (tid < nn) {
        m->list[d - 1][loff] = (unsigned char)tid;
        for (int j = 0; j < d; ++j) {
            // rank of the edge (tid, j) among the in-edges of its target whose source has the same degree
            const int v = nei[(size_t)base + j];
            const int icv = min(in_cnt[v], 4);
            int r = 0;
            for (int t = 0; t < icv; ++t) {
                const int u = in_src[4 * (size_t)v + t];
                if (u == t0 + tid && in_j[4 * (size_t)v + t] == j) break;
                if (deg[u] == d) ++r;
            }
            crank[j] = min(r, 3);
        }
        for (int j = 0; j < d; ++j) {
            w_nl |= (uint32_t)((nei[(size_t)base + j] - t0) & 0xff) << (8 * j);
            const float4* src = reinterpret_cast<const float4*>(ehat + ((size_t)base + j) * EP);
            float4* dst = reinterpret_cast<float4*>(ehat_node + ((size_t)s_e0 + soff + j) * EP);
            dst[0] = src[0];
            dst[1] = src[1];
        }
        ic = min(in_cnt[t0 + tid], 4);
        for (int t = 0; t < ic; ++t) {
            w_in |= (uint32_t)((in_src[4 * (size_t)(t0 + tid) + t] - t0) & 0xff) << (8 * t);
            ij[t] = (unsigned char)in_j[4 * (size_t)(t0 + tid) + t];
        }
    }
    if (tid < nn && node_tile) node_tile[t0 + tid] = (tile << 8) | tid;
    m->nl[tid] = w_nl;
    m->posl[tid] = R;
    m->eslot[tid] = (unsigned short)soff;
    m->degl[tid] = (unsigned char)d;
    m->tsg[tid] = (d == 4) ? tsign[R] : 0;
    m->inl[tid] = w_in;
#pragma unroll
    for (int t = 0; t < 4; ++t) m->inj[tid][t] = ij[t];
    m->incnt[tid] = (unsigned char)ic;
    m->lidx[tid] = (unsigned char)loff;
    m->cr[tid] = (unsigned char)(crank[0] | (crank[1] << 2) | (crank[2] << 4) | (crank[3] << 6));
    // collision chains (tile.cuh): this thread is the TARGET v; per source degree dd the in-edges of v whose source has that
    // degree, in in-edge order -- the first is the rank-0 store of its own pair thread, the others are the chain's followers
    uint32_t chain[4] = {0u, 0u, 0u, 0u};
    if (tid < nn) {
        int seen[4] = {0, 0, 0, 0};
        for (int t = 0; t < ic; ++t) {
            const int ul = (int)((w_in >> (8 * t)) & 0xffu);
            const int du = deg[t0 + ul];
            if (du < 1 || du > 4) continue;
            const int k = seen[du - 1]++;
            if (k >= 1) chain[du - 1] |= (uint32_t)((ul << 2) | ij[t]) << (9 * (k - 1));
        }
        for (int dd = 0; dd < 4; ++dd)
            if (seen[dd] >= 2) { chain[dd] |= (uint32_t)(seen[dd] - 2) << 27; atomicAdd(&s_ccnt[dd], 1); }
            else chain[dd] = 0xffffffffu;
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int dd = 0; dd < 4; ++dd) { s_coff[dd] = run; m->choff[dd] = run; run += s_ccnt[dd]; }
        s_coff[4] = run; m->choff[4] = run;
    }
    __syncthreads();
    if (tid < nn) {
        for (int dd = 0; dd < 4; ++dd)
            if (chain[dd] != 0xffffffffu) m->chains[s_coff[dd] + atomicAdd(&s_cfill[dd], 1)] = chain[dd];
    }
}

__global__ void k_export(int d, int n, int boff, int eoff, const int* __restrict__ sel, const int* __restrict__ nei,
                         const int* __restrict__ nei_eid, const float* __restrict__ p, int p_dim,
                         const float* __restrict__ edge_attr, int Fe, int64_t* selected_index, int64_t* nei_index,
                         float* p_focal, float* nei_p, float* nei_edge_attr) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int v = sel[boff + r];
    if (selected_index) selected_index[r] = v;
    if (p_focal) for (int c = 0; c < p_dim; ++c) p_focal[(size_t)r * p_dim + c] = p[(size_t)v * p_dim + c];
    for (int j = 0; j < d; ++j) {
        size_t row = (size_t)eoff + (size_t)r * d + j;
        int u = nei[row], e = nei_eid[row];
        size_t o = (size_t)r * d + j;
        if (nei_index) nei_index[o] = u;
        if (nei_p) for (int c = 0; c < p_dim; ++c) nei_p[o * p_dim + c] = p[(size_t)u * p_dim + c];
        if (nei_edge_attr)
            for (int c = 0; c < Fe; ++c) nei_edge_attr[o * Fe + c] = edge_attr[(size_t)(2 * (e / 2)) * Fe + c];
    }
}

// ---- plan from reference-format bucket tensors -------------------------------------------------------------
__global__ void k_from_buckets(int d, int n, int boff, int eoff, const int64_t* __restrict__ selected_index,
                               const int64_t* __restrict__ nei_index, const float* __restrict__ p_focal,
                               const float* __restrict__ nei_p, int p_dim, const float* __restrict__ nei_ea, int Fe,
                               int N, int* deg, int* pos, int* sel, int* nei, int* nei_eid, float* ehat, int8_t* tsign,
                               int* in_cnt, int* in_key, int* err) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    long long v = selected_index[r];
    if (v < 0 || v >= N) { atomicOr(err, 1); return; }
    deg[v] = d;
    pos[v] = r;
    sel[boff + r] = (int)v;
    for (int j = 0; j < d; ++j) {
        size_t o = (size_t)r * d + j, row = (size_t)eoff + o;
        long long u = nei_index[o];
        if (u < 0 || u >= N) { atomicOr(err, 1); u = 0; }
        nei[row] = (int)u;
        nei_eid[row] = -1;
        float t[EP];
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < EP; ++c) { t[c] = c < Fe ? nei_ea[o * Fe + c] : 0.f; ss += t[c] * t[c]; }
        float nrm = fmaxf(sqrtf(ss), MOLKGNN_COS_EPS);
#pragma unroll
        for (int c = 0; c < EP; ++c) ehat[row * EP + c] = t[c] / nrm;
        int si = atomicAdd(&in_cnt[u], 1);
        if (si < 4) in_key[4 * u + si] = (int)(row);  // global neighbour-row id identifies (source bucket row, j)
        else atomicOr(err, 4);
    }
    if (d == 4 && p_dim == 3 && tsign) {
        float q[3][3];
        for (int j = 0; j < 3; ++j)
            for (int c = 0; c < 3; ++c)
                q[j][c] = __fsub_rn(nei_p[((size_t)r * 4 + j) * 3 + c], p_focal[(size_t)r * 3 + c]);
        float cx = __fsub_rn(__fmul_rn(q[0][1], q[1][2]), __fmul_rn(q[0][2], q[1][1]));
        float cy = __fsub_rn(__fmul_rn(q[0][2], q[1][0]), __fmul_rn(q[0][0], q[1][2]));
        float cz = __fsub_rn(__fmul_rn(q[0][0], q[1][1]), __fmul_rn(q[0][1], q[1][0]));
        float dt = __fadd_rn(__fadd_rn(__fmul_rn(q[2][0], cx), __fmul_rn(q[2][1], cy)), __fmul_rn(q[2][2], cz));
        tsign[r] = dt > 0.f ? 1 : (dt < 0.f ? -1 : 0);
    }
}

// in_key holds neighbour-row ids; sort them (deterministic order) and decode into (source node, j)
__global__ void k_from_buckets_inlists(int N, const int* __restrict__ in_cnt, int* in_key, const int* __restrict__ sel,
                                       int4 boff, int4 eoff, int4 nd, int* in_src, int* in_j) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    int di = min(in_cnt[v], 4);
    int a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = j < di ? in_key[4 * v + j] : 0x7fffffff;
    sort4(a, di);
    const int bo[4] = {boff.x, boff.y, boff.z, boff.w};
    const int eo[4] = {eoff.x, eoff.y, eoff.z, eoff.w};
    const int nn[4] = {nd.x, nd.y, nd.z, nd.w};
    for (int t = 0; t < 4; ++t) {
        int u = -1, jj = 0;
        if (t < di) {
            int row = a[t];
            for (int d = 1; d <= 4; ++d) {
                int lo = eo[d - 1], hi = lo + nn[d - 1] * d;
                if (row >= lo && row < hi) {
                    int o = row - lo;
                    u = sel[bo[d - 1] + o / d];
                    jj = o % d;
                }
            }
        }
        in_src[4 * v + t] = u;
        in_j[4 * v + t] = jj;
    }
}

}  // namespace mk

using namespace mk;

extern "C" int64_t molkgnn_bucket_scratch_bytes(int32_t N, int32_t E) {
    int64_t nblk = (N + BT - 1) / BT;
    // out_cnt[N] out_eid[4N] in_eid[4N] blk_counts[5*nblk] blk_off[5*nblk] totals[16] err[1] tinfo[8] (32 ints)
    // far[N] cutpos[N+1]
    return (int64_t)sizeof(int) * ((int64_t)N * 11 + nblk * 10 + 64);
}

// Side stream of the tile-cut chain (valid cuts -> greedy walk -> per-tile maxima).  The walk is a single-CTA sequential
// kernel; on its own stream it runs beside the counting kernels and whatever the caller queues before _finish (parameter
// packing), instead of holding up the launching stream.  One set per device, created on first use; the library is driven
// by one host thread per process (SURVEY 8(b) threading).
struct SideStream { cudaStream_t st; cudaEvent_t far_ready, deg_ready, done; bool ok; };
static SideStream* side_stream() {
    static SideStream s_side[16];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    SideStream& s = s_side[dev];
    if (!s.ok) {
        if (cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&s.far_ready, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.deg_ready, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        s.ok = true;
    }
    return &s;
}

// Balanced schedule of the tile-major kernels (tile.cuh TileWalk): stable counting sort of the tiles by node count, then
// n = q G + r: CTAs 0..r-1 (q + 1 tiles each) take the (q + 1) r smallest tiles, the other CTAs the rest; inside each group
// the sorted tiles are dealt boustrophedon (round k left to right, round k + 1 right to left) so that per-CTA node totals
// agree to ~1 %.  One CTA; the stable ranks come from warp 0 walking the tiles in index order (deterministic).
constexpr int ORDER_MAX_TILES = 16384;
static int g_tile_order = -1;       // 0 = round robin, 1 = schedule where it pays (default), 2 = always
static int tile_order_mode() {
    if (g_tile_order < 0) {
        const char* e = getenv("MOLKGNN_TILE_ORDER");
        g_tile_order = (e && e[0] == '0') ? 0 : (e && e[0] == '2') ? 2 : 1;
    }
    return g_tile_order;
}
__global__ void __launch_bounds__(1024) k_tile_order(int n_tiles, int G, const int* __restrict__ tile_start,
                                                     int* __restrict__ order) {
    __shared__ int hist[TNODES + 2];
    __shared__ unsigned char snn[ORDER_MAX_TILES];
    const int tid = threadIdx.x;
    for (int i = tid; i < TNODES + 2; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int t = tid; t < n_tiles; t += blockDim.x) {
        const int nn = min(max(tile_start[t + 1] - tile_start[t], 0), TNODES);
        snn[t] = (unsigned char)nn;
        atomicAdd(&hist[nn + 1], 1);
    }
    __syncthreads();
    if (tid == 0) for (int i = 1; i < TNODES + 2; ++i) hist[i] += hist[i - 1];     // hist[nn] = first sorted position of size nn
    __syncthreads();
    if (tid >= 32) return;
    const int q = n_tiles / G, r = n_tiles - q * G;
    const int nlong = (q + 1) * r, Gs = G - r;
    const unsigned lt = (1u << tid) - 1u;
    for (int t0 = 0; t0 < n_tiles; t0 += 32) {
        const int t = t0 + tid;
        const bool ok = t < n_tiles;
        const int nn = ok ? (int)snn[t] : TNODES + 1;
        const unsigned same = __match_any_sync(0xffffffffu, nn);
        const int s = hist[nn] + __popc(same & lt);
        __syncwarp();
        if (ok && (same & lt) == 0u) hist[nn] += __popc(same);
        __syncwarp();
        if (!ok) continue;
        int pos;
        if (s < nlong) {
            const int k = s / r, i = s - k * r;
            const int c = (k & 1) ? r - 1 - i : i;
            pos = c * (q + 1) + k;
        } else {
            const int idx = s - nlong;
            const int k = idx / Gs, i = idx - k * Gs;
            const int c = (k & 1) ? Gs - 1 - i : i;
            pos = nlong + c * q + k;
        }
        order[pos] = t;
    }
}


// phases: bit 0 = counting kernels (asynchronous), bit 1 = host round trip of the bucket sizes + assignment kernels
static int bucket_build_phases(molkgnn_plan_t* plan, const int64_t* edge_index, const float* p, int32_t p_dim,
                               const float* edge_attr, int32_t Fe, void* scratch, void* stream_, int phases,
                               const RefRows* ref_rows = nullptr, const int64_t* ref_nrows = nullptr) {
    cudaStream_t st = (cudaStream_t)stream_;
    const int N = plan->N, E = plan->E;
    MK_REQUIRE(N > 0 && E >= 0, "bucket_build: empty batch (N=%d E=%d)", N, E);
    MK_REQUIRE(Fe >= 1 && Fe <= EP, "bucket_build: edge_attr_dim %d not in 1..%d", Fe, EP);
    MK_REQUIRE(p_dim == 3, "bucket_build: only 3-D coordinates are supported (got %d)", p_dim);
    const int nblk = (N + BT - 1) / BT;
    int* s = (int*)scratch;
    int* out_cnt = s;             s += N;
    int* out_eid = s;             s += 4 * (size_t)N;
    int* in_eid = s;              s += 4 * (size_t)N;
    int* blk_counts = s;          s += 5 * (size_t)nblk;
    int* blk_off = s;             s += 5 * (size_t)nblk;
    int* totals = s;              s += 16;
    int* err = s;                 s += 1;
    int* tinfo = s;               s += 15;     // totals .. tinfo are one 32-int block copied to the host
    int* far = s;                 s += N;
    int* cutpos = s;
    const bool tiles = plan->tile_start != nullptr;
    const int tile_cap = N / TILE_MIN_STRIDE + 4;
    if (phases & 1) {
    ProfScope prof("bucket_count", st);
    MK_CHECK_CUDA(cudaMemsetAsync(out_cnt, 0, sizeof(int) * (size_t)N, st));
    MK_CHECK_CUDA(cudaMemsetAsync(plan->in_cnt, 0, sizeof(int) * (size_t)N, st));
    MK_CHECK_CUDA(cudaMemsetAsync(totals, 0, sizeof(int) * 32, st));
    MK_CHECK_CUDA(cudaMemsetAsync(far, 0, sizeof(int) * (size_t)N, st));
    count_launches((E > 0 ? 4 : 3) + (tiles ? 3 : 0));
    SideStream* side = tiles ? side_stream() : nullptr;
    MK_REQUIRE(!tiles || side, "bucket_build: cannot create the side stream");
    if (E > 0)
        k_edge_slots<<<(E + 255) / 256, 256, 0, st>>>(edge_index, E, N, out_cnt, out_eid, plan->in_cnt, in_eid, far, err);
    if (tiles) MK_CHECK_CUDA(cudaEventRecord(side->far_ready, st));
    k_node_prepare<<<nblk, BT, 0, st>>>(N, out_cnt, out_eid, plan->in_cnt, in_eid, plan->deg, blk_counts, err);
    if (tiles) MK_CHECK_CUDA(cudaEventRecord(side->deg_ready, st));
    k_scan_blocks<<<1, 160, 0, st>>>(nblk, blk_counts, blk_off, totals);
    if (tiles) {
        cudaStream_t main_st = st;
        cudaStream_t st = side->st;                       // the cut chain runs on the side stream from here on
        MK_CHECK_CUDA(cudaStreamWaitEvent(st, side->far_ready, 0));
        k_valid_cuts<<<(N + 1 + 255) / 256, 256, 0, st>>>(N, far, cutpos, tinfo);
        k_cut_gaps<<<(N + 1 + 255) / 256, 256, 0, st>>>(N, cutpos, tinfo);
        MK_CHECK_CUDA(cudaStreamWaitEvent(st, side->deg_ready, 0));
        const int64_t bitmap = (int64_t)((N + 32) / 32) * 4;
        static int s_budget = 0;
        if (!s_budget) s_budget = device_max_smem_optin();
        if (bitmap <= s_budget - 16384) {                  // + the kernel's static tables
            static int64_t s_attr_dev[16] = {0};
    int64_t& s_attr = s_attr_dev[device_index()];      // function attributes are per device
            if (bitmap > s_attr) {
                MK_CHECK_CUDA(cudaFuncSetAttribute(k_tile_starts_greedy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bitmap));
                s_attr = bitmap;
            }
            k_tile_starts_greedy<<<1, 1024, bitmap, st>>>(N, tile_cap, cutpos, plan->tile_start, tinfo);
            k_tile_stats<<<(tile_cap + 3) / 4, 128, 0, st>>>(tile_cap, plan->tile_start, plan->deg, tinfo);
            count_launches(1);
        } else {
            k_tile_starts<<<(tile_cap + 3) / 4, 128, 0, st>>>(N, tile_cap, cutpos, plan->deg, plan->tile_start, tinfo);
        }
        MK_CHECK_CUDA(cudaEventRecord(side->done, st));
        // The launching stream picks the cut chain's results up HERE, in the first half: the events are one set per device, and
        // a second _begin (another batch staged ahead on another stream) re-records them before this batch's _finish runs --
        // waiting there made _finish wait for the LATER batch's chain (and its host -> device copies).
        MK_CHECK_CUDA(cudaStreamWaitEvent(main_st, side->done, 0));
    }
    MK_CHECK_CUDA(cudaGetLastError());
    }
    if (!(phases & 2)) return 0;
    int host[32];
    MK_CHECK_CUDA(cudaMemcpyAsync(host, totals, sizeof(int) * 32, cudaMemcpyDeviceToHost, st));
    MK_CHECK_CUDA(cudaStreamSynchronize(st));
    plan->n_tiles = 0; plan->tile_max_nodes = 0;
    for (int d = 0; d < 4; ++d) plan->tile_max_deg[d] = 0;
    if (tiles && host[17 + 2] > 0 && host[17 + 2] < tile_cap && host[17 + 3] <= TILE_CAP) {
        plan->n_tiles = host[17 + 2];
        plan->tile_max_nodes = host[17 + 3];
        for (int d = 0; d < 4; ++d) plan->tile_max_deg[d] = host[17 + 4 + d];
    }
    MK_REQUIRE(host[16] == 0 && host[4] == 0,
               "bucket_build: unsupported graph (flags=0x%x: 1=node id out of range, 2=out-degree>4, 4=in-degree>4, "
               "8=node with out-degree 0 or >4; %d offending nodes)", host[16], host[4]);
    for (int d = 0; d < 4; ++d) { plan->n[d] = host[d]; plan->boff[d] = host[5 + d]; plan->eoff[d] = host[9 + d]; }
    ProfScope prof2("bucket_assign", st);
    RefRows ref;
    memset(&ref, 0, sizeof(ref));
    if (ref_rows) ref = *ref_rows;
    for (int d = 0; d < 4; ++d)
        MK_REQUIRE(!ref_rows || plan->n[d] == 0 || ref.nea[d],
                   "bucket_build: the batch has %d nodes of degree %d but no nei_edge_attr_deg%d tensor was given", plan->n[d],
                   d + 1, d + 1);
    for (int d = 0; d < 4 && ref_nrows; ++d)      // checked BEFORE the kernel reads the rows
        MK_REQUIRE(ref_nrows[d] == (int64_t)plan->n[d] * (d + 1),
                   "bucket_build: nei_edge_attr_deg%d has %lld rows but edge_index gives %d nodes of degree %d (inconsistent batch)",
                   d + 1, (long long)ref_nrows[d], plan->n[d], d + 1);
    MK_REQUIRE(ref.nea[0] || ref.nea[1] || ref.nea[2] || ref.nea[3] || edge_attr, "bucket_build: no bond attributes");
    k_assign<<<nblk, BT, 0, st>>>(N, E, edge_index, plan->deg, out_eid, plan->in_cnt, in_eid, blk_off, totals, p, p_dim,
                                  edge_attr, Fe, plan->pos, plan->sel, plan->nei, plan->nei_eid, plan->ehat,
                                  plan->tsign, plan->in_src, plan->in_j, ref);
    if (plan->n_tiles > 0 && plan->tile_meta && plan->ehat_node) {
        count_launches(1);
        k_tile_meta<<<plan->n_tiles, TNODES, 0, st>>>(N, plan->tile_start, plan->deg, plan->pos, plan->nei, plan->ehat,
                                                    plan->tsign, plan->in_cnt, plan->in_src, plan->in_j, blk_off, totals,
                                                    reinterpret_cast<TileMetaG*>(plan->tile_meta), plan->ehat_node,
                                                    plan->node_tile);
        plan->tile_grid = 0;
        static int s_sms = 0;
        if (!s_sms) s_sms = device_num_sms();
        const int s_order = tile_order_mode();
        // Used where it was measured to pay: a few CTAs with one tile more than the rest (0 < r <= G / 2), where those CTAs
        // alone decide the kernel time (892 tiles on 148 SMs: 1.328 -> 1.303 ms per step).  With most CTAs on the longer
        // walk (887 tiles: r = 147) round robin measured the same or slightly better.  MOLKGNN_TILE_ORDER=2 forces it.
        const int G = std::max(1, std::min(plan->n_tiles, s_sms));
        const int r = plan->n_tiles % G;
        if (s_order && plan->tile_order && s_sms > 0 && plan->n_tiles <= ORDER_MAX_TILES && (s_order == 2 || (r > 0 && 2 * r <= G))) {
            count_launches(1);
            k_tile_order<<<1, 1024, 0, st>>>(plan->n_tiles, G, plan->tile_start, plan->tile_order);
            plan->tile_grid = G;
        }
    } else {
        plan->n_tiles = 0;
        plan->tile_grid = 0;
    }
    MK_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int molkgnn_bucket_build(molkgnn_plan_t* plan, const int64_t* edge_index, const float* p, int32_t p_dim,
                                    const float* edge_attr, int32_t Fe, void* scratch, void* stream) {
    return bucket_build_phases(plan, edge_index, p, p_dim, edge_attr, Fe, scratch, stream, 3);
}
extern "C" int molkgnn_bucket_build_begin(molkgnn_plan_t* plan, const int64_t* edge_index, const float* p, int32_t p_dim,
                                          const float* edge_attr, int32_t Fe, void* scratch, void* stream) {
    return bucket_build_phases(plan, edge_index, p, p_dim, edge_attr, Fe, scratch, stream, 1);
}
extern "C" int molkgnn_bucket_build_finish(molkgnn_plan_t* plan, const int64_t* edge_index, const float* p, int32_t p_dim,
                                           const float* edge_attr, int32_t Fe, void* scratch, void* stream) {
    return bucket_build_phases(plan, edge_index, p, p_dim, edge_attr, Fe, scratch, stream, 2);
}

// _finish with the reference batch's own data tensors as the source of the bond rows / degree-4 coordinates (see RefRows).
// n_rows[d] = rows of nei_edge_attr[d] the caller holds (n_d * (d+1), checked against the bucket sizes found on the GPU).
extern "C" int molkgnn_bucket_build_finish_ref(molkgnn_plan_t* plan, const int64_t* edge_index, const float* p, int32_t p_dim,
                                               const float* const nei_edge_attr[4], const int64_t n_rows[4], int32_t Fe,
                                               const float* p_focal4, const float* nei_p4, void* scratch, void* stream) {
    RefRows ref;
    for (int d = 0; d < 4; ++d) ref.nea[d] = nei_edge_attr[d];
    ref.pf4 = p_focal4; ref.np4 = nei_p4;
    return bucket_build_phases(plan, edge_index, p, p_dim, nullptr, Fe, scratch, stream, 2, &ref, n_rows);
}

extern "C" int molkgnn_set_tile_order(int mode) {
    const int old = tile_order_mode();
    g_tile_order = mode < 0 ? 0 : mode > 2 ? 2 : mode;
    return old;
}

extern "C" int64_t molkgnn_tile_meta_bytes(void) { return (int64_t)sizeof(TileMetaG); }

extern "C" int molkgnn_bucket_export(const molkgnn_plan_t* plan, int32_t d, const float* p, int32_t p_dim,
                                     const float* edge_attr, int32_t Fe, int64_t* selected_index, int64_t* nei_index,
                                     float* p_focal, float* nei_p, float* nei_edge_attr, void* stream_) {
    MK_REQUIRE(d >= 1 && d <= 4, "bucket_export: degree %d", d);
    int n = plan->n[d - 1];
    if (n == 0) return 0;
    count_launches(1);
    k_export<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream_>>>(d, n, plan->boff[d - 1], plan->eoff[d - 1], plan->sel,
                                                                plan->nei, plan->nei_eid, p, p_dim, edge_attr, Fe,
                                                                selected_index, nei_index, p_focal, nei_p,
                                                                nei_edge_attr);
    MK_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int molkgnn_plan_from_buckets(molkgnn_plan_t* plan, const int64_t* const selected_index[4],
                                         const int64_t* const nei_index[4], const float* const p_focal[4],
                                         const float* const nei_p[4], int32_t p_dim,
                                         const float* const nei_edge_attr[4], int32_t Fe, void* stream_) {
    // plan->N, plan->n[] must be set by the caller (tensor shapes); boff/eoff are derived here.
    cudaStream_t st = (cudaStream_t)stream_;
    plan->n_tiles = 0; plan->tile_max_nodes = 0;     // bucket tensors carry no node order guarantees: no tiling
    for (int d = 0; d < 4; ++d) plan->tile_max_deg[d] = 0;
    const int N = plan->N;
    MK_REQUIRE(Fe >= 1 && Fe <= EP, "plan_from_buckets: edge_attr_dim %d not in 1..%d", Fe, EP);
    int bo = 0, eo = 0, tot = 0;
    for (int d = 0; d < 4; ++d) {
        plan->boff[d] = bo; plan->eoff[d] = eo;
        bo += plan->n[d]; eo += plan->n[d] * (d + 1); tot += plan->n[d];
    }
    // nodes outside every bucket (deg 0) are allowed here: they only ever act as neighbours (KernelConv.forward)
    MK_REQUIRE(tot <= N, "plan_from_buckets: buckets hold %d focal nodes but the batch has %d nodes", tot, N);
    MK_REQUIRE(eo == plan->E, "plan_from_buckets: plan->E=%d but buckets hold %d neighbour rows", plan->E, eo);
    // in_j doubles as the key scratch (4 ints per node) until the in-lists are decoded; err flag lives in in_cnt? no:
    // use the last int of nei_eid's allocation is not safe either -> use a static device flag.
    static int* d_err_dev[16] = {nullptr};
    int*& d_err = d_err_dev[device_index()];                  // one error word per device
    if (!d_err) MK_CHECK_CUDA(cudaMalloc(&d_err, sizeof(int)));
    MK_CHECK_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), st));
    MK_CHECK_CUDA(cudaMemsetAsync(plan->in_cnt, 0, sizeof(int) * (size_t)N, st));
    MK_CHECK_CUDA(cudaMemsetAsync(plan->deg, 0, sizeof(int) * (size_t)N, st));
    MK_CHECK_CUDA(cudaMemsetAsync(plan->pos, 0, sizeof(int) * (size_t)N, st));
    for (int d = 1; d <= 4; ++d) {
        int n = plan->n[d - 1];
        if (!n) continue;
        count_launches(1);
        k_from_buckets<<<(n + 127) / 128, 128, 0, st>>>(d, n, plan->boff[d - 1], plan->eoff[d - 1], selected_index[d - 1],
                                                       nei_index[d - 1], p_focal[d - 1], nei_p[d - 1], p_dim,
                                                       nei_edge_attr[d - 1], Fe, N, plan->deg, plan->pos, plan->sel,
                                                       plan->nei, plan->nei_eid, plan->ehat, plan->tsign, plan->in_cnt,
                                                       plan->in_j, d_err);
    }
    count_launches(1);
    k_from_buckets_inlists<<<(N + 127) / 128, 128, 0, st>>>(
        N, plan->in_cnt, plan->in_j, plan->sel, make_int4(plan->boff[0], plan->boff[1], plan->boff[2], plan->boff[3]),
        make_int4(plan->eoff[0], plan->eoff[1], plan->eoff[2], plan->eoff[3]),
        make_int4(plan->n[0], plan->n[1], plan->n[2], plan->n[3]), plan->in_src, plan->in_j);
    int herr = 0;
    MK_CHECK_CUDA(cudaMemcpyAsync(&herr, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    MK_CHECK_CUDA(cudaStreamSynchronize(st));
    MK_REQUIRE(herr == 0, "plan_from_buckets: invalid bucket tensors (flags=0x%x: 1=index out of range, 4=in-degree>4)", herr);
    return 0;
}
