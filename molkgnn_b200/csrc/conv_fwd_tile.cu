// Molecule-tile forward of the molecular-kernel convolution (sm_100a, tcgen05 + TMEM).  Same contract as k_conv_fwd
// (conv_fwd.cu): replaces KernelConv.calculate_total_score (reference kernels.py:353-425) and the bucket gathers / output
// assembly of BaseKernelSetConv.forward (kernels.py:519-548, 674-747) for plans that carry molecule tiles.
//
// Formulation.  A tile is a run of <= 128 consecutive nodes holding whole molecules, so every neighbour of a tile node
// is itself a tile node.  Per tile ONE dense GEMM on the tensor cores gives every cosine the layer needs,
//        T[(k,s), v] = shat[k,s,:] . xhat[v,:]            (M = kernel rows, N = tile nodes, K = F)
// with the kernel rows on the TMEM lanes and the nodes on the TMEM columns: the d x d similarity tile of a
// (node, kernel) pair is then rows (k,0..d-1) = 4 adjacent lanes, columns nei(n,0..d-1) -- a thread reads "its" row at
// its node's neighbour columns straight from TMEM (tcgen05.ld, one column per load) and there is no gather of
// neighbour feature rows at all: x is read once per role, contiguously.  fp32 accuracy: both operands are split into
// unscaled fp16 pairs v = hi + lo and three UMMAs per K step (hi*hi, lo*hi, hi*lo) accumulate into one fp32 accumulator.
//
// The kernel rows do not fit shared memory together with a node tile, so they are split into two roles (TileRows,
// common.cuh: role 0 = degree 4, role 1 = degrees 3, 2, 1); every persistent CTA runs role 0 over a dynamic queue of
// tiles with that role's images resident in shared memory, then role 1.  Per tile: x rows -> normalised fp16 images in
// shared memory, 2 x (F/16) x 3 UMMAs into one of two TMEM accumulator sets (the next tile's MMAs overlap the second
// half of the current tile's epilogue), and the epilogue: per (node, kernel) the 4 lanes of a kernel exchange rows by
// shuffle, each evaluates its share of the permutations in the reference's arithmetic (sequential mean, first-max
// arg-max kernels.py:373), then bond cosine at the arg-max (kernels.py:382-390), chirality (kernels.py:279-350) and the
// softmax mix (kernels.py:402-425).
#include <algorithm>
#include "common.cuh"
#include "tc.cuh"

namespace mk {

bool tile_layer_ok(const molkgnn_layer_t* layer);

constexpr int TF_THREADS = 512;
constexpr int TF_WARPS = TF_THREADS / 32;
constexpr int TNODES = MOLKGNN_TILE_NODES;      // 128
constexpr int TF_ESLOTS = 4 * TNODES;           // neighbour (bond) slots of a tile

// per-tile metadata in shared memory (two copies: the next tile is prepared while the current one is still in its epilogue)
struct __align__(128) TileMeta {
    float ehat[TF_ESLOTS][EP];        // normalised bond rows of the role's nodes, slot = eslot[n] + j
    uint32_t nl[TNODES];              // 4 local neighbour ids, 8 bits each
    int posl[TNODES];                 // bucket row R of the node
    unsigned short eslot[TNODES];
    unsigned char list[4][TNODES];    // local ids of the degree-d nodes, ascending
    signed char tsg[TNODES];          // degree-4 only: sign of the neighbour triple product
    unsigned char dup[TNODES];        // degree-4 only: two identical neighbour rows (chirality gate)
    int cnt[4];
    int t0, nn;
    int wcnt[4][5];                   // scratch: per-warp counts (4 degrees + slots)
};

struct FwdTileArgs {
    const float* x; const float* xnorm; int ldx;
    int F, Fp, Fk;
    const int* deg; const int* pos; const int* nei; const float* ehat; const int8_t* tsign;
    const int* tile_start; int n_tiles;
    int n[4], boff[4], eoff[4], L[4], koff[4];
    const float* packed[4];
    const unsigned char* img;
    int img_one;                      // bytes of one image (hi or lo) of one role
    int is_last;
    float* sc; int sc_mode; int ld_sc; long long scoff[4];
    uint8_t* argmax; uint8_t* argmax_free; const uint8_t* argmax_in;
    int* counter;                     // [2] tile queues
    int sm_img, sm_x, sm_meta;        // byte offsets into dynamic shared memory
    int x_one;                        // bytes of one node-tile image
};

// per-lane description of "its" kernel row, fixed for a whole role
struct LaneRow {
    int d, k, slot;                   // d = 0: unused lane
    float es[EP];                     // normalised bond support row (support lanes)
    float ws, wc, we, W, rW;
    uint32_t pc[3];                   // packed permutation codes of the permutations this lane evaluates
    const int8_t* supsign;            // degree 4: chirality sign of every (kernel, permutation)
};

template <int D> __device__ __forceinline__ uint32_t perm_code_rt(int p) {
    uint32_t c = 0;
#pragma unroll
    for (int q = 0; q < Perm<D>::P; ++q) if (q == p) c = perm_code<D>(q);
    return c;
}
template <int D> __device__ __forceinline__ uint32_t perm_inv_code_rt(int p) {
    uint32_t c = 0;
#pragma unroll
    for (int q = 0; q < Perm<D>::P; ++q) if (q == p) c = perm_inv_code<D>(q);
    return c;
}

__device__ __forceinline__ void tf_copy16(unsigned char* dst, const unsigned char* src, int64_t bytes) {
    for (int64_t i = (int64_t)threadIdx.x * 16; i < bytes; i += (int64_t)TF_THREADS * 16)
        *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(src + i);
}

__device__ __forceinline__ size_t tf_sc_index(const FwdTileArgs& a, int d, int R, int L, int k, int node) {
    return a.sc_mode == 0 ? (size_t)a.scoff[d - 1] + (size_t)R * L + k : (size_t)node * a.ld_sc + a.koff[d - 1] + k;
}

// ---- tile preparation: metadata + normalised fp16 images of the tile's x rows ---------------------------------------
__device__ __forceinline__ void tf_prepare(const FwdTileArgs& a, unsigned char* smem, TileMeta* mt, int tile, int role) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t0 = a.tile_start[tile], t1 = a.tile_start[tile + 1];
    const int nn = t1 - t0;
    // (1) node-tile images: thread (row r, column group cg); 8 consecutive rows x 16 B are contiguous in the layout
    {
        const int r = tid & (TNODES - 1), cg = tid >> 7;
        const bool ok = r < nn;
        float rinv = 0.f;
        const float* xr = a.x;
        if (ok) {
            rinv = 1.0f / fmaxf(a.xnorm[t0 + r], MOLKGNN_COS_EPS);
            xr = a.x + (size_t)(t0 + r) * a.ldx;
        }
        unsigned char* Xhi = smem + a.sm_x;
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
    // (2) per-node metadata (threads 0..127 = local node id)
    int d = 0, R = 0, base = 0, rank[4] = {0, 0, 0, 0}, srank = 0;
    bool mine = false;
    if (tid < TNODES) {
        if (tid < nn) {
            d = a.deg[t0 + tid];
            R = a.pos[t0 + tid];
            base = a.eoff[d - 1] + R * d;
        }
        mine = d > 0 && (role == 0 ? d == 4 : d <= 3) && a.L[d - 1] > 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const unsigned m = __ballot_sync(0xffffffffu, mine && d == c + 1);
            rank[c] = __popc(m & ((1u << lane) - 1u));
            if (lane == 0) mt->wcnt[warp][c] = __popc(m);
        }
        // exclusive prefix of the bond slots inside the warp
        int s = mine ? d : 0;
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        srank = incl - s;
        if (lane == 31) mt->wcnt[warp][4] = incl;
    }
    __syncthreads();
    if (tid < TNODES) {
        int soff = srank;
        int loff[4] = {rank[0], rank[1], rank[2], rank[3]};
        for (int w = 0; w < warp; ++w) {
            soff += mt->wcnt[w][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) loff[c] += mt->wcnt[w][c];
        }
        if (tid == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) mt->cnt[c] = mt->wcnt[0][c] + mt->wcnt[1][c] + mt->wcnt[2][c] + mt->wcnt[3][c];
            mt->t0 = t0; mt->nn = nn;
        }
        if (mine) {
            mt->list[d - 1][loff[d - 1]] = (unsigned char)tid;
            mt->posl[tid] = R;
            mt->eslot[tid] = (unsigned short)soff;
            uint32_t w = 0;
            for (int j = 0; j < d; ++j) {
                const int u = a.nei[(size_t)base + j] - t0;
                w |= (uint32_t)(u & 0xff) << (8 * j);
                const float4* src = reinterpret_cast<const float4*>(a.ehat + ((size_t)base + j) * EP);
                float4* dst = reinterpret_cast<float4*>(&mt->ehat[soff + j][0]);
                dst[0] = __ldg(src);
                dst[1] = __ldg(src + 1);
            }
            mt->nl[tid] = w;
            if (d == 4) mt->tsg[tid] = a.is_last ? a.tsign[R] : 0;
        }
    }
    // (3) chirality gate of the degree-4 nodes: any two of the four neighbour feature rows bit-equal (torch.equal,
    //     kernels.py:310-317); one warp per node, raw rows from global memory
    if (role == 0 && a.is_last) {
        __syncthreads();
        const int n4 = mt->cnt[3];
        for (int i = warp; i < n4; i += TF_WARPS) {
            const int nl_ = mt->list[3][i];
            const uint32_t w = mt->nl[nl_];
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
            if (lane == 0) mt->dup[nl_] = (neq != 0x3fu) ? 1 : 0;
        }
    }
}

// one thread: 2 M blocks x (Fk/16) K steps x 3 UMMAs into accumulator set `set`
__device__ __forceinline__ void tf_issue(const FwdTileArgs& a, unsigned char* smem, int role, int nn, uint32_t tmem,
                                         int set, uint64_t* bar) {
    const TileRows tr(a.L);
    const int used = tr.rows_used(role);
    const int nmb = used > 128 ? 2 : 1;
    const uint32_t sbo = (uint32_t)(a.Fk >> 3) * 128u;
    const uint32_t ihi = tc::smem_u32(smem + a.sm_img), ilo = ihi + (uint32_t)a.img_one;
    const uint32_t xhi = tc::smem_u32(smem + a.sm_x), xlo = xhi + (uint32_t)a.x_one;
    const int N = max(16, (nn + 15) & ~15);
    const uint32_t idesc = tc::idesc_f16(128, N, 0, 0);
    const int nks = a.Fk >> 4;
    for (int mb = 0; mb < nmb; ++mb) {
        const uint32_t d = tmem + (uint32_t)(set * 256 + mb * 128);
        const uint32_t ro = (uint32_t)mb * 16u * sbo;      // 128 rows = 16 row groups
        for (int ks = 0; ks < nks; ++ks) {
            const uint32_t o = (uint32_t)ks * 256u;
            const uint64_t dAh = tc::smem_desc(ihi + ro + o, 128u, sbo), dAl = tc::smem_desc(ilo + ro + o, 128u, sbo);
            const uint64_t dBh = tc::smem_desc(xhi + o, 128u, sbo), dBl = tc::smem_desc(xlo + o, 128u, sbo);
            tc::umma_f16(d, dAh, dBh, idesc, ks > 0 ? 1u : 0u);
            tc::umma_f16(d, dAl, dBh, idesc, 1u);
            tc::umma_f16(d, dAh, dBl, idesc, 1u);
        }
    }
    tc::umma_commit(bar);
}

// ---- epilogue ------------------------------------------------------------------------------------------------------
// Degree-4 centre rows: C[n,k] parked in the score slot of (n,k); the leader lane of the kernel reads it back.
__device__ __forceinline__ void tf_centre_pass(const FwdTileArgs& a, const TileMeta* mt, const LaneRow& lr, uint32_t tb,
                                               int slice, int nslices) {
    const bool cen = lr.d == 4 && lr.slot == 4;
    if (!__any_sync(0xffffffffu, cen)) return;
    const int n4 = mt->cnt[3], L = a.L[3];
    for (int i = slice; i < n4; i += nslices) {
        const int nl_ = mt->list[3][i];
        uint32_t v = tc::tmem_ld1(tb + (uint32_t)nl_);
        tc::tmem_ld_wait();
        asm volatile("" : "+r"(v) :: "memory");
        if (cen) a.sc[tf_sc_index(a, 4, mt->posl[nl_], L, lr.k, mt->t0 + nl_)] = __uint_as_float(v);
    }
}

// all (node, kernel) pairs of the degree-D nodes list[i0], list[i0 + istep], ... (< i1) against this warp's kernel rows
template <int D>
__device__ __forceinline__ void tf_visit(const FwdTileArgs& a, const TileMeta* mt, const LaneRow& lr, uint32_t tb,
                                         int i0, int i1, int istep) {
    constexpr int P = Perm<D>::P, PPL = P / D;
    const int lane = threadIdx.x & 31;
    const bool act = lr.d == D;
    if (!__any_sync(0xffffffffu, act)) return;
    const bool sup = act && lr.slot < D;
    const bool leader = sup && lr.slot == 0;
    const int gbase = lane & ~3;
    const int L = a.L[D - 1];
    const float NEG = -3.0e38f;
    for (int i = i0; i < i1; i += istep) {
        const int nl_ = mt->list[D - 1][i];
        const uint32_t nw = mt->nl[nl_];
        const int R = mt->posl[nl_];
        const int e0 = mt->eslot[nl_];
        const size_t cidx = (size_t)a.scoff[D - 1] + (size_t)R * L + lr.k;
        const size_t oidx = tf_sc_index(a, D, R, L, lr.k, mt->t0 + nl_);
        float cdot = 0.f;
        if (D == 4 && leader) cdot = __ldcg(a.sc + oidx);
        int forced = -1;
        if (a.argmax_in && act) forced = a.argmax_in[cidx] & 0x7f;
        // this lane's kernel row at the D neighbour columns (and at the node's own column for the centre row)
        uint32_t tv[D], tcen = 0;
#pragma unroll
        for (int j = 0; j < D; ++j) tv[j] = tc::tmem_ld1(tb + ((nw >> (8 * j)) & 0xffu));
        if (D < 4) tcen = tc::tmem_ld1(tb + (uint32_t)nl_);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < D; ++j) asm volatile("" : "+r"(tv[j]) :: "memory");   // consumers stay behind the wait
        asm volatile("" : "+r"(tcen) :: "memory");
        float t[D];
#pragma unroll
        for (int j = 0; j < D; ++j) t[j] = __uint_as_float(tv[j]);
        // mean over j for this lane's permutations (pi(0) = slot): sequential sum, true division (kernels.py:194)
        float val[PPL];
#pragma unroll
        for (int q = 0; q < PPL; ++q) {
            float s = t[0];
#pragma unroll
            for (int j = 1; j < D; ++j) s += __shfl_sync(0xffffffffu, t[j], gbase + (int)((lr.pc[q] >> (2 * j)) & 3u));
            val[q] = div_deg<D>(s);
        }
        float best = NEG;
        int bi = 127;
        if (sup) {
            best = val[0]; bi = PPL * lr.slot;
#pragma unroll
            for (int q = 1; q < PPL; ++q) if (val[q] > best) { best = val[q]; bi = PPL * lr.slot + q; }   // first max wins
        }
        if (D > 1) {
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int obi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && obi < bi)) { best = ob; bi = obi; }
            }
        }
        if (a.argmax_in) {          // teacher forcing (parity harness / replay): warp-uniform branch
            float mv = 0.f;
#pragma unroll
            for (int q = 0; q < PPL; ++q) if (forced == PPL * lr.slot + q) mv = val[q];
            const int fl = (forced >= 0 && forced < P) ? forced / PPL : 0;
            const float used = __shfl_sync(0xffffffffu, mv, gbase + fl);
            if (leader && a.argmax_free) a.argmax_free[cidx] = (uint8_t)bi;
            if (forced >= 0 && forced < P) { bi = forced; best = used; }
        } else if (leader && a.argmax_free) {
            a.argmax_free[cidx] = (uint8_t)bi;
        }
        if (bi >= P) bi = 0;        // unused lanes: keep the table lookups in range
        // bond-attribute cosine at the chosen permutation (kernels.py:382-390): support lane s pairs with neighbour
        // j = pi^-1(s); the leader sums the D terms in neighbour order
        const uint32_t code = perm_code_rt<D>(bi);
        const uint32_t inv = perm_inv_code_rt<D>(bi);
        float dj = 0.f;
        if (sup) {
            const int j = (int)((inv >> (2 * lr.slot)) & 3u);
            const float4 q0 = *reinterpret_cast<const float4*>(&mt->ehat[e0 + j][0]);
            const float4 q1 = *reinterpret_cast<const float4*>(&mt->ehat[e0 + j][4]);
            dj = fmaf(q0.x, lr.es[0], dj); dj = fmaf(q0.y, lr.es[1], dj); dj = fmaf(q0.z, lr.es[2], dj); dj = fmaf(q0.w, lr.es[3], dj);
            dj = fmaf(q1.x, lr.es[4], dj); dj = fmaf(q1.y, lr.es[5], dj); dj = fmaf(q1.z, lr.es[6], dj); dj = fmaf(q1.w, lr.es[7], dj);
        }
        float esum = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const float v = __shfl_sync(0xffffffffu, dj, gbase + (int)((code >> (2 * j)) & 3u));
            esum = j == 0 ? v : esum + v;
        }
        if (D < 4) cdot = __shfl_sync(0xffffffffu, __uint_as_float(tcen), gbase + D);
        if (leader) {
            const float E = div_deg<D>(esum);
            float sc = div_by((best * lr.ws + cdot * lr.wc) + E * lr.we, lr.W, lr.rW);
            uint8_t am = (uint8_t)bi;
            if (D == 4 && a.is_last) {
                // chirality (kernels.py:279-350): +1 if any two neighbours are identical, else sign agreement
                int chi = 1;
                if (!mt->dup[nl_]) chi = (mt->tsg[nl_] == lr.supsign[lr.k * 12 + bi]) ? 1 : -1;
                if (chi < 0) { sc = -sc; am |= 0x80; }
            }
            a.argmax[cidx] = am;
            a.sc[oidx] = sc;
        }
    }
}

__device__ __forceinline__ void tf_epilogue(const FwdTileArgs& a, const TileMeta* mt, const LaneRow& lr, uint32_t tb,
                                            int role, int slice, int nslices, int part) {
    // the node list of every degree is cut in two parts: the next tile is prepared between them
    if (role == 0) {
        const int c = mt->cnt[3], h = (c + 1) >> 1;
        tf_visit<4>(a, mt, lr, tb, (part ? h : 0) + slice, part ? c : h, nslices);
    } else {
        int c = mt->cnt[2], h = (c + 1) >> 1;
        tf_visit<3>(a, mt, lr, tb, (part ? h : 0) + slice, part ? c : h, nslices);
        c = mt->cnt[1]; h = (c + 1) >> 1;
        tf_visit<2>(a, mt, lr, tb, (part ? h : 0) + slice, part ? c : h, nslices);
        c = mt->cnt[0]; h = (c + 1) >> 1;
        tf_visit<1>(a, mt, lr, tb, (part ? h : 0) + slice, part ? c : h, nslices);
    }
}

__device__ __forceinline__ void tf_lane_row(const FwdTileArgs& a, int role, int row, LaneRow& lr) {
    const TileRows tr(a.L);
    tr.describe(role, row, lr.d, lr.k, lr.slot);
#pragma unroll
    for (int c = 0; c < EP; ++c) lr.es[c] = 0.f;
    lr.ws = lr.wc = lr.we = 0.f; lr.W = 1.f; lr.rW = 1.f;
    lr.pc[0] = lr.pc[1] = lr.pc[2] = 0;
    lr.supsign = nullptr;
    if (lr.d == 0) return;
    const int L = a.L[lr.d - 1];
    const PackedLayout pl(lr.d, L, a.Fp);
    const float* pk = a.packed[lr.d - 1];
    lr.ws = pk[pl.w + 0]; lr.wc = pk[pl.w + 1]; lr.we = pk[pl.w + 2]; lr.W = pk[pl.w + 3];
    lr.rW = 1.0f / lr.W;
    lr.supsign = reinterpret_cast<const int8_t*>(pk + pl.sign);
    if (lr.slot < 4) {
        const float* es = pk + pl.es + (size_t)(lr.slot * L + lr.k) * EP;
#pragma unroll
        for (int c = 0; c < EP; ++c) lr.es[c] = es[c];
        const int ppl = num_perms(lr.d) / lr.d;
        for (int q = 0; q < ppl; ++q) {
            const int p = ppl * lr.slot + q;
            lr.pc[q] = lr.d == 4 ? perm_code_rt<4>(p) : lr.d == 3 ? perm_code_rt<3>(p) : lr.d == 2 ? perm_code_rt<2>(p) : 0u;
        }
    }
}

__global__ void __launch_bounds__(TF_THREADS, 1) k_conv_fwd_tile(const __grid_constant__ FwdTileArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tslot;
    __shared__ int s_tile[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = warp & 3, mb = (warp >> 2) & 1, slice = warp >> 3;
    constexpr int NSL = TF_WARPS / 8;
    if (tid == 0) { tc::mbar_init(&bars[0], 1); tc::mbar_init(&bars[1], 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&tslot, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tslot;
    TileMeta* meta = reinterpret_cast<TileMeta*>(smem + a.sm_meta);
    uint32_t phase[2] = {0u, 0u};

    for (int role = 0; role < 2; ++role) {
        const TileRows tr(a.L);
        if (tr.rows_used(role) == 0) continue;
        __syncthreads();                       // previous role completely finished (its MMAs were all waited for)
        tf_copy16(smem + a.sm_img, a.img + (size_t)role * 2 * a.img_one, 2 * (int64_t)a.img_one);
        LaneRow lr;
        tf_lane_row(a, role, mb * 128 + q * 32 + lane, lr);
        if (tid == 0) s_tile[0] = atomicAdd(a.counter + role, 1);
        __syncthreads();
        int t = s_tile[0];
        int set = 0;
        if (t < a.n_tiles) {
            tf_prepare(a, smem, &meta[0], t, role);
            tc::fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                tc::fence_after_sync();
                tf_issue(a, smem, role, meta[0].nn, tmem, 0, &bars[0]);
            }
        }
        while (t < a.n_tiles) {
            const TileMeta* mt = &meta[set];
            if (tid == 0) s_tile[1] = atomicAdd(a.counter + role, 1);
            tc::mbar_wait(&bars[set], phase[set]);
            phase[set] ^= 1u;
            tc::fence_after_sync();
            const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * 256 + mb * 128);
            if (role == 0) tf_centre_pass(a, mt, lr, tb, slice, NSL);
            __syncthreads();                   // centre scores parked; next tile id published
            const int tn = s_tile[1];
            tf_epilogue(a, mt, lr, tb, role, slice, NSL, 0);
            tc::fence_before_sync();
            __syncthreads();                   // every warp is past the previous tile: its metadata / TMEM set are free
            if (tn < a.n_tiles) {
                tf_prepare(a, smem, &meta[set ^ 1], tn, role);
                tc::fence_async_smem();
                __syncthreads();
                if (tid == 0) {
                    tc::fence_after_sync();
                    tf_issue(a, smem, role, meta[set ^ 1].nn, tmem, set ^ 1, &bars[set ^ 1]);
                }
            }
            tf_epilogue(a, mt, lr, tb, role, slice, NSL, 1);
            tc::fence_before_sync();
            t = tn;
            set ^= 1;
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ---- host side -----------------------------------------------------------------------------------------------------
// returns 1 if launched, 0 if the plan / layer is not eligible (caller falls back to the bucket-order kernels), <0 on error
int launch_conv_fwd_tile(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                         const float* xnorm, int32_t is_last_layer, float* sc, int32_t sc_mode, int32_t ld_sc,
                         const int64_t scoff[4], uint8_t* argmax, uint8_t* argmax_free, const uint8_t* argmax_in,
                         int32_t* counter, cudaStream_t st) {
    if (plan->n_tiles <= 0 || !plan->tile_start || !layer->tile_img || !tile_layer_ok(layer)) return 0;
    if (plan->tile_max_nodes > TNODES) return 0;
    static int s_budget = 0, s_sms = 0;
    if (!s_budget) {
        s_budget = device_max_smem_optin();
        s_sms = device_num_sms();
        MK_REQUIRE(s_budget > 0 && s_sms > 0, "conv_fwd_tile: no CUDA device");
    }
    FwdTileArgs a;
    a.x = x; a.xnorm = xnorm; a.ldx = ldx;
    a.F = layer->F; a.Fp = layer->Fp; a.Fk = tile_fk(layer->Fp);
    a.deg = plan->deg; a.pos = plan->pos; a.nei = plan->nei; a.ehat = plan->ehat; a.tsign = plan->tsign;
    a.tile_start = plan->tile_start; a.n_tiles = plan->n_tiles;
    for (int d = 0; d < 4; ++d) {
        a.n[d] = plan->n[d]; a.boff[d] = plan->boff[d]; a.eoff[d] = plan->eoff[d];
        a.L[d] = layer->L[d]; a.koff[d] = layer->koff[d];
        a.packed[d] = layer->packed[d];
        a.scoff[d] = scoff[d];
    }
    a.img = reinterpret_cast<const unsigned char*>(layer->tile_img);
    a.img_one = (int)tile_img_bytes_one(a.Fk);
    a.x_one = (int)tc::il_tile_bytes(TNODES, a.Fk);
    a.is_last = is_last_layer;
    a.sc = sc; a.sc_mode = sc_mode; a.ld_sc = ld_sc;
    a.argmax = argmax; a.argmax_free = argmax_free; a.argmax_in = argmax_in;
    a.counter = counter;
    int64_t off = 0;
    a.sm_img = (int)off; off += 2 * (int64_t)a.img_one;
    a.sm_x = (int)off; off += 2 * (int64_t)a.x_one;
    a.sm_meta = (int)off; off += 2 * (int64_t)((sizeof(TileMeta) + 127) / 128 * 128);
    if (off > s_budget - 1024) return 0;
    MK_CHECK_CUDA(cudaMemsetAsync(counter, 0, 2 * sizeof(int), st));
    static int64_t s_attr = 0;
    if (off > s_attr) {
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_conv_fwd_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)off));
        s_attr = off;
    }
    const int grid = std::min(plan->n_tiles, s_sms);
    count_launches(1);
    k_conv_fwd_tile<<<grid, TF_THREADS, off, st>>>(a);
    MK_CHECK_CUDA(cudaGetLastError());
    return 1;
}

}  // namespace mk
