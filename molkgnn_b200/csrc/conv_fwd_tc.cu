// Tensor-core forward of the molecular-kernel convolution (sm_100a, tcgen05 + TMEM).
// Same contract as k_conv_fwd (conv_fwd.cu); replaces KernelConv.calculate_total_score (reference kernels.py:353-425)
// and the bucket gathers / output assembly of BaseKernelSetConv.forward (kernels.py:519-548, 674-747).
//
// Formulation: for the degree-d bucket every cosine is a dot product of L2-normalised rows, so the d x d similarity
// tile of every (node, kernel) pair is one entry block of the GEMM
//        T[(n,j), (k,s)] = xhat[nei(n,j), :] . shat[k, s, :]          (M = gathered neighbour rows, N = kernel rows, K = F)
// and the centre term C[n,k] = xhat[n,:] . chat[k,:] is a second GEMM on the focal rows.  fp32 accuracy comes from
// splitting every operand into two fp16 numbers, v = hi + lo * 2^-11 (|error| <= 2^-24 |v|, |v| <= 1 after
// normalisation), and running three fp16 UMMAs per K step into two fp32 TMEM accumulators:
//        D1 += A_hi B_hi        D2 += A_lo B_hi + A_hi B_lo        T = D1 + 2^-11 D2       (dropped term ~2^-22).
//
// One persistent CTA per SM (16 warps).  The fp16 images of the current degree's kernel set (written by
// k_param_pack_tc) stay resident in shared memory; a dynamic queue hands out units of ~128 nodes (degree 4 first).
// A unit is a list of jobs: one centre job (128 focal rows), then per tile of 4*floor(32/d) nodes one support job per
// kernel range (<= 128 accumulator columns).  Jobs are software pipelined over two TMEM buffers: while the tensor
// core runs job i+1, all warps run the epilogue of job i: TMEM -> registers, the d rows of a node are exchanged
// between the d adjacent lanes that hold them (shuffle for d = 2, shared-memory bounce for d = 3, 4), then exactly
// the reference arithmetic on the d x d tile: sequential mean per permutation, first-max arg-max (kernels.py:373),
// bond cosine at the arg-max permutation (kernels.py:382-390), chirality (kernels.py:279-350), softmax mix
// (kernels.py:402-425).
#include <algorithm>
#include <stdlib.h>
#include "common.cuh"
#include "tc.cuh"

namespace mk {

constexpr int TCF_THREADS = 512;
constexpr int TCF_WARPS = TCF_THREADS / 32;
constexpr int TCF_NG = TCF_WARPS / 4;      // warps sharing one TMEM lane quadrant split the accumulator columns
constexpr int TCF_SCR = 2048;              // exchange scratch per warp (bytes)

template <int D> struct TcGeo {
    static constexpr int DS = D == 3 ? 4 : D;          // accumulator columns per kernel
    static constexpr int NPW = 32 / D;                 // nodes per warp quadrant (32 TMEM lanes)
    static constexpr int TN = 4 * NPW;                 // nodes per support tile: 128, 64, 40, 32
    static constexpr int TPU = D;                      // support tiles per unit
    static constexpr int UN = TN * TPU;                // nodes per unit: 128, 128, 120, 128
    static constexpr int KC = D == 1 ? 4 : D;          // kernels per epilogue chunk
    static constexpr int CW = D <= 2 ? 4 : 16;         // accumulator columns loaded per chunk
};

struct FwdTcArgs {
    const float* x; const float* xnorm; int ldx;
    int F, Fp, Fk;
    const int* sel; const int* nei; const float* ehat; const int8_t* tsign;
    int n[4], boff[4], eoff[4], L[4], koff[4];
    const float* packed[4];
    int unit_begin[5];      // queue position q = 0..3 <-> degree 4-q
    int KR[4], nr[4];       // kernels per range, number of ranges
    int is_last;
    float* sc; int sc_mode; int ld_sc; long long scoff[4];
    uint8_t* argmax; uint8_t* argmax_free; const uint8_t* argmax_in;
    int* counter;
    int sm_Ahi, sm_Alo, sm_Bhi, sm_Blo, sm_Chi, sm_Clo, sm_scr, sm_dup;   // byte offsets
    int sm_ES[4];           // normalised support bond rows of every degree (resident for the whole kernel)
};

struct TcJob {
    int valid, d, u0, kind, tile, range, newA, buf;   // kind 0 = centre, 1 = support
};

struct TcIter {      // uniform across the CTA
    int d, u0, step, nsteps, unit_valid, njobs;
};

__device__ __forceinline__ int tc_tpu(int d) { return d; }
__device__ __forceinline__ int tc_tn(int d) { return 4 * (32 / d); }

// next job of the CTA's stream; fetches a new unit from the global queue when the current one is exhausted
__device__ __forceinline__ TcJob tc_next_job(const FwdTcArgs& a, TcIter& it, int* s_unit) {
    TcJob j;
    j.valid = 0; j.d = 1; j.u0 = 0; j.kind = 0; j.tile = 0; j.range = 0; j.newA = 0; j.buf = 0;
    while (true) {
        if (!it.unit_valid) {
            __syncthreads();
            if (threadIdx.x == 0) *s_unit = atomicAdd(a.counter, 1);
            __syncthreads();
            const int u = *s_unit;
            if (u >= a.unit_begin[4]) return j;
            int q = 0;
            while (u >= a.unit_begin[q + 1]) ++q;
            it.d = 4 - q;
            it.u0 = (u - a.unit_begin[q]) * tc_tn(it.d) * tc_tpu(it.d);
            it.step = 0;
            it.nsteps = 1 + tc_tpu(it.d) * a.nr[it.d - 1];
            it.unit_valid = 1;
        }
        if (it.step >= it.nsteps) { it.unit_valid = 0; continue; }
        const int st = it.step++;
        const int nr = a.nr[it.d - 1];
        int kind = 0, tile = 0, range = 0;
        if (st > 0) { kind = 1; tile = (st - 1) / nr; range = (st - 1) % nr; }
        if (kind == 1 && it.u0 + tile * tc_tn(it.d) >= a.n[it.d - 1]) { it.unit_valid = 0; continue; }   // past the bucket end
        j.valid = 1; j.d = it.d; j.u0 = it.u0; j.kind = kind; j.tile = tile; j.range = range;
        j.newA = (kind == 0) || (range == 0);
        j.buf = it.njobs & 1;
        ++it.njobs;
        return j;
    }
}

__device__ __forceinline__ void copy16(unsigned char* dst, const unsigned char* src, int64_t bytes) {
    for (int64_t i = (int64_t)threadIdx.x * 16; i < bytes; i += (int64_t)TCF_THREADS * 16)
        *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(src + i);
}

// ---- operand staging ---------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ void tc_stage_B(const FwdTcArgs& a, unsigned char* smem) {
    const int L = a.L[D - 1];
    const PackedLayout pl(D, L, a.Fp);
    const float* pk = a.packed[D - 1];
    const unsigned char* img = reinterpret_cast<const unsigned char*>(pk + pl.tc);
    copy16(smem + a.sm_Bhi, img + pl.tc_sup_hi(), pl.tc_sup_bytes);
    copy16(smem + a.sm_Blo, img + pl.tc_sup_lo(), pl.tc_sup_bytes);
    copy16(smem + a.sm_Chi, img + pl.tc_cen_hi(), pl.tc_cen_bytes);
    copy16(smem + a.sm_Clo, img + pl.tc_cen_lo(), pl.tc_cen_bytes);
}

// 128 A rows: support job -> row (quadrant q, lane l) = neighbour j = l % D of node slot q*NPW + l / D of the tile;
// centre job -> row r = focal node r of the unit.  Rows are normalised, split into fp16 (hi, lo) and written in the
// interleaved layout: 8 consecutive rows x 16 bytes are contiguous, so a warp's stores are conflict free.
template <int D>
__device__ __forceinline__ void tc_gather_A(const FwdTcArgs& a, unsigned char* smem, const TcJob& job) {
    using G = TcGeo<D>;
    const int tid = threadIdx.x;
    const int r = tid & 127, cg = tid >> 7;
    const int n = a.n[D - 1];
    int node = -1;
    if (job.kind == 0) {
        const int R = job.u0 + r;
        if (r < G::UN && R < n) node = a.sel[a.boff[D - 1] + R];
    } else {
        const int q = r >> 5, l = r & 31;
        const int i = l / D, j = l - i * D;
        const int R = job.u0 + job.tile * G::TN + q * G::NPW + i;
        if (l < G::NPW * D && R < n) node = a.nei[(size_t)a.eoff[D - 1] + (size_t)R * D + j];
    }
    float rinv = 0.f;
    const float* xr = a.x;
    if (node >= 0) {
        rinv = 1.0f / fmaxf(a.xnorm[node], MOLKGNN_COS_EPS);
        xr = a.x + (size_t)node * a.ldx;
    }
    const int nch = a.Fk >> 3;
    unsigned char* Ahi = smem + a.sm_Ahi;
    unsigned char* Alo = smem + a.sm_Alo;
    constexpr int UNR = 4;                       // chunks in flight per thread (8 x LDG.128)
    for (int c0 = cg; c0 < nch; c0 += 4 * UNR) {
        float4 v[UNR][2];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int c = c0 + 4 * u;
            v[u][0] = v[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (node >= 0 && c < nch) {
                if (8 * c + 4 <= a.Fp) v[u][0] = ld4(xr + 8 * c);
                if (8 * c + 8 <= a.Fp) v[u][1] = ld4(xr + 8 * c + 4);
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int c = c0 + 4 * u;
            if (c < nch) {
                const float f[8] = {v[u][0].x * rinv, v[u][0].y * rinv, v[u][0].z * rinv, v[u][0].w * rinv,
                                    v[u][1].x * rinv, v[u][1].y * rinv, v[u][1].z * rinv, v[u][1].w * rinv};
                __align__(16) __half2 hi[4];
                __align__(16) __half2 lo[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    hi[t] = __floats2half2_rn(f[2 * t], f[2 * t + 1]);
                    const float2 hf = __half22float2(hi[t]);
                    lo[t] = __floats2half2_rn((f[2 * t] - hf.x) * tc::LO_SCALE, (f[2 * t + 1] - hf.y) * tc::LO_SCALE);
                }
                const uint32_t off = tc::il_off(r, 8 * c, a.Fk);
                *reinterpret_cast<uint4*>(Ahi + off) = *reinterpret_cast<const uint4*>(hi);
                *reinterpret_cast<uint4*>(Alo + off) = *reinterpret_cast<const uint4*>(lo);
            }
        }
    }
}

// chirality gate of a degree-4 tile: any two of the four neighbour feature rows bit-equal (torch.equal, kernels.py:310-317)
__device__ __forceinline__ void tc_dup_flags(const FwdTcArgs& a, unsigned char* dupf, const TcJob& job) {
    using G = TcGeo<4>;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = a.n[3];
    for (int nl = warp; nl < G::TN; nl += TCF_WARPS) {
        const int R = job.u0 + job.tile * G::TN + nl;
        bool dup = false;
        if (R < n) {
            int u[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) u[j] = a.nei[(size_t)a.eoff[3] + (size_t)R * 4 + j];
            unsigned neq = 0;
            for (int f = lane; f < a.F; f += 32) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = a.x[(size_t)u[j] * a.ldx + f];
                int b = 0;
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int q = p + 1; q < 4; ++q, ++b) if (!(v[p] == v[q])) neq |= 1u << b;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) neq |= __shfl_xor_sync(0xffffffffu, neq, o);
            dup = (neq != 0x3fu);
        }
        if (lane == 0) dupf[nl] = dup ? 1 : 0;
    }
}

template <int D>
__device__ __forceinline__ void tc_prepare(const FwdTcArgs& a, unsigned char* smem, const TcJob& job, int& staged_deg) {
    if (staged_deg != D) { tc_stage_B<D>(a, smem); staged_deg = D; }
    if (job.newA) {
        tc_gather_A<D>(a, smem, job);
        if (D == 4 && a.is_last && job.kind == 1) tc_dup_flags(a, smem + a.sm_dup + (job.tile & 1) * 32, job);
    }
}

// issued by one thread: 3 fp16 UMMAs per K step into (D1, D2) of the job's TMEM buffer
template <int D>
__device__ __forceinline__ void tc_issue(const FwdTcArgs& a, unsigned char* smem, const TcJob& job, uint32_t tmem, uint64_t* bars) {
    using G = TcGeo<D>;
    const int L = a.L[D - 1];
    const uint32_t sbo = (uint32_t)(a.Fk >> 3) * 128u;
    uint32_t bhi, blo;
    int N;
    if (job.kind == 0) {
        N = (L + 15) / 16 * 16;
        bhi = tc::smem_u32(smem + a.sm_Chi);
        blo = tc::smem_u32(smem + a.sm_Clo);
    } else {
        const int k0 = job.range * a.KR[D - 1];
        const int kn = min(a.KR[D - 1], L - k0);
        N = (kn * G::DS + 15) / 16 * 16;
        const uint32_t roff = (uint32_t)(k0 * G::DS / 8) * sbo;
        bhi = tc::smem_u32(smem + a.sm_Bhi) + roff;
        blo = tc::smem_u32(smem + a.sm_Blo) + roff;
    }
    const uint32_t ahi = tc::smem_u32(smem + a.sm_Ahi), alo = tc::smem_u32(smem + a.sm_Alo);
    const uint32_t idesc = tc::idesc_f16(128, N, 0, 0);
    const uint32_t d1 = tmem + (uint32_t)job.buf * 256u, d2 = d1 + 128u;
    const int nks = a.Fk >> 4;
    for (int ks = 0; ks < nks; ++ks) {
        const uint32_t o = (uint32_t)ks * 256u;
        const uint64_t dAh = tc::smem_desc(ahi + o, 128u, sbo), dAl = tc::smem_desc(alo + o, 128u, sbo);
        const uint64_t dBh = tc::smem_desc(bhi + o, 128u, sbo), dBl = tc::smem_desc(blo + o, 128u, sbo);
        tc::umma_f16(d1, dAh, dBh, idesc, ks > 0 ? 1u : 0u);
        tc::umma_f16(d2, dAl, dBh, idesc, ks > 0 ? 1u : 0u);
        tc::umma_f16(d2, dAh, dBl, idesc, 1u);
    }
    tc::umma_commit(&bars[job.buf]);
}

// ---- epilogues ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t tc_sc_index(const FwdTcArgs& a, int d, int R, int L, int k, int focal) {
    return a.sc_mode == 0 ? (size_t)a.scoff[d - 1] + (size_t)R * L + k : (size_t)focal * a.ld_sc + a.koff[d - 1] + k;
}

// centre job: C[n,k] parked in the score buffer (the support epilogue of the same CTA reads it back)
template <int D>
__device__ __forceinline__ void tc_epilogue_centre(const FwdTcArgs& a, const TcJob& job, uint32_t tmem) {
    using G = TcGeo<D>;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = warp & 3, g = warp >> 2;
    const int L = a.L[D - 1], n = a.n[D - 1];
    const int r = q * 32 + lane;
    const int R = job.u0 + r;
    const bool ok = r < G::UN && R < n;
    int focal = 0;
    if (ok && a.sc_mode != 0) focal = a.sel[a.boff[D - 1] + R];
    const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)job.buf * 256u;
    for (int c = g; c * 16 < L; c += TCF_NG) {
        uint32_t v1[16], v2[16];
        tc::tmem_ld16(tbase + c * 16, v1);
        tc::tmem_ld16(tbase + 128 + c * 16, v2);
        tc::tmem_ld_wait();
        if (ok) {
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const int k = c * 16 + t;
                if (k < L)
                    a.sc[tc_sc_index(a, D, R, L, k, focal)] = fmaf(__uint_as_float(v2[t]), tc::LO_UNSCALE, __uint_as_float(v1[t]));
            }
        }
    }
}

// exchange-scratch unit (16 B) of (node slot i, chunk kernel kk, neighbour row j)
template <int D> __device__ __forceinline__ int tc_scr_unit(int i, int kk, int j) {
    if (D == 4) return i * 16 + (kk >> 1) * 8 + (((kk & 1) ^ (i & 1)) << 2) + ((j ^ kk) & 3);   // conflict-free both ways
    return (i * D + kk) * 4 + j;
}

// the reference arithmetic on one (node, kernel) pair given its d x d similarity tile T[j][s] and centre dot
template <int D>
__device__ __forceinline__ void tc_score_pair(const FwdTcArgs& a, const float (&T)[D][D], float cdot, int R, int k, int L,
                                              int focal, const float* ESs, const unsigned char* dupf, int nl,
                                              float ws, float wc, float we, float W, float rW, const int8_t* supsign) {
    constexpr int P = Perm<D>::P;
    const size_t cidx = (size_t)a.scoff[D - 1] + (size_t)R * L + k;
    const int forced = a.argmax_in ? (a.argmax_in[cidx] & 0x7f) : -1;
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
        if (p == forced) used = s;
    }
    if (a.argmax_free) a.argmax_free[cidx] = (uint8_t)bi;
    if (forced >= 0 && forced < P) { bi = forced; best = used; }
    // bond-attribute cosine at the chosen permutation (kernels.py:382-390)
    uint32_t code = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) if (p == bi) code = perm_code<D>(p);
    const float* en = a.ehat + ((size_t)a.eoff[D - 1] + (size_t)R * D) * EP;
    float esum = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const int s = (code >> (2 * j)) & 3;
        const float* es = ESs + (size_t)(s * L + k) * EP;
        const float4 e0 = __ldg(reinterpret_cast<const float4*>(en + j * EP)), e1 = __ldg(reinterpret_cast<const float4*>(en + j * EP + 4));
        const float4 s0 = ld4(es), s1 = ld4(es + 4);
        float dd = 0.f;
        dd = fmaf(e0.x, s0.x, dd); dd = fmaf(e0.y, s0.y, dd); dd = fmaf(e0.z, s0.z, dd); dd = fmaf(e0.w, s0.w, dd);
        dd = fmaf(e1.x, s1.x, dd); dd = fmaf(e1.y, s1.y, dd); dd = fmaf(e1.z, s1.z, dd); dd = fmaf(e1.w, s1.w, dd);
        esum = j == 0 ? dd : esum + dd;
    }
    const float E = div_deg<D>(esum);
    float sc = div_by((best * ws + cdot * wc) + E * we, W, rW);
    uint8_t am = (uint8_t)bi;
    if (D == 4 && a.is_last) {
        // chirality (kernels.py:279-350): +1 if any two neighbours are identical, else sign agreement
        int chi = 1;
        if (!dupf[nl]) chi = (a.tsign[R] == supsign[k * 12 + bi]) ? 1 : -1;
        if (chi < 0) { sc = -sc; am |= 0x80; }
    }
    a.argmax[cidx] = am;
    a.sc[tc_sc_index(a, D, R, L, k, focal)] = sc;
}

template <int D>
__device__ __forceinline__ void tc_epilogue_support(const FwdTcArgs& a, unsigned char* smem, const TcJob& job, uint32_t tmem) {
    using G = TcGeo<D>;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = warp & 3, g = warp >> 2;
    const int L = a.L[D - 1], n = a.n[D - 1];
    const int i = lane / D, j = lane - i * D;
    const int nl = q * G::NPW + i;                        // node slot inside the tile
    const int R = job.u0 + job.tile * G::TN + nl;
    const bool node_ok = (lane < G::NPW * D) && R < n;
    const int k0 = job.range * a.KR[D - 1];
    const int kend = min(L, k0 + a.KR[D - 1]);
    const int nch = (kend - k0 + G::KC - 1) / G::KC;
    const PackedLayout pl(D, L, a.Fp);
    const float* pk = a.packed[D - 1];
    const float ws = pk[pl.w + 0], wc = pk[pl.w + 1], we = pk[pl.w + 2], W = pk[pl.w + 3];
    const float rW = 1.0f / W;
    const int8_t* supsign = reinterpret_cast<const int8_t*>(pk + pl.sign);
    const float* ESs = reinterpret_cast<const float*>(smem + a.sm_ES[D - 1]);
    const unsigned char* dupf = smem + a.sm_dup + (job.tile & 1) * 32;
    float4* scr = reinterpret_cast<float4*>(smem + a.sm_scr + warp * TCF_SCR);
    int focal = 0;
    if (node_ok && a.sc_mode != 0) focal = a.sel[a.boff[D - 1] + R];
    const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)job.buf * 256u;

    for (int c = g; c < nch; c += TCF_NG) {
        float v[G::CW];
        {
            uint32_t v1[G::CW], v2[G::CW];
            const uint32_t col = (uint32_t)(c * G::KC * G::DS);
            if (G::CW == 4) { tc::tmem_ld4(tbase + col, v1); tc::tmem_ld4(tbase + 128 + col, v2); }
            else { tc::tmem_ld16(tbase + col, v1); tc::tmem_ld16(tbase + 128 + col, v2); }
            tc::tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < G::CW; ++t) v[t] = fmaf(__uint_as_float(v2[t]), tc::LO_UNSCALE, __uint_as_float(v1[t]));
        }
        if (D == 1) {
            // lane = node: four kernels per chunk, no exchange
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int k = k0 + c * 4 + kk;
                if (node_ok && k < kend) {
                    float T[D][D];
                    T[0][0] = v[kk];
                    const float cdot = __ldcg(a.sc + tc_sc_index(a, D, R, L, k, focal));
                    tc_score_pair<D>(a, T, cdot, R, k, L, focal, ESs, dupf, nl, ws, wc, we, W, rW, supsign);
                }
            }
        } else {
            float T[D][D];
            if (D == 2) {
                // lane (i, j) keeps its own row for kernel kk = j and swaps the other kernel's row with its partner
                const float m0 = j == 0 ? v[0] : v[2], m1 = j == 0 ? v[1] : v[3];
                const float s0 = j == 0 ? v[2] : v[0], s1 = j == 0 ? v[3] : v[1];
                const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
                T[0][0] = j == 0 ? m0 : r0; T[0][1] = j == 0 ? m1 : r1;
                T[1][0] = j == 0 ? r0 : m0; T[1][1] = j == 0 ? r1 : m1;
            } else {
                __syncwarp();
                if (lane < G::NPW * D) {
#pragma unroll
                    for (int kk = 0; kk < G::KC; ++kk)
                        scr[tc_scr_unit<D>(i, kk, j)] = make_float4(v[kk * 4 + 0], v[kk * 4 + 1], v[kk * 4 + 2], v[kk * 4 + 3]);
                }
                __syncwarp();
                if (lane < G::NPW * D) {
#pragma unroll
                    for (int jj = 0; jj < D; ++jj) {
                        const float4 t = scr[tc_scr_unit<D>(i, j, jj)];
                        T[jj][0] = t.x; T[jj][1] = t.y; T[jj][2] = t.z;
                        if (D == 4) T[jj][D - 1] = t.w;
                    }
                }
            }
            const int k = k0 + c * G::KC + j;
            if (node_ok && k < kend) {
                const float cdot = __ldcg(a.sc + tc_sc_index(a, D, R, L, k, focal));
                tc_score_pair<D>(a, T, cdot, R, k, L, focal, ESs, dupf, nl, ws, wc, we, W, rW, supsign);
            }
        }
    }
}

template <int D>
__device__ __forceinline__ void tc_epilogue(const FwdTcArgs& a, unsigned char* smem, const TcJob& job, uint32_t tmem) {
    if (job.kind == 0) tc_epilogue_centre<D>(a, job, tmem);
    else tc_epilogue_support<D>(a, smem, job, tmem);
}

__global__ void __launch_bounds__(TCF_THREADS, 1) k_conv_fwd_tc(const __grid_constant__ FwdTcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tslot;
    __shared__ int s_unit;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { tc::mbar_init(&bars[0], 1); tc::mbar_init(&bars[1], 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&tslot, 512);
    for (int d = 1; d <= 4; ++d) {
        const int L = a.L[d - 1];
        if (L == 0) continue;
        const PackedLayout pl(d, L, a.Fp);
        copy16(smem + a.sm_ES[d - 1], reinterpret_cast<const unsigned char*>(a.packed[d - 1] + pl.es), (int64_t)d * L * EP * 4);
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tslot;

    TcIter it;
    it.d = 1; it.u0 = 0; it.step = 0; it.nsteps = 0; it.unit_valid = 0; it.njobs = 0;
    int staged_deg = 0;
    uint32_t phase[2] = {0u, 0u};

    auto prepare_and_issue = [&](const TcJob& job) {
        switch (job.d) {
            case 1: tc_prepare<1>(a, smem, job, staged_deg); break;
            case 2: tc_prepare<2>(a, smem, job, staged_deg); break;
            case 3: tc_prepare<3>(a, smem, job, staged_deg); break;
            default: tc_prepare<4>(a, smem, job, staged_deg); break;
        }
        tc::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            switch (job.d) {
                case 1: tc_issue<1>(a, smem, job, tmem, bars); break;
                case 2: tc_issue<2>(a, smem, job, tmem, bars); break;
                case 3: tc_issue<3>(a, smem, job, tmem, bars); break;
                default: tc_issue<4>(a, smem, job, tmem, bars); break;
            }
        }
    };

    TcJob cur = tc_next_job(a, it, &s_unit);
    if (cur.valid) prepare_and_issue(cur);
    while (cur.valid) {
        TcJob nxt = tc_next_job(a, it, &s_unit);
        if (nxt.valid) {
            // the tensor core still reads A / B of `cur`: wait for it before overwriting either
            if (nxt.newA || nxt.d != staged_deg) tc::mbar_wait(&bars[cur.buf], phase[cur.buf]);
            prepare_and_issue(nxt);
        }
        tc::mbar_wait(&bars[cur.buf], phase[cur.buf]);
        phase[cur.buf] ^= 1u;
        tc::fence_after_sync();
        switch (cur.d) {
            case 1: tc_epilogue<1>(a, smem, cur, tmem); break;
            case 2: tc_epilogue<2>(a, smem, cur, tmem); break;
            case 3: tc_epilogue<3>(a, smem, cur, tmem); break;
            default: tc_epilogue<4>(a, smem, cur, tmem); break;
        }
        tc::fence_before_sync();
        __syncthreads();
        cur = nxt;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ---- host side -----------------------------------------------------------------------------------------------------
static int tcf_ds(int d) { return d == 3 ? 4 : d; }
static int tcf_un(int d) { return 4 * (32 / d) * d; }

// shared-memory plan; returns bytes or -1 if the layer does not fit the resident-kernel-set design
static int64_t tcf_configure(const molkgnn_layer_t* layer, int budget, FwdTcArgs* a) {
    const int Fk = (layer->Fp + 15) / 16 * 16;
    if (Fk > 256) return -1;
    int64_t supb = 0, cenb = 0;
    int64_t esb[4] = {0, 0, 0, 0};
    for (int d = 1; d <= 4; ++d) {
        const int L = layer->L[d - 1];
        if (L == 0) { a->KR[d - 1] = 1; a->nr[d - 1] = 0; continue; }
        if (L > 128) return -1;
        const PackedLayout pl(d, L, layer->Fp);
        supb = std::max(supb, pl.tc_sup_bytes);
        cenb = std::max(cenb, pl.tc_cen_bytes);
        esb[d - 1] = ((int64_t)d * L * EP * 4 + 127) / 128 * 128;
        const int ds = tcf_ds(d);
        // degree 3: a chunk of 3 kernels is loaded as 16 accumulator columns (12 used): the last chunk must end inside the 128
        // columns of the job, i.e. at most 10 chunks = 30 kernels per range
        const int krmax = d == 3 ? 30 : 128 / ds;
        const int align = d == 1 ? 8 : d == 2 ? 4 : 2;          // range starts on an 8-row boundary of the image
        const int nr = (L + krmax - 1) / krmax;
        int kr = ((L + nr - 1) / nr + align - 1) / align * align;
        kr = std::min(kr, krmax / align * align);
        a->KR[d - 1] = kr;
        a->nr[d - 1] = (L + kr - 1) / kr;
    }
    const int64_t abytes = (int64_t)16 * (Fk / 8) * 128;
    int64_t off = 0;
    a->sm_Ahi = (int)off; off += abytes;
    a->sm_Alo = (int)off; off += abytes;
    a->sm_Bhi = (int)off; off += supb;
    a->sm_Blo = (int)off; off += supb;
    a->sm_Chi = (int)off; off += cenb;
    a->sm_Clo = (int)off; off += cenb;
    for (int d = 0; d < 4; ++d) { a->sm_ES[d] = (int)off; off += esb[d]; }
    a->sm_scr = (int)off; off += (int64_t)TCF_WARPS * TCF_SCR;
    a->sm_dup = (int)off; off += 128;
    a->Fk = Fk;
    return off <= budget ? off : -1;
}

// returns 1 if launched, 0 if the layer is not eligible (caller falls back to the SIMT kernel), <0 on error
int launch_conv_fwd_tc(const molkgnn_plan_t* plan, const molkgnn_layer_t* layer, const float* x, int32_t ldx,
                       const float* xnorm, int32_t is_last_layer, float* sc, int32_t sc_mode, int32_t ld_sc,
                       const int64_t scoff[4], uint8_t* argmax, uint8_t* argmax_free, const uint8_t* argmax_in,
                       int32_t* counter, cudaStream_t st) {
    static int s_budget = 0, s_sms = 0;
    if (!s_budget) {
        s_budget = device_max_smem_optin();
        s_sms = device_num_sms();
        MK_REQUIRE(s_budget > 0 && s_sms > 0, "conv_fwd_tc: no CUDA device");
    }
    FwdTcArgs a;
    const int64_t smem = tcf_configure(layer, s_budget - 1024, &a);
    if (smem < 0) return 0;
    a.x = x; a.xnorm = xnorm; a.ldx = ldx;
    a.F = layer->F; a.Fp = layer->Fp;
    a.sel = plan->sel; a.nei = plan->nei; a.ehat = plan->ehat; a.tsign = plan->tsign;
    int ub = 0;
    for (int q = 0; q < 4; ++q) {
        const int d = 4 - q;
        a.unit_begin[q] = ub;
        if (plan->n[d - 1] > 0 && layer->L[d - 1] > 0) ub += (plan->n[d - 1] + tcf_un(d) - 1) / tcf_un(d);
    }
    a.unit_begin[4] = ub;
    for (int d = 0; d < 4; ++d) {
        a.n[d] = plan->n[d]; a.boff[d] = plan->boff[d]; a.eoff[d] = plan->eoff[d];
        a.L[d] = layer->L[d]; a.koff[d] = layer->koff[d];
        a.packed[d] = layer->packed[d];
        a.scoff[d] = scoff[d];
    }
    a.is_last = is_last_layer;
    a.sc = sc; a.sc_mode = sc_mode; a.ld_sc = ld_sc;
    a.argmax = argmax; a.argmax_free = argmax_free; a.argmax_in = argmax_in;
    a.counter = counter;
    if (ub == 0) return 1;
    MK_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
    static int64_t s_attr_dev[16] = {0};
    int64_t& s_attr = s_attr_dev[device_index()];      // function attributes are per device
    if (smem > s_attr) {
        MK_CHECK_CUDA(cudaFuncSetAttribute(k_conv_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        s_attr = smem;
    }
    const int grid = std::min(ub, s_sms);
    count_launches(1);
    k_conv_fwd_tc<<<grid, TCF_THREADS, smem, st>>>(a);
    MK_CHECK_CUDA(cudaGetLastError());
    return 1;
}

}  // namespace mk
