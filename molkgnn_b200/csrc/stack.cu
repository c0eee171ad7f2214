// The whole conv stack behind two host calls: MolGCN.forward (reference KernelLayer.py:107-120: per layer
// `sim_sc = layer(...)`, `h = propagate(edge_index, sim_sc)`) and the backward autograd derives from it.  Pure host code:
// carves ONE caller-owned workspace and issues the launches of the per-layer entry points (conv_fwd.cu, activations.cu,
// conv_bwd.cu, params.cu) back to back, so that a training step costs three host calls (bucket pass, forward, backward)
// instead of ~45 calls and ~60 allocations driven from Python.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "tile.cuh"

using namespace mk;

namespace mk {
int launch_stack_fwd_fused(const molkgnn_plan_t* plan, const molkgnn_layer_t* layers, int nl, const float* x, int32_t ldx,
                           void* const* ximg, float* const* hnorm, uint8_t* const* amT, float* hgate, int32_t ld_hgate,
                           float* h_out, int32_t ldh, float* const* sc, uint8_t* const* argmax, uint8_t* const* argmax_free,
                           const uint8_t* const* argmax_in, const int64_t (*scoff)[4], cudaStream_t st);
extern long long g_path_counts[4];
bool tile_layer_ok(const molkgnn_layer_t* layer);
}

namespace {
inline int64_t al128(int64_t v) { return (v + 127) / 128 * 128; }
inline int rup4(int v) { return (v + 3) / 4 * 4; }
}  // namespace

extern "C" int molkgnn_stack_layout(const molkgnn_plan_t* plan, const molkgnn_layer_t* layers, int32_t nl, int32_t flags,
                                    molkgnn_stack_layout_t* out) {
    MK_REQUIRE(nl >= 1 && nl <= MOLKGNN_MAX_LAYERS, "stack_layout: %d layers (1..%d)", nl, MOLKGNN_MAX_LAYERS);
    memset(out, 0, sizeof(*out));
    const int64_t N = plan->N;
    int64_t off = 0;
    auto take = [&](int64_t bytes) { const int64_t o = off; off += al128(std::max<int64_t>(bytes, 16)); return o; };
    int64_t sc_max = 0;
    for (int i = 0; i < nl; ++i) {
        const molkgnn_layer_t& ly = layers[i];
        MK_REQUIRE(ly.Fp == rup4(ly.F) && ly.K > 0, "stack_layout: layer %d: Fp=%d must be roundup4(F=%d), K=%d > 0", i,
                   ly.Fp, ly.F, ly.K);
        MK_REQUIRE(i == 0 || ly.F == layers[i - 1].K, "stack_layout: layer %d: node_attr_dim %d != kernels of layer %d (%d)",
                   i, ly.F, i - 1, layers[i - 1].K);
        int64_t tot = 0;
        for (int d = 0; d < 4; ++d) { out->scoff[i][d] = tot; tot += (int64_t)plan->n[d] * ly.L[d]; }
        out->sc_elems[i] = tot;
        sc_max = std::max(sc_max, tot);
    }
    for (int i = 0; i < nl; ++i) {
        const molkgnn_layer_t& ly = layers[i];
        out->h[i] = take(N * ly.Fp * 4);
        out->hnorm[i] = take(N * 4);
        const int64_t one = molkgnn_tile_ximg_bytes(plan, &ly);
        out->ximg[i] = (one > 0 && ly.tile_img) ? take((int64_t)plan->n_tiles * one) : -1;
        out->argmax[i] = take(out->sc_elems[i]);
        out->argmax_free[i] = (flags & MOLKGNN_STACK_WANT_FREE) ? take(out->sc_elems[i]) : -1;
        const int64_t amt = out->ximg[i] >= 0 ? molkgnn_tile_argmax_bytes(plan, &ly) : 0;
        out->argmax_tile[i] = amt > 0 ? take(amt) : -1;
        if (flags & MOLKGNN_STACK_KEEP_SC) out->sc[i] = take(out->sc_elems[i] * 4);
    }
    out->hnorm[nl] = take(N * 4);
    if (!(flags & MOLKGNN_STACK_KEEP_SC)) {
        const int64_t o = take(sc_max * 4);
        for (int i = 0; i < nl; ++i) out->sc[i] = o;
    }
    out->counter = take(MOLKGNN_MAX_LAYERS * 8 * 4);
    out->fwd_bytes = off;
    // backward scratch
    off = 0;
    int64_t part_max = 0, scr_max = 0, gx_max = 0, coef_max = 0;
    for (int i = 0; i < nl; ++i) {
        const int64_t pf = molkgnn_conv_bwd_partial_floats(plan, &layers[i]);
        MK_REQUIRE(pf >= 0, "stack_layout: conv_bwd_partial_floats failed");
        part_max = std::max(part_max, pf);
        coef_max = std::max(coef_max, molkgnn_conv_bwd_coef_floats(plan, &layers[i]));
        scr_max = std::max(scr_max, N * (int64_t)tile_fk(layers[i].Fp));
        if (i > 0) gx_max = std::max(gx_max, N * (int64_t)layers[i].Fp);
    }
    out->coef = take(coef_max * 4);
    out->partials = take(part_max * 4);
    out->partials_alt = take(part_max * 4);
    out->scratch = take(scr_max * 4);
    out->gx[0] = take(gx_max * 4);
    out->gx[1] = take(gx_max * 4);
    out->bwd_bytes = off;
    // flat parameter gradients: reference layouts, degree by degree
    int64_t g = 0;
    for (int i = 0; i < nl; ++i) {
        const molkgnn_layer_t& ly = layers[i];
        for (int d = 0; d < 4; ++d) {
            const int64_t L = ly.L[d];
            out->g_x_center[i][d] = g; g += L * ly.F;
            out->g_x_support[i][d] = g; g += L * (d + 1) * ly.F;
            out->g_edge_attr_support[i][d] = g; g += L * (d + 1) * ly.Fe;
            out->g_w[i][d] = g; g += L > 0 ? 4 : 0;
            g = (g + 3) / 4 * 4;
        }
    }
    out->grad_floats = g;
    return 0;
}

extern "C" int molkgnn_stack_fwd(const molkgnn_plan_t* plan, const molkgnn_layer_t* layers, int32_t nl,
                                 const molkgnn_stack_layout_t* lay, int32_t flags, const float* x, int32_t ldx,
                                 void* workspace, float* h_out, int32_t ldh, const uint8_t* const* argmax_in,
                                 int32_t* tile_fwd, void* stream) {
    MK_REQUIRE(nl >= 1 && nl <= MOLKGNN_MAX_LAYERS && workspace && h_out && x, "stack_fwd: bad arguments");
    MK_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 127) == 0, "stack_fwd: workspace must be 128-byte aligned");
    MK_REQUIRE(ldh % 4 == 0 && ldh >= layers[nl - 1].K && (reinterpret_cast<uintptr_t>(h_out) & 15) == 0,
               "stack_fwd: h_out must be 16-byte aligned with ldh %% 4 == 0 and ldh >= K (ldh=%d)", ldh);
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    const int N = plan->N;
    int rc;
    if (!(flags & MOLKGNN_STACK_PACKED) && (rc = molkgnn_param_pack_layers(layers, nl, 7, stream))) return rc;
    // ---- layer-fused path: the whole stack in ONE launch (csrc/stack_fwd_fused.cu) ----
    if (molkgnn_get_fwd_path() >= 3) {
        bool all_img = true;
        for (int i = 0; i < nl; ++i) all_img = all_img && lay->ximg[i] >= 0 && lay->argmax_tile[i] >= 0;
        if (all_img && nl <= MOLKGNN_MAX_LAYERS) {
            void* ximg[MOLKGNN_MAX_LAYERS]; float* hnorm[MOLKGNN_MAX_LAYERS]; uint8_t* amT[MOLKGNN_MAX_LAYERS];
            float* scp[MOLKGNN_MAX_LAYERS]; uint8_t* am[MOLKGNN_MAX_LAYERS]; uint8_t* amf[MOLKGNN_MAX_LAYERS];
            const uint8_t* amin[MOLKGNN_MAX_LAYERS];
            const bool aux = (flags & MOLKGNN_STACK_KEEP_SC) || (flags & MOLKGNN_STACK_WANT_FREE) || argmax_in;
            const bool train = !(flags & MOLKGNN_STACK_INFERENCE);
            for (int i = 0; i < nl; ++i) {
                ximg[i] = train ? ws + lay->ximg[i] : nullptr;
                hnorm[i] = train ? reinterpret_cast<float*>(ws + lay->hnorm[i]) : nullptr;
                amT[i] = train ? ws + lay->argmax_tile[i] : nullptr;
                scp[i] = (flags & MOLKGNN_STACK_KEEP_SC) ? reinterpret_cast<float*>(ws + lay->sc[i]) : nullptr;
                am[i] = aux ? ws + lay->argmax[i] : nullptr;
                amf[i] = lay->argmax_free[i] >= 0 ? ws + lay->argmax_free[i] : nullptr;
                amin[i] = argmax_in ? argmax_in[i] : nullptr;
            }
            float* hgate = nl > 1 ? reinterpret_cast<float*>(ws + lay->h[nl - 1]) : nullptr;
            rc = launch_stack_fwd_fused(plan, layers, nl, x, ldx, ximg, hnorm, amT, hgate, nl > 1 ? layers[nl - 1].Fp : 0, h_out, ldh,
                                        aux ? scp : nullptr, aux ? am : nullptr, aux ? amf : nullptr, argmax_in ? amin : nullptr,
                                        lay->scoff, (cudaStream_t)stream);
            if (rc < 0) return rc;
            if (rc > 0) {
                mk::g_path_counts[0] += nl;
                for (int i = 0; i < nl && tile_fwd; ++i) tile_fwd[i] = train ? 1 : 0;
                return 0;
            }
        }
    }
    MK_CHECK_CUDA(cudaMemsetAsync(ws + lay->counter, 0, sizeof(int32_t) * 8 * (size_t)nl, (cudaStream_t)stream));
    float* h = reinterpret_cast<float*>(ws + lay->h[0]);
    float* hn = reinterpret_cast<float*>(ws + lay->hnorm[0]);
    if (lay->ximg[0] >= 0 && layers[0].Fp <= 64 && mk::tile_layer_ok(&layers[0])) {      // one pass over x: padded copy, norms, tensor-core images
        if ((rc = molkgnn_tile_ximg_build_raw(plan, &layers[0], x, ldx, h, hn, ws + lay->ximg[0], stream))) return rc;
    } else {
        if ((rc = molkgnn_pad_norm(x, N, layers[0].F, ldx, h, layers[0].Fp, hn, stream))) return rc;
        if (lay->ximg[0] >= 0 &&
            (rc = molkgnn_tile_ximg_build(plan, &layers[0], h, layers[0].Fp, hn, ws + lay->ximg[0], stream)))
            return rc;
    }
    // tile queues of every layer: one memset for the whole stack, queued before the first kernel of the chain
    struct CountersZeroed {
        CountersZeroed() { mk::g_fwd_counters_zeroed = true; }
        ~CountersZeroed() { mk::g_fwd_counters_zeroed = false; }
    } counters_zeroed;
    for (int i = 0; i < nl; ++i) {
        const molkgnn_layer_t& ly = layers[i];
        const bool last = i == nl - 1;
        float* sc = reinterpret_cast<float*>(ws + lay->sc[i]);
        int64_t pc0[4], pc1[4];
        molkgnn_path_counts(pc0);
        if ((rc = molkgnn_conv_fwd(plan, &ly, h, ly.Fp, hn, last ? 1 : 0, sc, 0, 0, lay->scoff[i], ws + lay->argmax[i],
                                   lay->argmax_free[i] >= 0 ? ws + lay->argmax_free[i] : nullptr,
                                   argmax_in ? argmax_in[i] : nullptr,
                                   reinterpret_cast<int32_t*>(ws + lay->counter) + 8 * i,
                                   lay->ximg[i] >= 0 ? ws + lay->ximg[i] : nullptr,
                                   lay->argmax_tile[i] >= 0 ? ws + lay->argmax_tile[i] : nullptr, stream)))
            return rc;
        molkgnn_path_counts(pc1);
        if (tile_fwd) tile_fwd[i] = pc1[0] > pc0[0] ? 1 : 0;      // the molecule-tile forward ran: argmax_tile[i] is valid
        float* hnext = last ? h_out : reinterpret_cast<float*>(ws + lay->h[i + 1]);
        const int ldn = last ? ldh : layers[i + 1].Fp;
        float* hnn = reinterpret_cast<float*>(ws + lay->hnorm[i + 1]);
        // the propagate kernel writes the next layer's tile images on the way out when the shapes allow it
        const bool fuse_img = !last && lay->ximg[i + 1] >= 0 && ldn <= 112 && ldn == rup4(ly.K);
        if ((rc = molkgnn_propagate_fwd(plan, &ly, sc, lay->scoff[i], hnext, ldn, hnn, fuse_img ? ws + lay->ximg[i + 1] : nullptr,
                                        stream)))
            return rc;
        if (!last && lay->ximg[i + 1] >= 0 && !fuse_img &&
            (rc = molkgnn_tile_ximg_build(plan, &layers[i + 1], hnext, ldn, hnn, ws + lay->ximg[i + 1], stream)))
            return rc;
        h = hnext; hn = hnn;
    }
    return 0;
}

// Side stream of the parameter finalisation: k_param_finalize / k_theta of layer i only read layer i's partial copies, so
// they run beside the coefficient pre-pass and the tile kernel of layer i-1 (whose CTAs leave registers and shared memory
// for the small finalize CTAs) instead of between the layers.  Partial-copy buffers alternate between layers.
namespace {
struct FinSide { cudaStream_t st; cudaEvent_t ready[2], done[2]; bool ok; };
FinSide* fin_side() {
    static FinSide s_side[16];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    FinSide& s = s_side[dev];
    if (!s.ok) {
        if (cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        for (int i = 0; i < 2; ++i)
            if (cudaEventCreateWithFlags(&s.ready[i], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&s.done[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        s.ok = true;
    }
    return &s;
}
}  // namespace

extern "C" int molkgnn_stack_bwd(const molkgnn_plan_t* plan, const molkgnn_layer_t* layers, int32_t nl,
                                 const molkgnn_stack_layout_t* lay, void* workspace, void* bwd_scratch, const float* grad_h,
                                 int32_t ldg, float* grad_x, float* grad_flat, const int32_t* tile_fwd, void* stream) {
    MK_REQUIRE(nl >= 1 && nl <= MOLKGNN_MAX_LAYERS && workspace && bwd_scratch && grad_h, "stack_bwd: bad arguments");
    MK_REQUIRE((reinterpret_cast<uintptr_t>(bwd_scratch) & 127) == 0 && (reinterpret_cast<uintptr_t>(grad_h) & 15) == 0 &&
               ldg >= layers[nl - 1].K, "stack_bwd: scratch must be 128-byte, grad_h 16-byte aligned, ldg >= K (ldg=%d)", ldg);
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    unsigned char* bs = reinterpret_cast<unsigned char*>(bwd_scratch);
    const float* g = grad_h;
    int ld = ldg;
    cudaStream_t main_st = (cudaStream_t)stream;
    static int s_use_side = -1;
    if (s_use_side < 0) { const char* e = getenv("MOLKGNN_FIN_SIDE"); s_use_side = (e && e[0] == '0') ? 0 : 1; }
    FinSide* side = (grad_flat && s_use_side) ? fin_side() : nullptr;
    MK_REQUIRE(!grad_flat || !s_use_side || side, "stack_bwd: cannot create the side stream");
    bool used[2] = {false, false};
    for (int i = nl - 1; i >= 0; --i) {
        const molkgnn_layer_t& ly = layers[i];
        float* gx = i == 0 ? grad_x : reinterpret_cast<float*>(bs + lay->gx[i & 1]);
        molkgnn_layer_grads_t gr;
        memset(&gr, 0, sizeof(gr));
        if (grad_flat) {
            for (int d = 0; d < 4; ++d) {
                if (ly.L[d] <= 0) continue;
                gr.x_center[d] = grad_flat + lay->g_x_center[i][d];
                gr.x_support[d] = grad_flat + lay->g_x_support[i][d];
                gr.edge_attr_support[d] = grad_flat + lay->g_edge_attr_support[i][d];
                gr.w_support[d] = grad_flat + lay->g_w[i][d];
                gr.w_center[d] = grad_flat + lay->g_w[i][d] + 1;
                gr.w_edge[d] = grad_flat + lay->g_w[i][d] + 2;
            }
        }
        if (!gx && !grad_flat) break;          // nothing below this layer needs a gradient
        const int pb = i & 1;
        float* partials = reinterpret_cast<float*>(bs + (pb ? lay->partials_alt : lay->partials));
        const uint8_t* amt = (lay->argmax_tile[i] >= 0 && tile_fwd && tile_fwd[i]) ? ws + lay->argmax_tile[i] : nullptr;
        // this layer's kernels overwrite partial buffer pb: the finalisation that last read it (layer i + 2) must be done
        if (side && used[pb]) MK_CHECK_CUDA(cudaStreamWaitEvent(main_st, side->done[pb], 0));
        int rc = molkgnn_conv_bwd(plan, &ly, reinterpret_cast<const float*>(ws + lay->h[i]), ly.Fp,
                                  reinterpret_cast<const float*>(ws + lay->hnorm[i]), g, ld, 1, ws + lay->argmax[i],
                                  lay->scoff[i], reinterpret_cast<float*>(bs + lay->coef), partials, gx, gx ? ly.Fp : 0,
                                  grad_flat ? &gr : nullptr, 5, lay->ximg[i] >= 0 ? ws + lay->ximg[i] : nullptr,
                                  reinterpret_cast<float*>(bs + lay->scratch), amt, stream);
        if (rc) return rc;
        if (!side && grad_flat) {                         // MOLKGNN_FIN_SIDE=0: finalisation in line on the launching stream
            rc = molkgnn_conv_bwd(plan, &ly, reinterpret_cast<const float*>(ws + lay->h[i]), ly.Fp,
                                  reinterpret_cast<const float*>(ws + lay->hnorm[i]), g, ld, 1, ws + lay->argmax[i],
                                  lay->scoff[i], reinterpret_cast<float*>(bs + lay->coef), partials, gx, gx ? ly.Fp : 0,
                                  &gr, 2, lay->ximg[i] >= 0 ? ws + lay->ximg[i] : nullptr,
                                  reinterpret_cast<float*>(bs + lay->scratch), amt, stream);
            if (rc) return rc;
        }
        if (side) {                                       // parameter finalisation on the side stream
            MK_CHECK_CUDA(cudaEventRecord(side->ready[pb], main_st));
            MK_CHECK_CUDA(cudaStreamWaitEvent(side->st, side->ready[pb], 0));
            rc = molkgnn_conv_bwd(plan, &ly, reinterpret_cast<const float*>(ws + lay->h[i]), ly.Fp,
                                  reinterpret_cast<const float*>(ws + lay->hnorm[i]), g, ld, 1, ws + lay->argmax[i],
                                  lay->scoff[i], reinterpret_cast<float*>(bs + lay->coef), partials, gx, gx ? ly.Fp : 0,
                                  &gr, 2, lay->ximg[i] >= 0 ? ws + lay->ximg[i] : nullptr,
                                  reinterpret_cast<float*>(bs + lay->scratch), amt, side->st);
            if (rc) return rc;
            MK_CHECK_CUDA(cudaEventRecord(side->done[pb], side->st));
            used[pb] = true;
        }
        g = gx; ld = ly.Fp;
    }
    for (int pb = 0; pb < 2; ++pb)                        // the launching stream owns the gradients again
        if (side && used[pb]) MK_CHECK_CUDA(cudaStreamWaitEvent(main_st, side->done[pb], 0));
    return 0;
}

extern "C" int64_t molkgnn_struct_bytes(int32_t which) {
    switch (which) {
        case 0: return (int64_t)sizeof(molkgnn_plan_t);
        case 1: return (int64_t)sizeof(molkgnn_layer_t);
        case 2: return (int64_t)sizeof(molkgnn_layer_grads_t);
        case 3: return (int64_t)sizeof(molkgnn_stack_layout_t);
        default: return -1;
    }
}
