// GPU-side batch assembly from a packed molecule store: replaces the PyG DataLoader's collate of the
// reference (data.py:136-229 builds torch_geometric DataLoaders; PyG `Batch.from_data_list` concatenates every attribute
// along its `__cat_dim__` and adds `__inc__` = num_nodes to every key containing "index").
//
// Store = all molecules of a dataset packed back to back (CSR): x [sumN, F], p [sumN, P], edge_attr [sumE, Fe],
// edge_index [2, sumE] with MOLECULE-LOCAL node ids, node_ptr / edge_ptr [M_total + 1].  A batch is a list of molecule ids
// (any order, repeats allowed: WeightedRandomSampler with replacement, data.py:150-167).  Two launches:
//   k_collate_offsets   one CTA: sizes of the chosen molecules and their exclusive prefix sums (out_node_off, out_edge_off)
//   k_collate           one warp per chosen molecule: rows of x / p / edge_attr copied (coalesced, float4 where aligned),
//                       edge_index rebased to batch-global ids (local + out_node_off: PyG __inc__), batch vector, ptr.
// Pure integer / byte work: bit-exact with the PyG semantics by construction (tests/test_store.py).
#include "common.cuh"

namespace mk {

constexpr int CO_THREADS = 1024;

__global__ void __launch_bounds__(CO_THREADS) k_collate_offsets(const int64_t* __restrict__ ids, int M, int64_t M_total,
                                                                const int64_t* __restrict__ node_ptr,
                                                                const int64_t* __restrict__ edge_ptr, int64_t* out_node_off,
                                                                int64_t* out_edge_off, int* err) {
    __shared__ int64_t wn[CO_THREADS / 32], we[CO_THREADS / 32];
    __shared__ int64_t carry_n, carry_e;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { carry_n = 0; carry_e = 0; }
    __syncthreads();
    for (int base = 0; base < M; base += CO_THREADS) {
        const int i = base + tid;
        int64_t n = 0, e = 0;
        if (i < M) {
            const int64_t m = ids[i];
            if (m < 0 || m >= M_total) atomicOr(err, 1);
            else { n = node_ptr[m + 1] - node_ptr[m]; e = edge_ptr[m + 1] - edge_ptr[m]; }
        }
        int64_t sn = n, se = e;                       // inclusive warp scans
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t tn = __shfl_up_sync(0xffffffffu, sn, o), te = __shfl_up_sync(0xffffffffu, se, o);
            if (lane >= o) { sn += tn; se += te; }
        }
        if (lane == 31) { wn[warp] = sn; we[warp] = se; }
        __syncthreads();
        int64_t on = carry_n, oe = carry_e;
        for (int w = 0; w < warp; ++w) { on += wn[w]; oe += we[w]; }
        if (i < M) { out_node_off[i] = on + sn - n; out_edge_off[i] = oe + se - e; }
        __syncthreads();
        if (tid == CO_THREADS - 1) { carry_n = on + sn; carry_e = oe + se; }
        __syncthreads();
    }
    if (tid == 0) { out_node_off[M] = carry_n; out_edge_off[M] = carry_e; }
}

struct CollateArgs {
    const int64_t* ids; int M; int64_t M_total;
    const int64_t* node_ptr; const int64_t* edge_ptr;
    const float* x; int F; const float* p; int P; const float* edge_attr; int Fe;
    const int64_t* edge_index; int64_t E_total;
    const float* y; int Y;                            // per-molecule targets [M_total, Y], nullable
    const int64_t* out_node_off; const int64_t* out_edge_off;
    float* x_out; float* p_out; float* ea_out; int64_t* ei_out; int64_t E_b; int64_t* batch_out; int64_t* ptr_out; float* y_out;
};

__device__ __forceinline__ void co_copy(float* dst, const float* src, int64_t n, int lane) {
    if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
        const int64_t n4 = n >> 2;
        for (int64_t i = lane; i < n4; i += 32) reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
        for (int64_t i = (n4 << 2) + lane; i < n; i += 32) dst[i] = __ldg(src + i);
    } else {
        for (int64_t i = lane; i < n; i += 32) dst[i] = __ldg(src + i);
    }
}

__global__ void __launch_bounds__(256) k_collate(const __grid_constant__ CollateArgs a) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= a.M) return;
    const int64_t m = a.ids[i];
    if (m < 0 || m >= a.M_total) return;              // flagged by k_collate_offsets
    const int64_t n0 = a.node_ptr[m], n = a.node_ptr[m + 1] - n0;
    const int64_t e0 = a.edge_ptr[m], e = a.edge_ptr[m + 1] - e0;
    const int64_t on = a.out_node_off[i], oe = a.out_edge_off[i];
    co_copy(a.x_out + on * a.F, a.x + n0 * a.F, n * a.F, lane);
    co_copy(a.p_out + on * a.P, a.p + n0 * a.P, n * a.P, lane);
    co_copy(a.ea_out + oe * a.Fe, a.edge_attr + e0 * a.Fe, e * a.Fe, lane);
    for (int64_t k = lane; k < e; k += 32) {          // PyG __inc__: "index" keys are offset by the nodes in front of the graph
        a.ei_out[oe + k] = __ldg(a.edge_index + e0 + k) + on;
        a.ei_out[a.E_b + oe + k] = __ldg(a.edge_index + a.E_total + e0 + k) + on;
    }
    for (int64_t k = lane; k < n; k += 32) a.batch_out[on + k] = i;
    if (lane == 0) {
        a.ptr_out[i] = on;
        if (i == a.M - 1) a.ptr_out[a.M] = on + n;
    }
    if (a.y && lane < a.Y) a.y_out[(int64_t)i * a.Y + lane] = a.y[m * a.Y + lane];
}

}  // namespace mk

using namespace mk;

extern "C" int molkgnn_collate(const int64_t* ids, int32_t M, int64_t M_total, const int64_t* node_ptr, const int64_t* edge_ptr,
                               const float* x, int32_t F, const float* p, int32_t P, const float* edge_attr, int32_t Fe,
                               const int64_t* edge_index, int64_t E_total, const float* y, int32_t Y, int64_t* node_off,
                               int64_t* edge_off, float* x_out, float* p_out, float* ea_out, int64_t* ei_out, int64_t E_b,
                               int64_t* batch_out, int64_t* ptr_out, float* y_out, int32_t* err, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    MK_REQUIRE(M > 0 && ids && node_ptr && edge_ptr && node_off && edge_off && err, "collate: bad arguments");
    MK_REQUIRE(Y >= 0 && Y <= 32, "collate: at most 32 targets per molecule (got %d)", Y);
    MK_CHECK_CUDA(cudaMemsetAsync(err, 0, sizeof(int32_t), st));
    count_launches(2);
    k_collate_offsets<<<1, CO_THREADS, 0, st>>>(ids, M, M_total, node_ptr, edge_ptr, node_off, edge_off, err);
    CollateArgs a;
    a.ids = ids; a.M = M; a.M_total = M_total; a.node_ptr = node_ptr; a.edge_ptr = edge_ptr;
    a.x = x; a.F = F; a.p = p; a.P = P; a.edge_attr = edge_attr; a.Fe = Fe;
    a.edge_index = edge_index; a.E_total = E_total; a.y = y; a.Y = Y;
    a.out_node_off = node_off; a.out_edge_off = edge_off;
    a.x_out = x_out; a.p_out = p_out; a.ea_out = ea_out; a.ei_out = ei_out; a.E_b = E_b; a.batch_out = batch_out;
    a.ptr_out = ptr_out; a.y_out = y_out;
    k_collate<<<(M + 7) / 8, 256, 0, st>>>(a);
    MK_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---- global_add_pool (MolKGNNNet.py:59,144-146; PyG: out[g] = sum of the rows of graph g) as a deterministic segmented sum:
// the nodes of a graph are contiguous in a collated batch (ptr), so every output element is one thread's in-order sum -- no
// atomics (torch's index_add_ on CUDA adds in arbitrary order).  One warp per graph, lanes over the columns.
namespace mk {
__global__ void __launch_bounds__(256) k_segment_sum(const float* __restrict__ z, int C, int ldz, const int64_t* __restrict__ ptr,
                                                     int B, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= B) return;
    const int64_t n0 = ptr[g], n1 = ptr[g + 1];
    for (int c = lane; c < C; c += 32) {
        float s = 0.f;
        for (int64_t i = n0; i < n1; ++i) s += __ldg(z + i * ldz + c);
        out[(int64_t)g * C + c] = s;
    }
}
}  // namespace mk

extern "C" int molkgnn_segment_sum(const float* z, int32_t C, int32_t ldz, const int64_t* ptr, int32_t B, float* out, void* stream_) {
    MK_REQUIRE(z && ptr && out && B > 0 && C > 0 && ldz >= C, "segment_sum: bad arguments");
    count_launches(1);
    mk::k_segment_sum<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream_>>>(z, C, ldz, ptr, B, out);
    MK_CHECK_CUDA(cudaGetLastError());
    return 0;
}
