// One-shot all-reduce of the flat kernel-parameter gradient buffer over NVLink peer memory (data-parallel step).
//
// The buffer is latency sized (0.49 MB for the base model): a ring all-reduce pays 2 (W - 1) hops for it, while over
// NVSwitch every GPU can simply READ the other W - 1 copies.  One kernel per rank and step:
//   1. copy the rank's own gradients into its EXCHANGE buffer (cudaMalloc'ed once, exported to the peers with CUDA IPC),
//   2. the last CTA to finish that copy raises this rank's flag in every peer's exchange buffer (release, system scope),
//   3. every CTA waits until all W flags of this step are up in its own buffer (acquire, system scope),
//   4. sums the W copies in RANK ORDER -- the same order on every rank, so all ranks hold bitwise identical results, as after
//      an NCCL all-reduce -- and writes the sum (or mean) over the rank's own gradients.
// Two exchange slots alternate by step parity: a rank can only reach the copy of step s + 2 after every rank raised its
// flag of step s + 1, i.e. after every rank's kernel of step s (the last reader of slot s & 1) has finished.
// The spin has a time limit: a missing peer sets the error word instead of hanging the GPU (checked by the host).
#include <cstdint>
#include <cstring>
#include "common.cuh"

namespace mk {

constexpr int OS_MAX_WORLD = 16;
constexpr int OS_CTAS = 32;
constexpr int OS_THREADS = 512;
constexpr int OS_HDR_BYTES = 1024;        // flags[2][OS_MAX_WORLD] + counters + error word, padded

struct OneShot {
    int rank, world;
    int64_t cap_bytes;                    // one data slot
    unsigned char* local;                 // own exchange buffer: [header][slot 0][slot 1]
    unsigned char* peer[OS_MAX_WORLD];    // everybody's exchange buffer as seen from this process (peer[rank] == local)
    unsigned int epoch;
    bool opened;
};

struct OneShotArgs {
    unsigned char* peer[OS_MAX_WORLD];
    int rank, world;
    unsigned int epoch;
    long long n4, slot_off;
    float scale;
    float4* flat;
    unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned long long os_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void os_store_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int os_load_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 os_load_peer(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// header layout (unsigned ints): [0 .. 2 W)  flags[parity][rank];  [64] CTA counter;  [65] error word
__global__ void __launch_bounds__(OS_THREADS) k_oneshot_allreduce(const __grid_constant__ OneShotArgs a) {
    __shared__ int s_last, s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    const int par = (int)(a.epoch & 1u);
    unsigned int* hdr = reinterpret_cast<unsigned int*>(a.peer[a.rank]);
    float4* mine = reinterpret_cast<float4*>(a.peer[a.rank] + OS_HDR_BYTES + par * a.slot_off);
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // 1. own gradients -> own exchange slot
    for (long long i = i0; i < a.n4; i += stride) mine[i] = a.flat[i];
    __threadfence_system();
    __syncthreads();
    // 2. the last CTA raises this rank's flag everywhere
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(&hdr[64], 1u);
        s_last = done == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        if (threadIdx.x == 0) hdr[64] = 0u;
        __threadfence_system();
        if (threadIdx.x < a.world)
            os_store_release_sys(reinterpret_cast<unsigned int*>(a.peer[threadIdx.x]) + par * OS_MAX_WORLD + a.rank, a.epoch);
    }
    // 3. wait for every rank's flag of this step
    if (threadIdx.x < a.world) {
        const unsigned int* f = hdr + par * OS_MAX_WORLD + threadIdx.x;
        const unsigned long long t0 = os_now();
        while (os_load_acquire_sys(f) != a.epoch) {
            if (os_now() - t0 > a.timeout_ns) { atomicExch(&hdr[65], 1u + (unsigned int)threadIdx.x); s_bad = 1; break; }
            __nanosleep(100);
        }
    }
    __syncthreads();
    // A flag that never arrived: the peers' slots hold stale or partial data.  Never reduce them into the gradients --
    // poison this rank's buffer with NaN instead, so that no optimizer step can silently consume a wrong, rank-divergent
    // sum (the error word says which rank was missing; GradBucket.check() / molkgnn_oneshot_error raise on it).
    if (s_bad) {
        const float qnan = __int_as_float(0x7fc00000);
        for (long long i = i0; i < a.n4; i += stride) a.flat[i] = make_float4(qnan, qnan, qnan, qnan);
        return;
    }
    // 4. sum the W copies in rank order
    for (long long i = i0; i < a.n4; i += stride) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < a.world; ++r) {
            const float4 v = os_load_peer(reinterpret_cast<const float4*>(a.peer[r] + OS_HDR_BYTES + par * a.slot_off) + i);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        a.flat[i] = make_float4(s.x * a.scale, s.y * a.scale, s.z * a.scale, s.w * a.scale);
    }
}

}  // namespace mk

using namespace mk;

extern "C" int molkgnn_oneshot_create(int32_t rank, int32_t world, int64_t bytes, void** handle) {
    MK_REQUIRE(handle && world >= 1 && world <= OS_MAX_WORLD && rank >= 0 && rank < world && bytes > 0,
               "oneshot_create: bad arguments (rank %d of %d, %lld bytes)", rank, world, (long long)bytes);
    OneShot* h = new OneShot();
    memset(h, 0, sizeof(*h));
    h->rank = rank; h->world = world;
    h->cap_bytes = (bytes + 255) / 256 * 256;
    const size_t total = OS_HDR_BYTES + 2 * (size_t)h->cap_bytes;
    if (cudaMalloc(reinterpret_cast<void**>(&h->local), total) != cudaSuccess) {
        delete h;
        MK_REQUIRE(false, "oneshot_create: cudaMalloc of %zu bytes failed", total);
    }
    MK_CHECK_CUDA(cudaMemset(h->local, 0, total));
    MK_CHECK_CUDA(cudaDeviceSynchronize());
    h->peer[rank] = h->local;
    *handle = h;
    return 0;
}

extern "C" int molkgnn_oneshot_ipc_handle(void* handle, void* out64) {
    OneShot* h = reinterpret_cast<OneShot*>(handle);
    MK_REQUIRE(h && out64, "oneshot_ipc_handle: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t ipc;
    MK_CHECK_CUDA(cudaIpcGetMemHandle(&ipc, h->local));
    memcpy(out64, &ipc, 64);
    return 0;
}

extern "C" int molkgnn_oneshot_open(void* handle, const void* handles /* world x 64 bytes, rank order */) {
    OneShot* h = reinterpret_cast<OneShot*>(handle);
    MK_REQUIRE(h && handles && !h->opened, "oneshot_open: bad arguments");
    for (int r = 0; r < h->world; ++r) {
        if (r == h->rank) continue;
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, reinterpret_cast<const unsigned char*>(handles) + 64 * (size_t)r, 64);
        void* p = nullptr;
        MK_CHECK_CUDA(cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess));
        h->peer[r] = reinterpret_cast<unsigned char*>(p);
    }
    h->opened = true;
    return 0;
}

// In place over flat[0 .. n): sum (average != 0: mean) over the ranks.  flat must be 16-byte aligned; n is rounded up to a
// multiple of 4 floats, so the allocation behind flat must hold that many (torch allocations are 512-byte granular).
extern "C" int molkgnn_oneshot_allreduce(void* handle, float* flat, int64_t n, int32_t average, void* stream) {
    OneShot* h = reinterpret_cast<OneShot*>(handle);
    MK_REQUIRE(h && flat && n > 0 && (h->opened || h->world == 1), "oneshot_allreduce: not opened");
    MK_REQUIRE((reinterpret_cast<uintptr_t>(flat) & 15) == 0, "oneshot_allreduce: flat must be 16-byte aligned");
    const int64_t n4 = (n + 3) / 4;
    MK_REQUIRE(n4 * 16 <= h->cap_bytes, "oneshot_allreduce: %lld floats exceed the exchange slot (%lld bytes)", (long long)n,
               (long long)h->cap_bytes);
    OneShotArgs a;
    for (int r = 0; r < OS_MAX_WORLD; ++r) a.peer[r] = r < h->world ? h->peer[r] : nullptr;
    a.rank = h->rank; a.world = h->world;
    a.epoch = ++h->epoch;
    if (a.epoch == 0u) a.epoch = h->epoch = 2u;          // never 0 (the cleared flags), parity kept
    a.n4 = n4; a.slot_off = h->cap_bytes;
    a.scale = average ? 1.0f / (float)h->world : 1.0f;
    a.flat = reinterpret_cast<float4*>(flat);
    a.timeout_ns = 5ull * 1000000000ull;
    count_launches(1);
    ProfScope prof("oneshot_allreduce", (cudaStream_t)stream);
    k_oneshot_allreduce<<<OS_CTAS, OS_THREADS, 0, (cudaStream_t)stream>>>(a);
    MK_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// 0 = fine; r + 1 = the flag of rank r did not arrive within the time limit in some step since the last call (clears it).
// Synchronises the device.
extern "C" int molkgnn_oneshot_error(void* handle) {
    OneShot* h = reinterpret_cast<OneShot*>(handle);
    MK_REQUIRE(h, "oneshot_error: bad arguments");
    unsigned int e = 0;
    MK_CHECK_CUDA(cudaDeviceSynchronize());
    MK_CHECK_CUDA(cudaMemcpy(&e, h->local + 65 * 4, 4, cudaMemcpyDeviceToHost));
    if (e) MK_CHECK_CUDA(cudaMemset(h->local + 65 * 4, 0, 4));
    return (int)e;
}

// Non-blocking variant for the training loop: queues a copy of the error word into `pinned_host4` (4 bytes of page-locked host
// memory) behind the work already on `stream`; the caller reads it once the stream has passed that point (an event it records
// itself) -- no device synchronisation inside the step loop.
extern "C" int molkgnn_oneshot_error_async(void* handle, void* pinned_host4, void* stream) {
    OneShot* h = reinterpret_cast<OneShot*>(handle);
    MK_REQUIRE(h && pinned_host4, "oneshot_error_async: bad arguments");
    MK_CHECK_CUDA(cudaMemcpyAsync(pinned_host4, h->local + 65 * 4, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return 0;
}

extern "C" int molkgnn_oneshot_destroy(void* handle) {
    OneShot* h = reinterpret_cast<OneShot*>(handle);
    if (!h) return 0;
    cudaDeviceSynchronize();
    for (int r = 0; r < h->world; ++r)
        if (r != h->rank && h->peer[r]) cudaIpcCloseMemHandle(h->peer[r]);
    if (h->local) cudaFree(h->local);
    delete h;
    return 0;
}
