// Frame-shard <-> pixel-shard layout exchange and GroupNorm-statistics all-reduce over NVLink PEER MEMORY
// (multi-GPU frame sharding of one sample, videomv_b200/parallel.py; SURVEY.md section 8e option B).
//
// Every rank owns an arena that its peers have mapped through CUDA IPC.  An exchange is ONE kernel per rank:
//   1. copy: each rank stores the slices the other ranks need straight into THEIR output tensors (16 B remote stores
//      over NVLink / NVSwitch, contiguous HWl*C chunks on both sides), its own slice locally;
//   2. signal: every CTA fences (system scope) and arrives on a local counter; the last one publishes this rank's epoch
//      in every peer's flag array (st.release.sys);
//   3. wait: the same thread spins (ld.acquire.sys) until all peers' epochs have arrived in the local flag array.
// The kernel therefore completes only when the whole output tensor of THIS rank is in its memory: later kernels of the
// stream / CUDA graph need nothing else.  No NCCL call, no staging buffer, no separate permute pass (the NCCL baseline is
// all_gather of the whole tensor + a strided copy, ~46 us per exchange at sizes where the wire time is <= 11 us).
// The 5-D GroupNorm statistics (2*32*B doubles per rank) use the same flags: store my partials into every rank's slot,
// signal, wait, sum the P slots in rank order (bit-identical on all ranks).
#include "common.cuh"
#include "../../include/videomv_b200.h"

#include <string.h>

namespace vmv {

void count_launch(int n = 1);

constexpr int PEER_MAX = VMV_PEER_MAX_RANKS;

struct PeerExArgs {
    const uint4* src;
    uint4* dst[PEER_MAX];
    unsigned int* flags[PEER_MAX];
    unsigned int* epoch;
    unsigned int* done;
    int world, rank, direction, nowait;
    int B, Fl;
    long long chunk_vecs;          // HWl * C / 8
};

// Publish `e` in slot `rank` of every rank's flag array, then wait until every rank's epoch has reached `e` in mine.
__device__ __forceinline__ void peer_signal_and_wait(unsigned int* const* flags, int world, int rank, unsigned int e, int nowait) {
    peer_publish(flags, world, rank, e);
    if (!nowait) peer_wait_all(flags, world, rank, e);
}

__global__ void __launch_bounds__(256)
peer_exchange_kernel(const PeerExArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const long long nchunks = (long long)a.world * a.B * a.Fl;
    const long long total = nchunks * a.chunk_vecs;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
        const long long ch = i / a.chunk_vecs, v = i - ch * a.chunk_vecs;
        const int f = (int)(ch % a.Fl);
        const int b = (int)((ch / a.Fl) % a.B);
        const int q = (int)(ch / ((long long)a.Fl * a.B));
        long long so, doff;
        if (a.direction == 0) {            // frames -> pixels: src [B, Fl, P(q), HWl, C]   dst_q [B, P(rank), Fl, HWl, C]
            so = (((long long)b * a.Fl + f) * a.world + q) * a.chunk_vecs;
            doff = (((long long)b * a.world + a.rank) * a.Fl + f) * a.chunk_vecs;
        } else {                           // pixels -> frames: src [B, P(q), Fl, HWl, C]   dst_q [B, Fl, P(rank), HWl, C]
            so = (((long long)b * a.world + q) * a.Fl + f) * a.chunk_vecs;
            doff = (((long long)b * a.Fl + f) * a.world + a.rank) * a.chunk_vecs;
        }
        a.dst[q][doff + v] = a.src[so + v];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();            // the CTA's remote stores are performed before it counts as arrived (cumulative over the barrier)
        const unsigned int prev = atomicAdd(a.done, 1u);
        if (prev == gridDim.x - 1) {       // last CTA of this rank: everything this rank had to send is on its way
            __threadfence();
            *a.done = 0;
            const unsigned int e = *a.epoch + 1;
            *a.epoch = e;
            peer_signal_and_wait(a.flags, a.world, a.rank, e, a.nowait);
        }
    }
}

struct PeerArArgs {
    double* data;
    double* slots[PEER_MAX];
    unsigned int* flags[PEER_MAX];
    unsigned int* epoch;
    int world, rank, n, nowait;
};

__global__ void __launch_bounds__(128)
peer_allreduce_kernel(const PeerArArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
        const double v = a.data[i];
        for (int q = 0; q < a.world; ++q) a.slots[q][(long long)a.rank * a.n + i] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int e = *a.epoch + 1;
        *a.epoch = e;
        peer_signal_and_wait(a.flags, a.world, a.rank, e, a.nowait);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
        double s = 0.0;
        for (int q = 0; q < a.world; ++q) s += __ldcg(a.slots[a.rank] + (long long)q * a.n + i);   // same order on every rank
        a.data[i] = s;
    }
}

// All-gather of a strided block: every rank stores its [nouter][inner] slice at (dst_base + outer * outer_stride) of the SAME
// destination tensor in every rank's arena (16 B remote stores), then the ranks meet at the epoch flags.  Used once per UNet
// call to assemble the output of all frame shards (and of the cond / uncond halves of a split CFG pair) on every rank.
struct PeerAgArgs {
    const uint4* src;
    uint4* dst[PEER_MAX];
    unsigned int* flags[PEER_MAX];
    unsigned int* epoch;
    unsigned int* done;
    int world, rank, nowait;
    long long nouter, inner_vecs, base_vecs, outer_stride_vecs;
};

__global__ void __launch_bounds__(256)
peer_allgather_kernel(const PeerAgArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    // Phase 0 ("ready to receive"): the destination is the SAME buffer every call, and a peer may still be reading the
    // previous call's result (its consumer kernels precede this kernel in ITS stream, not in mine).  Nobody writes before
    // every rank has entered this call.  Epochs advance by 2 per call: odd = ready, even = data delivered.
    if (threadIdx.x == 0) {
        const unsigned int e0 = *a.epoch + 1;               // stable until the last CTA of this launch has finished copying
        if (blockIdx.x == 0) peer_publish(a.flags, a.world, a.rank, e0);
        if (!a.nowait) peer_wait_all(a.flags, a.world, a.rank, e0);
    }
    __syncthreads();
    const long long total = a.nouter * a.inner_vecs * a.world;
    const long long per_rank = a.nouter * a.inner_vecs;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
        const int q = (int)(i / per_rank);
        const long long j = i - (long long)q * per_rank;
        const long long o = j / a.inner_vecs, v = j - o * a.inner_vecs;
        a.dst[q][a.base_vecs + o * a.outer_stride_vecs + v] = a.src[j];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int prev = atomicAdd(a.done, 1u);
        if (prev == gridDim.x - 1) {
            __threadfence();
            *a.done = 0;
            const unsigned int e = *a.epoch + 2;
            *a.epoch = e;
            peer_signal_and_wait(a.flags, a.world, a.rank, e, a.nowait);
        }
    }
}

typedef CUresult (*GetRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);

static GetRangeFn get_range_fn() {
    static GetRangeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<GetRangeFn>(p);
    }
    return fn;
}

}  // namespace vmv

using namespace vmv;

extern "C" int vmv_ipc_export(const void* ptr, void* handle64, int64_t* offset) {
    VMV_CHECK_ARG(ptr && handle64 && offset, "vmv_ipc_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    GetRangeFn fn = get_range_fn();
    if (!fn) { set_error("vmv_ipc_export: cuMemGetAddressRange not available"); return VMV_ERR_CUDA; }
    CUdeviceptr base = 0;
    size_t size = 0;
    CUresult r = fn(&base, &size, reinterpret_cast<CUdeviceptr>(ptr));
    if (r != CUDA_SUCCESS) { set_error("vmv_ipc_export: cuMemGetAddressRange failed (CUresult %d)", (int)r); return VMV_ERR_CUDA; }
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base));
    if (e != cudaSuccess) { set_error("vmv_ipc_export: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); return VMV_ERR_CUDA; }
    memcpy(handle64, &h, 64);
    *offset = (int64_t)(reinterpret_cast<CUdeviceptr>(ptr) - base);
    return VMV_OK;
}

extern "C" int vmv_ipc_import(const void* handle64, int64_t offset, void** out) {
    VMV_CHECK_ARG(handle64 && out && offset >= 0, "vmv_ipc_import: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* base = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_error("vmv_ipc_import: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e)); return VMV_ERR_CUDA; }
    *out = static_cast<char*>(base) + offset;
    return VMV_OK;
}

extern "C" int vmv_peer_exchange(const vmv_peer_exchange_params* p, void* stream) {
    VMV_CHECK_ARG(p && p->src && p->epoch && p->done, "vmv_peer_exchange: null pointer");
    VMV_CHECK_ARG(p->world >= 1 && p->world <= PEER_MAX && p->rank >= 0 && p->rank < p->world, "vmv_peer_exchange: bad world/rank");
    VMV_CHECK_ARG(p->B > 0 && p->Fl > 0 && p->HWl > 0 && p->C > 0 && ((int64_t)p->HWl * p->C) % 8 == 0,
                  "vmv_peer_exchange: bad geometry (HWl*C must be a multiple of 8)");
    VMV_CHECK_ARG(p->direction == 0 || p->direction == 1, "vmv_peer_exchange: direction must be 0 or 1");
    PeerExArgs a;
    memset(&a, 0, sizeof(a));
    a.src = static_cast<const uint4*>(p->src);
    uintptr_t al = reinterpret_cast<uintptr_t>(p->src);
    for (int q = 0; q < p->world; ++q) {
        VMV_CHECK_ARG(p->dst[q] && p->flags[q], "vmv_peer_exchange: null dst/flags for rank %d", q);
        a.dst[q] = static_cast<uint4*>(p->dst[q]);
        a.flags[q] = static_cast<unsigned int*>(p->flags[q]);
        al |= reinterpret_cast<uintptr_t>(p->dst[q]);
    }
    VMV_CHECK_ARG((al & 15) == 0, "vmv_peer_exchange: src/dst must be 16 B aligned");
    a.epoch = static_cast<unsigned int*>(p->epoch);
    a.done = static_cast<unsigned int*>(p->done);
    a.world = p->world; a.rank = p->rank; a.direction = p->direction; a.nowait = p->nowait;
    a.B = p->B; a.Fl = p->Fl;
    a.chunk_vecs = (long long)p->HWl * p->C / 8;
    const long long total = (long long)p->world * p->B * p->Fl * a.chunk_vecs;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
    }
    long long ctas = (total + 256 * 8 - 1) / (256 * 8);          // >= 8 vectors per thread
    if (ctas > 4LL * num_sms) ctas = 4LL * num_sms;
    if (ctas < 1) ctas = 1;
    launch_kernel(peer_exchange_kernel, dim3((unsigned)ctas), dim3(256), 0, static_cast<cudaStream_t>(stream), a);
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_peer_exchange");
    return VMV_OK;
}

extern "C" int vmv_peer_allgather(const vmv_peer_allgather_params* p, void* stream) {
    VMV_CHECK_ARG(p && p->src && p->epoch && p->done, "vmv_peer_allgather: null pointer");
    VMV_CHECK_ARG(p->world >= 1 && p->world <= PEER_MAX && p->rank >= 0 && p->rank < p->world, "vmv_peer_allgather: bad world/rank");
    VMV_CHECK_ARG(p->nouter > 0 && p->inner_bytes > 0 && p->inner_bytes % 16 == 0 && p->dst_offset_bytes % 16 == 0 &&
                      p->dst_outer_stride_bytes % 16 == 0 && p->dst_outer_stride_bytes >= p->inner_bytes,
                  "vmv_peer_allgather: sizes / offsets must be positive multiples of 16 bytes");
    PeerAgArgs a;
    memset(&a, 0, sizeof(a));
    a.src = static_cast<const uint4*>(p->src);
    uintptr_t al = reinterpret_cast<uintptr_t>(p->src);
    for (int q = 0; q < p->world; ++q) {
        VMV_CHECK_ARG(p->dst[q] && p->flags[q], "vmv_peer_allgather: null dst/flags for rank %d", q);
        a.dst[q] = static_cast<uint4*>(p->dst[q]);
        a.flags[q] = static_cast<unsigned int*>(p->flags[q]);
        al |= reinterpret_cast<uintptr_t>(p->dst[q]);
    }
    VMV_CHECK_ARG((al & 15) == 0, "vmv_peer_allgather: src/dst must be 16 B aligned");
    a.epoch = static_cast<unsigned int*>(p->epoch);
    a.done = static_cast<unsigned int*>(p->done);
    a.world = p->world; a.rank = p->rank; a.nowait = p->nowait;
    a.nouter = p->nouter; a.inner_vecs = p->inner_bytes / 16; a.base_vecs = p->dst_offset_bytes / 16;
    a.outer_stride_vecs = p->dst_outer_stride_bytes / 16;
    const long long total = a.nouter * a.inner_vecs * a.world;
    long long ctas = (total + 256 * 4 - 1) / (256 * 4);
    if (ctas > 296) ctas = 296;
    if (ctas < 1) ctas = 1;
    launch_kernel(peer_allgather_kernel, dim3((unsigned)ctas), dim3(256), 0, static_cast<cudaStream_t>(stream), a);
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_peer_allgather");
    return VMV_OK;
}

extern "C" int vmv_peer_allreduce_f64(const vmv_peer_allreduce_params* p, void* stream) {
    VMV_CHECK_ARG(p && p->data && p->epoch, "vmv_peer_allreduce_f64: null pointer");
    VMV_CHECK_ARG(p->world >= 1 && p->world <= PEER_MAX && p->rank >= 0 && p->rank < p->world, "vmv_peer_allreduce_f64: bad world/rank");
    VMV_CHECK_ARG(p->n > 0 && p->n <= 65536, "vmv_peer_allreduce_f64: bad n");
    PeerArArgs a;
    memset(&a, 0, sizeof(a));
    a.data = p->data;
    for (int q = 0; q < p->world; ++q) {
        VMV_CHECK_ARG(p->slots[q] && p->flags[q], "vmv_peer_allreduce_f64: null slots/flags for rank %d", q);
        a.slots[q] = p->slots[q];
        a.flags[q] = static_cast<unsigned int*>(p->flags[q]);
    }
    a.epoch = static_cast<unsigned int*>(p->epoch);
    a.world = p->world; a.rank = p->rank; a.n = p->n; a.nowait = p->nowait;
    launch_kernel(peer_allreduce_kernel, dim3(1), dim3(128), 0, static_cast<cudaStream_t>(stream), a);
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_peer_allreduce_f64");
    return VMV_OK;
}
