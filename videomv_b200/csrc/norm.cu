// GroupNorm(32) statistics / apply(+SiLU) and LayerNorm on channels-last fp16 rows.  HBM-bound kernels:
// 16-byte vector loads/stores along C, one fixed channel vector per thread so the per-channel scale/shift
// live in registers, fp32 partials per CTA, fp64 atomics across CTAs.
#include "common.cuh"
#include "../../include/videomv_b200.h"

#include <stdlib.h>
#include <string.h>

namespace vmv {

void count_launch(int n = 1);

constexpr int GN_THREADS = 256;
constexpr int GN_GROUPS = 32;

struct GnGeom {
    int C, C1, nvec, slabs, vw, lanes;   // vw = channel vectors per CTA slab, lanes = row lanes per CTA
    int rows_per_cta;
};

// max_ctas > 0: never exceed that many CTAs in total (fused kernel: all CTAs must be co-resident)
static GnGeom gn_geom(int C1, int C2, long long rows_per_batch, int nbatch, long long max_ctas = 0) {
    GnGeom g;
    g.C = C1 + C2;
    g.C1 = C1;
    g.nvec = g.C / 8;
    g.slabs = (g.nvec + GN_THREADS - 1) / GN_THREADS;
    g.vw = (g.nvec + g.slabs - 1) / g.slabs;
    g.lanes = GN_THREADS / g.vw;
    // aim for >= ~4 CTAs per SM overall while keeping >= 8 rows per row lane
    long long target_ctas = 148LL * 4;
    long long per_batch = (target_ctas + nbatch - 1) / nbatch;
    if (max_ctas > 0) {
        per_batch = max_ctas / ((long long)nbatch * g.slabs);
        if (per_batch < 1) per_batch = 1;
    }
    long long rpc = (rows_per_batch + per_batch - 1) / per_batch;
    long long min_rpc = (long long)g.lanes * 8;
    if (rpc < min_rpc) rpc = min_rpc;
    if (rpc > rows_per_batch) rpc = rows_per_batch;
    g.rows_per_cta = (int)rpc;
    return g;
}

__device__ __forceinline__ const uint4* gn_src(const __half* x1, long long ld1, int C1, const __half* x2,
                                               long long ld2, long long row, int c) {
    return (c < C1) ? reinterpret_cast<const uint4*>(x1 + row * ld1 + c)
                    : reinterpret_cast<const uint4*>(x2 + row * ld2 + (c - C1));
}

__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const __half* __restrict__ x1, long long ld1, int C1, const __half* __restrict__ x2, long long ld2,
                int C, long long rows_per_batch, int rows_per_cta, int vw, int lanes, double* __restrict__ stats) {
    __shared__ float s_sum[GN_GROUPS], s_sq[GN_GROUPS];
    const int t = threadIdx.x;
    if (t < GN_GROUPS) { s_sum[t] = 0.f; s_sq[t] = 0.f; }
    pdl_launch_dependents();
    pdl_wait();
    __syncthreads();
    const int batch = blockIdx.y;
    const int tx = t % vw, ty = t / vw;
    const int vec = blockIdx.z * vw + tx;
    const int c0 = vec * 8;
    const int cpg = C / GN_GROUPS;
    if (ty < lanes && c0 < C) {
        const long long r_begin = (long long)blockIdx.x * rows_per_cta;
        long long r_end = r_begin + rows_per_cta;
        if (r_end > rows_per_batch) r_end = rows_per_batch;
        const long long base = (long long)batch * rows_per_batch;
        float s[8], q[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
        auto acc = [&](const uint4& u) {
            uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 f = unpack_half2(w[j]);
                s[2 * j] += f.x; q[2 * j] += f.x * f.x;
                s[2 * j + 1] += f.y; q[2 * j + 1] += f.y * f.y;
            }
        };
        long long r = r_begin + ty;
        for (; r + 3LL * lanes < r_end; r += 4LL * lanes) {       // 4 independent 16B loads in flight per thread
            uint4 u0 = __ldg(gn_src(x1, ld1, C1, x2, ld2, base + r, c0));
            uint4 u1 = __ldg(gn_src(x1, ld1, C1, x2, ld2, base + r + lanes, c0));
            uint4 u2 = __ldg(gn_src(x1, ld1, C1, x2, ld2, base + r + 2LL * lanes, c0));
            uint4 u3 = __ldg(gn_src(x1, ld1, C1, x2, ld2, base + r + 3LL * lanes, c0));
            acc(u0); acc(u1); acc(u2); acc(u3);
        }
        for (; r < r_end; r += lanes) acc(__ldg(gn_src(x1, ld1, C1, x2, ld2, base + r, c0)));
        // fold the 8 channels into their groups (a vector may straddle two groups)
        int g_prev = c0 / cpg;
        float as = 0.f, aq = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int g = (c0 + j) / cpg;
            if (g != g_prev) {
                atomicAdd(&s_sum[g_prev], as); atomicAdd(&s_sq[g_prev], aq);
                as = 0.f; aq = 0.f; g_prev = g;
            }
            as += s[j]; aq += q[j];
        }
        atomicAdd(&s_sum[g_prev], as); atomicAdd(&s_sq[g_prev], aq);
    }
    __syncthreads();
    if (t < GN_GROUPS) {
        // only groups touched by this slab are non-zero; skip exact zeros to save atomics
        float a = s_sum[t], b = s_sq[t];
        if (a != 0.f || b != 0.f) {
            atomicAdd(&stats[((long long)batch * GN_GROUPS + t) * 2], (double)a);
            atomicAdd(&stats[((long long)batch * GN_GROUPS + t) * 2 + 1], (double)b);
        }
    }
}

__global__ void __launch_bounds__(GN_THREADS)
gn_apply_kernel(const __half* __restrict__ x1, long long ld1, int C1, const __half* __restrict__ x2, long long ld2,
                int C, long long rows_per_batch, long long stat_rows, int rows_per_cta, int vw, int lanes,
                const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                float eps, int silu, __half* __restrict__ out, long long ldo) {
    // Group mean / rstd once per CTA (32 threads do the fp64 part: E[x^2] - mean^2 cancels in fp32), not once per
    // channel per thread: 256 threads x 8 channels of fp64 div/sqrt cost more than streaming the CTA's rows.
    __shared__ float s_mean[GN_GROUPS], s_rstd[GN_GROUPS];
    const int t = threadIdx.x;
    const int batch = blockIdx.y;
    const int cpg = C / GN_GROUPS;
    pdl_launch_dependents();
    pdl_wait();
    if (t < GN_GROUPS) {
        const double inv_cnt = 1.0 / ((double)stat_rows * cpg);   // stat_rows > rows_per_batch: stats all-reduced over shards
        const double mean = stats[((long long)batch * GN_GROUPS + t) * 2] * inv_cnt;
        double var = stats[((long long)batch * GN_GROUPS + t) * 2 + 1] * inv_cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[t] = (float)mean;
        s_rstd[t] = rsqrtf((float)var + eps);
    }
    __syncthreads();
    const int tx = t % vw, ty = t / vw;
    const int vec = blockIdx.z * vw + tx;
    const int c0 = vec * 8;
    if (ty >= lanes || c0 >= C) return;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        const int g = c / cpg;
        const float ga = __ldg(gamma + c) * s_rstd[g];
        sc[j] = ga;
        sh[j] = __ldg(beta + c) - s_mean[g] * ga;
    }
    const long long r_begin = (long long)blockIdx.x * rows_per_cta;
    long long r_end = r_begin + rows_per_cta;
    if (r_end > rows_per_batch) r_end = rows_per_batch;
    const long long base = (long long)batch * rows_per_batch;
    auto apply = [&](const uint4& u, long long r) {
        uint32_t w[4] = {u.x, u.y, u.z, u.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 f = unpack_half2(w[j]);
            float a = f.x * sc[2 * j] + sh[2 * j];
            float b = f.y * sc[2 * j + 1] + sh[2 * j + 1];
            if (silu) { a = silu_f(a); b = silu_f(b); }
            o[j] = pack_half2(a, b);
        }
        *reinterpret_cast<uint4*>(out + (base + r) * ldo + c0) = make_uint4(o[0], o[1], o[2], o[3]);
    };
    long long r = r_begin + ty;
    for (; r + 3LL * lanes < r_end; r += 4LL * lanes) {
        uint4 u0 = __ldg(gn_src(x1, ld1, C1, x2, ld2, base + r, c0));
        uint4 u1 = __ldg(gn_src(x1, ld1, C1, x2, ld2, base + r + lanes, c0));
        uint4 u2 = __ldg(gn_src(x1, ld1, C1, x2, ld2, base + r + 2LL * lanes, c0));
        uint4 u3 = __ldg(gn_src(x1, ld1, C1, x2, ld2, base + r + 3LL * lanes, c0));
        apply(u0, r); apply(u1, r + lanes); apply(u2, r + 2LL * lanes); apply(u3, r + 3LL * lanes);
    }
    for (; r < r_end; r += lanes) apply(__ldg(gn_src(x1, ld1, C1, x2, ld2, base + r, c0)), r);
}

// Single-launch GroupNorm: statistics, a per-batch arrival barrier, then apply -- the second read of x comes from L2
// instead of HBM and the memset + two launches of the split path collapse into one graph node.  The barrier is a
// plain global counter, so every CTA of the grid must be co-resident: the host wrapper checks the grid against the
// occupancy of this kernel and refuses otherwise (the caller then uses the split kernels).  `stats` / `arrive` must
// be zero on entry (the engine zeroes one arena per forward).
__global__ void __launch_bounds__(GN_THREADS)
gn_fused_kernel(const __half* __restrict__ x1, long long ld1, int C1, const __half* __restrict__ x2, long long ld2,
                int C, long long rows_per_batch, int rows_per_cta, int vw, int lanes, double* __restrict__ stats,
                unsigned int* __restrict__ arrive, const float* __restrict__ gamma, const float* __restrict__ beta,
                float eps, int silu, __half* __restrict__ out, long long ldo, int opt) {
    // opt (VMV_GN_OPT, experiments): 1 = nanosleep back-off in the barrier spin; timing-only (wrong results):
    // 4 = do not wait at the barrier, 8 = skip the apply phase, 16 = skip the global statistics atomics
    __shared__ float s_sum[GN_GROUPS], s_sq[GN_GROUPS];
    const int t = threadIdx.x;
    if (t < GN_GROUPS) { s_sum[t] = 0.f; s_sq[t] = 0.f; }
    pdl_launch_dependents();
    pdl_wait();
    __syncthreads();
    const int batch = blockIdx.y;
    const int tx = t % vw, ty = t / vw;
    const int vec = blockIdx.z * vw + tx;
    const int c0 = vec * 8;
    const int cpg = C / GN_GROUPS;
    const bool active = ty < lanes && c0 < C;
    const long long r_begin = (long long)blockIdx.x * rows_per_cta;
    long long r_end = r_begin + rows_per_cta;
    if (r_end > rows_per_batch) r_end = rows_per_batch;
    const long long base = (long long)batch * rows_per_batch;
    // ---- phase 1: partial sums of this CTA's rows
    if (active) {
        float s[8], q[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
        auto acc = [&](const uint4& u) {
            uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 f = unpack_half2(w[j]);
                s[2 * j] += f.x; q[2 * j] += f.x * f.x;
                s[2 * j + 1] += f.y; q[2 * j + 1] += f.y * f.y;
            }
        };
        long long r = r_begin + ty;
        for (; r + 3LL * lanes < r_end; r += 4LL * lanes) {
            uint4 u0 = *gn_src(x1, ld1, C1, x2, ld2, base + r, c0);
            uint4 u1 = *gn_src(x1, ld1, C1, x2, ld2, base + r + lanes, c0);
            uint4 u2 = *gn_src(x1, ld1, C1, x2, ld2, base + r + 2LL * lanes, c0);
            uint4 u3 = *gn_src(x1, ld1, C1, x2, ld2, base + r + 3LL * lanes, c0);
            acc(u0); acc(u1); acc(u2); acc(u3);
        }
        for (; r < r_end; r += lanes) acc(*gn_src(x1, ld1, C1, x2, ld2, base + r, c0));
        int g_prev = c0 / cpg;
        float as = 0.f, aq = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int g = (c0 + j) / cpg;
            if (g != g_prev) {
                atomicAdd(&s_sum[g_prev], as); atomicAdd(&s_sq[g_prev], aq);
                as = 0.f; aq = 0.f; g_prev = g;
            }
            as += s[j]; aq += q[j];
        }
        atomicAdd(&s_sum[g_prev], as); atomicAdd(&s_sq[g_prev], aq);
    }
    __syncthreads();
    if (t < GN_GROUPS) {
        float a = s_sum[t], b = s_sq[t];
        if ((a != 0.f || b != 0.f) && !(opt & 16)) {
            atomicAdd(&stats[((long long)batch * GN_GROUPS + t) * 2], (double)a);
            atomicAdd(&stats[((long long)batch * GN_GROUPS + t) * 2 + 1], (double)b);
        }
        __threadfence();
    }
    __syncthreads();
    // ---- barrier over the CTAs of this batch (release: the fences above; acquire: the load below)
    __shared__ float s_mean[GN_GROUPS], s_rstd[GN_GROUPS];
    if (t == 0) {
        const unsigned int expected = gridDim.x * gridDim.z;
        atomicAdd(&arrive[batch], 1u);
        unsigned int seen, spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(arrive + batch) : "memory");
            if (++spins > (1u << 26)) __trap();               // co-residency was checked on the host; never a silent hang
            if (opt & 4) break;
            if ((opt & 1) && seen < expected) __nanosleep(100);
        } while (seen < expected);
    }
    __syncthreads();
    if (opt & 8) return;
    // ---- phase 2: normalise; x is re-read (L2) -- not through the non-coherent path, `out` may alias x
    if (t < GN_GROUPS) {
        const double inv_cnt = 1.0 / ((double)rows_per_batch * cpg);
        const double mean = __ldcg(&stats[((long long)batch * GN_GROUPS + t) * 2]) * inv_cnt;
        double var = __ldcg(&stats[((long long)batch * GN_GROUPS + t) * 2 + 1]) * inv_cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[t] = (float)mean;
        s_rstd[t] = rsqrtf((float)var + eps);
    }
    __syncthreads();
    if (!active) return;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        const int g = c / cpg;
        const float ga = __ldg(gamma + c) * s_rstd[g];
        sc[j] = ga;
        sh[j] = __ldg(beta + c) - s_mean[g] * ga;
    }
    auto apply = [&](const uint4& u, long long r) {
        uint32_t w[4] = {u.x, u.y, u.z, u.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 f = unpack_half2(w[j]);
            float a = f.x * sc[2 * j] + sh[2 * j];
            float b = f.y * sc[2 * j + 1] + sh[2 * j + 1];
            if (silu) { a = silu_f(a); b = silu_f(b); }
            o[j] = pack_half2(a, b);
        }
        *reinterpret_cast<uint4*>(out + (base + r) * ldo + c0) = make_uint4(o[0], o[1], o[2], o[3]);
    };
    long long r = r_begin + ty;
    for (; r + 3LL * lanes < r_end; r += 4LL * lanes) {
        uint4 u0 = *gn_src(x1, ld1, C1, x2, ld2, base + r, c0);
        uint4 u1 = *gn_src(x1, ld1, C1, x2, ld2, base + r + lanes, c0);
        uint4 u2 = *gn_src(x1, ld1, C1, x2, ld2, base + r + 2LL * lanes, c0);
        uint4 u3 = *gn_src(x1, ld1, C1, x2, ld2, base + r + 3LL * lanes, c0);
        apply(u0, r); apply(u1, r + lanes); apply(u2, r + 2LL * lanes); apply(u3, r + 3LL * lanes);
    }
    for (; r < r_end; r += lanes) apply(*gn_src(x1, ld1, C1, x2, ld2, base + r, c0), r);
}


// ------------------------------------------------------------------------------------------------
// Single-pass GroupNorm: the CTA's rows are brought into shared memory ONCE with bulk async copies (one mbarrier, the
// whole slab in flight at a time -- no per-thread load latency chain), statistics are taken from smem, the CTAs of a
// chunk meet at the same arrival barrier as gn_fused_kernel, and the apply phase reads smem and writes global: HBM sees
// one read and one write of the tensor.  One CTA per SM; a [49152, 320] level-0 activation is 213 KB per SM, just inside
// the 227 KB a CTA may own, the lower levels are smaller.  Inputs that do not fit (the widest skip concatenations) use
// gn_fused_kernel.  `stats` / `arrive` must be zero on entry.
// ------------------------------------------------------------------------------------------------
constexpr int GNS_THREADS = 512;
constexpr int GNS_MAX_DYN_SMEM = 227 * 1024 - 2048;       // the kernel's static smem (barrier + 4 x 32 floats, 1.7 KB) comes on top
constexpr int GNS_CHUNK_BYTES = 16384;

// Cross-GPU part of a 5-D GroupNorm in the pixel-sharded layout (multi-GPU frame sharding, csrc/peer.cu): every rank holds
// all frames of HW/P pixels, so the statistics of a sample are the sum over ranks.  The first CTA of a chunk publishes this
// rank's 64 partial sums into every rank's slot and raises an epoch flag there; every CTA of the chunk waits for the P epochs
// in the local flag array and sums the P slots in rank order.  world <= 1: single-GPU behaviour.
struct GnPeer {
    int world, rank;
    double* slots[VMV_PEER_MAX_RANKS];            // [world][nbatch][64] doubles in every rank's arena
    unsigned int* flags[VMV_PEER_MAX_RANKS];      // [nbatch][16] u32 in every rank's arena (8 used per chunk)
    unsigned int* epoch;                          // local [nbatch]
    long long stat_rows;                          // rows per chunk summed over all ranks
};

__device__ __forceinline__ void gn_st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int gn_ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(GNS_THREADS, 1)
gn_smem_kernel(const __half* __restrict__ x1, long long ld1, int C1, const __half* __restrict__ x2, long long ld2, int C2,
               long long rows_per_batch, int rows_per_cta, double* __restrict__ stats, unsigned int* __restrict__ arrive,
               const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu,
               __half* __restrict__ out, long long ldo, const GnPeer pe) {
    extern __shared__ __align__(128) uint8_t gsm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ float s_sum[GN_GROUPS], s_sq[GN_GROUPS], s_mean[GN_GROUPS], s_rstd[GN_GROUPS];
    const int t = threadIdx.x;
    const int C = C1 + C2;
    const int batch = blockIdx.y;
    const long long r_begin = (long long)blockIdx.x * rows_per_cta;
    long long left = rows_per_batch - r_begin;
    const int nrows = left <= 0 ? 0 : (left < rows_per_cta ? (int)left : rows_per_cta);
    const long long base_row = (long long)batch * rows_per_batch + r_begin;
    if (t == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (t < GN_GROUPS) { s_sum[t] = 0.f; s_sq[t] = 0.f; }
    __syncthreads();
    pdl_launch_dependents();
    pdl_wait();
    // the epoch this launch will use: read before this CTA arrives anywhere, i.e. before the publisher can advance it
    unsigned int peer_epoch = 0;
    if (pe.world > 1 && t == 0) peer_epoch = *(pe.epoch + batch) + 1;
    // ---- load: the whole slab in flight at once
    if (t < 32) {
        const uint32_t row_bytes = (uint32_t)C * 2u;
        if (t == 0) mbar_arrive_expect_tx(&bar, (uint32_t)nrows * row_bytes);
        __syncwarp();
        if (C2 == 0 && ld1 == C1) {                      // rows are contiguous in memory: 16 KB chunks
            const int chunk_rows = max(1, GNS_CHUNK_BYTES / (int)row_bytes);
            const int nchunks = (nrows + chunk_rows - 1) / chunk_rows;
            for (int i = t; i < nchunks; i += 32) {
                const int r0 = i * chunk_rows;
                const int n = min(chunk_rows, nrows - r0);
                bulk_g2s(gsm + (size_t)r0 * row_bytes, x1 + (base_row + r0) * ld1, (uint32_t)n * row_bytes, &bar);
            }
        } else {                                         // strided and / or two sources: one or two copies per row
            for (int r = t; r < nrows; r += 32) {
                bulk_g2s(gsm + (size_t)r * row_bytes, x1 + (base_row + r) * ld1, (uint32_t)C1 * 2u, &bar);
                if (C2) bulk_g2s(gsm + (size_t)r * row_bytes + (size_t)C1 * 2, x2 + (base_row + r) * ld2, (uint32_t)C2 * 2u, &bar);
            }
        }
    }
    mbar_wait(&bar, 0);
    // ---- phase 1: statistics from smem.  Thread = one 8-channel vector x every `lanes`-th row.
    const int vw = C / 8;
    const int lanes = GNS_THREADS / vw;                  // >= 1: the host guarantees C <= 8 * GNS_THREADS
    const int tx = t % vw, ty = t / vw;
    const int c0 = tx * 8;
    const int cpg = C / GN_GROUPS;
    const bool active = ty < lanes;
    const uint8_t* colp = gsm + (size_t)c0 * 2;
    const size_t rstride = (size_t)C * 2;
    if (active) {
        float s[8], q[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
#pragma unroll 4
        for (int r = ty; r < nrows; r += lanes) {
            const uint4 u = *reinterpret_cast<const uint4*>(colp + r * rstride);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_half2(w[j]);
                s[2 * j] += f.x; q[2 * j] = fmaf(f.x, f.x, q[2 * j]);
                s[2 * j + 1] += f.y; q[2 * j + 1] = fmaf(f.y, f.y, q[2 * j + 1]);
            }
        }
        int g_prev = c0 / cpg;
        float as = 0.f, aq = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int g = (c0 + j) / cpg;
            if (g != g_prev) {
                atomicAdd(&s_sum[g_prev], as); atomicAdd(&s_sq[g_prev], aq);
                as = 0.f; aq = 0.f; g_prev = g;
            }
            as += s[j]; aq += q[j];
        }
        atomicAdd(&s_sum[g_prev], as); atomicAdd(&s_sq[g_prev], aq);
    }
    __syncthreads();
    const unsigned int expected = gridDim.x;
    if (expected > 1) {
        if (t < GN_GROUPS) {
            atomicAdd(&stats[((long long)batch * GN_GROUPS + t) * 2], (double)s_sum[t]);
            atomicAdd(&stats[((long long)batch * GN_GROUPS + t) * 2 + 1], (double)s_sq[t]);
            __threadfence();
        }
        __syncthreads();
        if (t == 0) {
            atomicAdd(&arrive[batch], 1u);
            unsigned int seen, spins = 0;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(arrive + batch) : "memory");
                if (++spins > (1u << 26)) __trap();           // co-residency is guaranteed by the host; never a silent hang
            } while (seen < expected);
        }
        __syncthreads();
    }
    if (pe.world > 1) {
        const int nb = gridDim.y;
        if (blockIdx.x == 0) {                               // publisher of this chunk
            if (t < GN_GROUPS) {
                double sum, sq;
                if (expected > 1) {
                    sum = __ldcg(&stats[((long long)batch * GN_GROUPS + t) * 2]);
                    sq = __ldcg(&stats[((long long)batch * GN_GROUPS + t) * 2 + 1]);
                } else {
                    sum = (double)s_sum[t];
                    sq = (double)s_sq[t];
                }
                const long long o = (((long long)pe.rank * nb + batch) * GN_GROUPS + t) * 2;
                for (int q = 0; q < pe.world; ++q) { pe.slots[q][o] = sum; pe.slots[q][o + 1] = sq; }
            }
            __threadfence_system();
            __syncthreads();
            if (t == 0)
                for (int q = 0; q < pe.world; ++q) gn_st_release_sys(pe.flags[q] + batch * 16 + pe.rank, peer_epoch);
        }
        if (t == 0) {
            for (int q = 0; q < pe.world; ++q) {
                unsigned long long spins = 0;
                while ((int)(gn_ld_acquire_sys(pe.flags[pe.rank] + batch * 16 + q) - peer_epoch) < 0) {
                    if (++spins > (1ull << 27)) __trap();        // seconds: a peer that never arrives fails the launch
                    __nanosleep(20);
                }
            }
            if (blockIdx.x == 0) *(pe.epoch + batch) = peer_epoch;
        }
        __syncthreads();
    }
    if (t < GN_GROUPS) {
        const double inv_cnt = 1.0 / ((double)(pe.world > 1 ? pe.stat_rows : rows_per_batch) * cpg);
        double sum, sq;
        if (pe.world > 1) {                                  // sum of the ranks' partials, in rank order
            const int nb = gridDim.y;
            sum = 0.0; sq = 0.0;
            for (int q = 0; q < pe.world; ++q) {
                const long long o = (((long long)q * nb + batch) * GN_GROUPS + t) * 2;
                sum += __ldcg(pe.slots[pe.rank] + o);
                sq += __ldcg(pe.slots[pe.rank] + o + 1);
            }
        } else if (expected > 1) {
            sum = __ldcg(&stats[((long long)batch * GN_GROUPS + t) * 2]);
            sq = __ldcg(&stats[((long long)batch * GN_GROUPS + t) * 2 + 1]);
        } else {                                             // the chunk is mine alone: no global round trip
            sum = (double)s_sum[t];
            sq = (double)s_sq[t];
        }
        const double mean = sum * inv_cnt;
        double var = sq * inv_cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[t] = (float)mean;
        s_rstd[t] = rsqrtf((float)var + eps);
    }
    __syncthreads();
    if (!active) return;
    // ---- phase 2: normalise from smem, write global
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        const int g = c / cpg;
        const float ga = __ldg(gamma + c) * s_rstd[g];
        sc[j] = ga;
        sh[j] = __ldg(beta + c) - s_mean[g] * ga;
    }
    __half* op = out + base_row * ldo + c0;
#pragma unroll 4
    for (int r = ty; r < nrows; r += lanes) {
        const uint4 u = *reinterpret_cast<const uint4*>(colp + r * rstride);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_half2(w[j]);
            float a = fmaf(f.x, sc[2 * j], sh[2 * j]);
            float b = fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1]);
            if (silu) { a = silu_f(a); b = silu_f(b); }
            o[j] = pack_half2(a, b);
        }
        *reinterpret_cast<uint4*>(op + r * ldo) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// One warp per row; the row lives in registers (<= 8 vectors of 8 halfs per lane => C <= 2048).
constexpr int LN_MAX_VEC = 8;
__global__ void __launch_bounds__(256)
layernorm_kernel(const __half* __restrict__ x, long long ldx, long long M, int C, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, __half* __restrict__ out, long long ldo) {
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();
    pdl_wait();
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int nvec = C / 8;
    float v[LN_MAX_VEC][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_VEC; ++i) {
        const int vec = lane + i * 32;
        if (vec < nvec) {
            uint4 u = __ldg(reinterpret_cast<const uint4*>(x + row * ldx + vec * 8));
            uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 f = unpack_half2(w[j]);
                v[i][2 * j] = f.x; v[i][2 * j + 1] = f.y;
                sum += f.x + f.y;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_VEC; ++i) {
        const int vec = lane + i * 32;
        if (vec < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { float d = v[i][j] - mean; sq += d * d; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)C + eps);
#pragma unroll
    for (int i = 0; i < LN_MAX_VEC; ++i) {
        const int vec = lane + i * 32;
        if (vec < nvec) {
            const int c = vec * 8;
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                o[j] = pack_half2((v[i][2 * j] - mean) * rstd * g[2 * j] + b[2 * j],
                                  (v[i][2 * j + 1] - mean) * rstd * g[2 * j + 1] + b[2 * j + 1]);
            *reinterpret_cast<uint4*>(out + row * ldo + c) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// mean / rstd of each row (exact two-pass in registers), no normalised output: the consumer GEMM folds the affine part.
// A warp handles ROWS rows at once and issues all of their 16 B loads before reducing (memory-level parallelism:
// this kernel is a pure read stream); NV = vectors per lane per row.
template <int NV, int ROWS>
__global__ void __launch_bounds__(256)
layernorm_stats_kernel(const __half* __restrict__ x, long long ldx, long long M, int C, float eps, float2* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();
    pdl_wait();
    const long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ROWS;
    if (row0 >= M) return;
    const int nvec = C / 8;
    uint4 u[ROWS][NV];
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
        const long long row = row0 + rr < M ? row0 + rr : M - 1;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int vec = lane + i * 32;
            u[rr][i] = vec < nvec ? __ldg(reinterpret_cast<const uint4*>(x + row * ldx + vec * 8)) : make_uint4(0, 0, 0, 0);
        }
    }
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
        float v[NV][8];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const uint32_t w[4] = {u[rr][i].x, u[rr][i].y, u[rr][i].z, u[rr][i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_half2(w[j]);
                v[i][2 * j] = f.x; v[i][2 * j + 1] = f.y;
                sum += f.x + f.y;                       // zero padding beyond nvec adds nothing
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum / (float)C;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (lane + i * 32 < nvec) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; sq += d * d; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (lane == 0 && row0 + rr < M) stats[row0 + rr] = make_float2(mean, rsqrtf(sq / (float)C + eps));
    }
}

}  // namespace vmv

using namespace vmv;

extern "C" int vmv_layernorm_stats(const void* x, int64_t ldx, int64_t M, int32_t C, float eps, void* stats, void* stream) {
    VMV_CHECK_ARG(x && stats, "vmv_layernorm_stats: null pointer");
    VMV_CHECK_ARG(C > 0 && C % 8 == 0 && C <= LN_MAX_VEC * 32 * 8 && ldx % 8 == 0 && M > 0, "vmv_layernorm_stats: bad C/ld/M");
    const int wpb = 8;
    const int nv = (C / 8 + 31) / 32;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const __half* xp = static_cast<const __half*>(x);
    float2* sp = static_cast<float2*>(stats);
#define VMV_LNS(NV_, R_) launch_kernel(layernorm_stats_kernel<NV_, R_>, dim3((unsigned)((M + wpb * R_ - 1) / (wpb * R_))), dim3(wpb * 32), 0, st, xp, ldx, M, C, eps, sp)
    if (nv <= 1) VMV_LNS(1, 4);
    else if (nv == 2) VMV_LNS(2, 4);
    else if (nv == 3) VMV_LNS(3, 2);
    else if (nv <= 5) VMV_LNS(5, 2);
    else VMV_LNS(8, 1);
#undef VMV_LNS
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_layernorm_stats");
    return VMV_OK;
}

static int gn_check(const char* who, const void* x1, int64_t ldx1, int C1, const void* x2, int64_t ldx2, int C2,
                    int64_t rows_per_batch, int nbatch) {
    VMV_CHECK_ARG(x1 && C1 > 0 && C1 % 8 == 0 && ldx1 % 8 == 0, "%s: bad x1/C1/ldx1", who);
    VMV_CHECK_ARG(C2 == 0 || (x2 && C2 % 8 == 0 && ldx2 % 8 == 0), "%s: bad x2/C2/ldx2", who);
    VMV_CHECK_ARG((C1 + C2) % GN_GROUPS == 0, "%s: C=%d not divisible by 32 groups", who, C1 + C2);
    VMV_CHECK_ARG(rows_per_batch > 0 && nbatch > 0 && nbatch <= 65535, "%s: bad rows_per_batch/nbatch", who);
    return VMV_OK;
}

extern "C" int vmv_groupnorm_stats(const void* x1, int64_t ldx1, int32_t C1, const void* x2, int64_t ldx2, int32_t C2,
                                   int64_t rows_per_batch, int32_t nbatch, double* stats, void* stream) {
    int rc = gn_check("vmv_groupnorm_stats", x1, ldx1, C1, x2, ldx2, C2, rows_per_batch, nbatch);
    if (rc) return rc;
    VMV_CHECK_ARG(stats != nullptr, "vmv_groupnorm_stats: null stats");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * 2 * GN_GROUPS * nbatch, st);
    if (e != cudaSuccess) { set_error("vmv_groupnorm_stats: memset failed: %s", cudaGetErrorString(e)); return VMV_ERR_CUDA; }
    GnGeom g = gn_geom(C1, C2, rows_per_batch, nbatch);
    dim3 grid((unsigned)((rows_per_batch + g.rows_per_cta - 1) / g.rows_per_cta), nbatch, g.slabs);
    launch_kernel(gn_stats_kernel, grid, dim3(GN_THREADS), 0, st, static_cast<const __half*>(x1), ldx1, C1,
                                                 static_cast<const __half*>(x2), ldx2, g.C, rows_per_batch,
                                                 g.rows_per_cta, g.vw, g.lanes, stats);
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_groupnorm_stats");
    return VMV_OK;
}

extern "C" int vmv_groupnorm_apply(const void* x1, int64_t ldx1, int32_t C1, const void* x2, int64_t ldx2, int32_t C2,
                                   int64_t rows_per_batch, int32_t nbatch, const double* stats, int64_t stat_rows,
                                   const float* gamma, const float* beta, float eps, int32_t silu, void* out, int64_t ldo,
                                   void* stream) {
    int rc = gn_check("vmv_groupnorm_apply", x1, ldx1, C1, x2, ldx2, C2, rows_per_batch, nbatch);
    if (rc) return rc;
    VMV_CHECK_ARG(stats && gamma && beta && out && ldo % 8 == 0 && ldo >= C1 + C2, "vmv_groupnorm_apply: bad stats/gamma/beta/out");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GnGeom g = gn_geom(C1, C2, rows_per_batch, nbatch);
    dim3 grid((unsigned)((rows_per_batch + g.rows_per_cta - 1) / g.rows_per_cta), nbatch, g.slabs);
    launch_kernel(gn_apply_kernel, grid, dim3(GN_THREADS), 0, st, static_cast<const __half*>(x1), ldx1, C1,
                                                 static_cast<const __half*>(x2), ldx2, g.C, rows_per_batch,
                                                 stat_rows > 0 ? stat_rows : rows_per_batch, g.rows_per_cta, g.vw,
                                                 g.lanes, stats, gamma, beta, eps, silu,
                                                 static_cast<__half*>(out), ldo);
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_groupnorm_apply");
    return VMV_OK;
}

static int gn_opt() {
    const char* e = getenv("VMV_GN_OPT");          // read per call: the micro-benchmark switches it between launches
    return e ? atoi(e) : 0;
}

// Scratch for one fused call: nbatch*64 doubles (sums) followed by nbatch uint32 arrival counters, all zero on entry.
extern "C" int64_t vmv_groupnorm_fused_scratch_bytes(int32_t nbatch) {
    return (int64_t)nbatch * 2 * GN_GROUPS * 8 + (((int64_t)nbatch * 4 + 7) / 8) * 8;
}

static int gn_fused_impl(const void* x1, int64_t ldx1, int32_t C1, const void* x2, int64_t ldx2, int32_t C2,
                         int64_t rows_per_batch, int32_t nbatch, void* scratch, const float* gamma, const float* beta,
                         float eps, int32_t silu, void* out, int64_t ldo, const vmv_gn_peer* peer, void* stream) {
    int rc = gn_check("vmv_groupnorm_fused", x1, ldx1, C1, x2, ldx2, C2, rows_per_batch, nbatch);
    if (rc) return rc;
    VMV_CHECK_ARG(scratch && gamma && beta && out && ldo % 8 == 0 && ldo >= C1 + C2, "vmv_groupnorm_fused: bad scratch/gamma/beta/out");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* stats = static_cast<double*>(scratch);
    unsigned int* arrive = reinterpret_cast<unsigned int*>(stats + (size_t)nbatch * 2 * GN_GROUPS);
    {
        // smem-resident single pass whenever one CTA per SM can hold the tensor (VMV_GN_SMEM=0: always the re-read kernel)
        static int use_smem = -1, num_sms = 0;
        if (use_smem < 0) {
            const char* e = getenv("VMV_GN_SMEM");
            use_smem = (e && e[0] == '0') ? 0 : 1;
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
            if (use_smem) {
                cudaError_t e2 = cudaFuncSetAttribute(gn_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GNS_MAX_DYN_SMEM);
                if (e2 != cudaSuccess) { set_error("vmv_groupnorm_fused: smem attribute: %s", cudaGetErrorString(e2)); return VMV_ERR_CUDA; }
            }
        }
        const int C = C1 + C2;
        const bool al16 = ((reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(x2) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
        if (use_smem && al16 && C <= 8 * GNS_THREADS && nbatch <= num_sms) {
            long long cpb = num_sms / nbatch;                      // CTAs per chunk; all CTAs co-resident at one per SM
            if (cpb > rows_per_batch) cpb = rows_per_batch;
            const long long rpc = (rows_per_batch + cpb - 1) / cpb;
            cpb = (rows_per_batch + rpc - 1) / rpc;                // drop CTAs that would own no rows
            const long long smem = rpc * C * 2;
            if (smem <= GNS_MAX_DYN_SMEM) {
                GnPeer pe;
                memset(&pe, 0, sizeof(pe));
                if (peer != nullptr && peer->world > 1) {
                    pe.world = peer->world; pe.rank = peer->rank; pe.stat_rows = peer->stat_rows;
                    pe.epoch = static_cast<unsigned int*>(peer->epoch);
                    for (int q = 0; q < peer->world; ++q) {
                        pe.slots[q] = peer->slots[q];
                        pe.flags[q] = static_cast<unsigned int*>(peer->flags[q]);
                    }
                }
                launch_kernel(gn_smem_kernel, dim3((unsigned)cpb, nbatch), dim3(GNS_THREADS), (size_t)smem, st,
                              static_cast<const __half*>(x1), ldx1, C1, static_cast<const __half*>(x2), ldx2, C2, rows_per_batch,
                              (int)rpc, stats, arrive, gamma, beta, eps, silu, static_cast<__half*>(out), ldo, pe);
                count_launch();
                VMV_CUDA_LAUNCH_CHECK("vmv_groupnorm_fused (smem)");
                return VMV_OK;
            }
        }
    }
    if (peer != nullptr && peer->world > 1) {
        set_error("vmv_groupnorm_fused_peer: the tensor does not fit the smem-resident kernel (%lld rows x %d channels per chunk); "
                  "use vmv_groupnorm_stats + vmv_peer_allreduce_f64 + vmv_groupnorm_apply", (long long)rows_per_batch, C1 + C2);
        return VMV_ERR_UNSUPPORTED;
    }
    static int capacity = 0;                                   // co-resident CTAs of gn_fused_kernel on this device
    if (capacity == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_fused_kernel, GN_THREADS, 0);
        if (e != cudaSuccess) { set_error("vmv_groupnorm_fused: occupancy query failed: %s", cudaGetErrorString(e)); return VMV_ERR_CUDA; }
        capacity = sms * per_sm;
    }
    GnGeom g = gn_geom(C1, C2, rows_per_batch, nbatch, capacity);
    dim3 grid((unsigned)((rows_per_batch + g.rows_per_cta - 1) / g.rows_per_cta), nbatch, g.slabs);
    if ((long long)grid.x * grid.y * grid.z > capacity) {
        set_error("vmv_groupnorm_fused: grid of %u CTAs exceeds the %d co-resident CTAs the in-kernel barrier needs; "
                  "use vmv_groupnorm_stats + vmv_groupnorm_apply", grid.x * grid.y * grid.z, capacity);
        return VMV_ERR_UNSUPPORTED;
    }
    launch_kernel(gn_fused_kernel, grid, dim3(GN_THREADS), 0, st, static_cast<const __half*>(x1), ldx1, C1,
                                                 static_cast<const __half*>(x2), ldx2, g.C, rows_per_batch,
                                                 g.rows_per_cta, g.vw, g.lanes, stats, arrive, gamma, beta, eps, silu,
                                                 static_cast<__half*>(out), ldo, gn_opt());
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_groupnorm_fused");
    return VMV_OK;
}

extern "C" int vmv_groupnorm_fused(const void* x1, int64_t ldx1, int32_t C1, const void* x2, int64_t ldx2, int32_t C2,
                                   int64_t rows_per_batch, int32_t nbatch, void* scratch, const float* gamma,
                                   const float* beta, float eps, int32_t silu, void* out, int64_t ldo, void* stream) {
    return gn_fused_impl(x1, ldx1, C1, x2, ldx2, C2, rows_per_batch, nbatch, scratch, gamma, beta, eps, silu, out, ldo, nullptr, stream);
}

extern "C" int vmv_groupnorm_fused_peer(const void* x1, int64_t ldx1, int32_t C1, const void* x2, int64_t ldx2, int32_t C2,
                                        int64_t rows_per_batch, int32_t nbatch, void* scratch, const float* gamma,
                                        const float* beta, float eps, int32_t silu, void* out, int64_t ldo,
                                        const vmv_gn_peer* peer, void* stream) {
    VMV_CHECK_ARG(peer && peer->world >= 1 && peer->world <= VMV_PEER_MAX_RANKS && peer->rank >= 0 && peer->rank < peer->world,
                  "vmv_groupnorm_fused_peer: bad world/rank");
    VMV_CHECK_ARG(peer->epoch && peer->stat_rows >= rows_per_batch, "vmv_groupnorm_fused_peer: bad epoch/stat_rows");
    for (int q = 0; q < peer->world; ++q)
        VMV_CHECK_ARG(peer->slots[q] && peer->flags[q], "vmv_groupnorm_fused_peer: null slots/flags for rank %d", q);
    return gn_fused_impl(x1, ldx1, C1, x2, ldx2, C2, rows_per_batch, nbatch, scratch, gamma, beta, eps, silu, out, ldo, peer, stream);
}

extern "C" int vmv_layernorm(const void* x, int64_t ldx, int64_t M, int32_t C, const float* gamma, const float* beta,
                             float eps, void* out, int64_t ldo, void* stream) {
    VMV_CHECK_ARG(x && out && gamma && beta, "vmv_layernorm: null pointer");
    VMV_CHECK_ARG(C > 0 && C % 8 == 0 && C <= LN_MAX_VEC * 32 * 8, "vmv_layernorm: C=%d must be a multiple of 8 and <= %d", C, LN_MAX_VEC * 256);
    VMV_CHECK_ARG(ldx % 8 == 0 && ldo % 8 == 0 && M > 0, "vmv_layernorm: bad ld/M");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int wpb = 8;
    launch_kernel(layernorm_kernel, dim3((unsigned)((M + wpb - 1) / wpb)), dim3(wpb * 32), 0, st,
        static_cast<const __half*>(x), ldx, M, C, gamma, beta, eps, static_cast<__half*>(out), ldo);
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_layernorm");
    return VMV_OK;
}
