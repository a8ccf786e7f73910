// GroupNorm(32) statistics / apply(+SiLU) and LayerNorm on channels-last fp16 rows.  HBM-bound kernels:
// 16-byte vector loads/stores along C, one fixed channel vector per thread so the per-channel scale/shift
// live in registers, fp32 partials per CTA, fixed-order fp64 sums across CTAs (no atomics: bit-reproducible).
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

// ------------------------------------------------------------------------------------------------
// Deterministic reductions (no floating-point atomics anywhere: two runs are bit-identical).
//   (1) inside a CTA: every thread folds its 8 channel sums into per-group partials and parks them in a shared scratch
//       (`ks` float2 slots per thread: a vector of 8 channels touches at most 2 groups when C/32 >= 8, up to 8 otherwise);
//       NT/32 threads per group then add the contributors in a fixed order and finish with a shuffle tree.
//   (2) across the CTAs of a chunk: each CTA writes its 32 {sum, sumsq} to its own slot; after the arrival barrier the
//       slots are added in CTA order (fp64), again NT/32 threads per group + shuffle tree.
//   The arrival barrier is self-resetting ({count, generation} per chunk: the last arriver zeroes the count and bumps the
//   generation), so its words are zeroed once at allocation and never again.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int gn_slots_per_thread(int C) { return (C / GN_GROUPS) >= 8 ? 2 : 8; }

template <int NT>
__device__ __forceinline__ void gn_cta_group_sums(const float (&s)[8], const float (&q)[8], bool active, int c0, int cpg,
                                                  int vec0, int vw, int lanes, int ks, float2* scr, float* s_sum, float* s_sq) {
    const int t = threadIdx.x;
    if (active) {
        const int g0 = c0 / cpg;
        int g_prev = g0;
        float as = 0.f, aq = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int g = (c0 + j) / cpg;
            if (g != g_prev) {
                scr[t * ks + (g_prev - g0)] = make_float2(as, aq);
                as = 0.f; aq = 0.f; g_prev = g;
            }
            as += s[j]; aq += q[j];
        }
        scr[t * ks + (g_prev - g0)] = make_float2(as, aq);
    }
    __syncthreads();
    constexpr int W = NT / GN_GROUPS;                    // threads per group (8 or 16: a shuffle tree inside one warp)
    const int g = t / W, i = t % W;
    int vlo = (g * cpg) / 8, vhi = ((g + 1) * cpg - 1) / 8;
    if (vlo < vec0) vlo = vec0;
    if (vhi > vec0 + vw - 1) vhi = vec0 + vw - 1;
    const int nv = vhi - vlo + 1;
    float ps = 0.f, pq = 0.f;
    if (nv > 0) {
        const int total = lanes * nv;
        for (int idx = i; idx < total; idx += W) {
            const int ty = idx / nv, v = vlo + (idx - ty * nv);
            const int j = g - (v * 8) / cpg;
            const float2 f = scr[(ty * vw + (v - vec0)) * ks + j];
            ps += f.x; pq += f.y;
        }
    }
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) {
        ps += __shfl_xor_sync(0xffffffffu, ps, o);
        pq += __shfl_xor_sync(0xffffffffu, pq, o);
    }
    if (i == 0) { s_sum[g] = ps; s_sq[g] = pq; }
    __syncthreads();
}

// Sum the per-CTA slots of one chunk in CTA order; the totals of group g land in tot[2g], tot[2g+1] (shared, fp64).
template <int NT>
__device__ __forceinline__ void gn_sum_slots(const float2* slots, int ncta, double* tot) {
    constexpr int W = NT / GN_GROUPS;
    const int t = threadIdx.x, g = t / W, i = t % W;
    double ds = 0.0, dq = 0.0;
    for (int c = i; c < ncta; c += W) {
        const float2 f = __ldcg(slots + (long long)c * GN_GROUPS + g);
        ds += (double)f.x; dq += (double)f.y;
    }
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) {
        ds += __shfl_xor_sync(0xffffffffu, ds, o);
        dq += __shfl_xor_sync(0xffffffffu, dq, o);
    }
    if (i == 0) { tot[2 * g] = ds; tot[2 * g + 1] = dq; }
    __syncthreads();
}

// Arrive at the chunk's barrier (all threads call; the CTA's slot writes precede it).  wait = true: returns once every CTA
// of the chunk has arrived.  wait = false: returns immediately; the result is true in the LAST CTA to arrive only.
__device__ __forceinline__ bool gn_barrier(unsigned int* bar /* {count, generation} */, unsigned int expected, bool wait) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int gen;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(bar + 1) : "memory");
        const unsigned int prev = atomicAdd(bar, 1u);
        const int last = prev == expected - 1;
        if (last) {
            __threadfence();
            atomicExch(bar, 0u);
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 1), "r"(gen + 1) : "memory");
        } else if (wait) {
            unsigned int seen, spins = 0;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar + 1) : "memory");
                if (++spins > (1u << 26)) __trap();               // co-residency is guaranteed by the host; never a silent hang
            } while (seen == gen);
        }
        s_last = last;
        __threadfence();
    }
    __syncthreads();
    return s_last != 0;
}

template <bool NC /* read-only path: x is not written by this kernel */>
__device__ __forceinline__ void gn_row_sums(const __half* x1, long long ld1, int C1, const __half* x2, long long ld2, int c0,
                                            long long base, long long r_begin, long long r_end, int ty, int lanes,
                                            float (&s)[8], float (&q)[8]) {
    auto ld = [&](long long r) { const uint4* p = gn_src(x1, ld1, C1, x2, ld2, base + r, c0); return NC ? __ldg(p) : *p; };
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
        uint4 u0 = ld(r), u1 = ld(r + lanes), u2 = ld(r + 2LL * lanes), u3 = ld(r + 3LL * lanes);
        acc(u0); acc(u1); acc(u2); acc(u3);
    }
    for (; r < r_end; r += lanes) acc(ld(r));
}

// statistics only (split form: a reduction across GPUs follows): stats[batch][32][2] fp64 {sum, sumsq}, written by the last
// CTA of the chunk to arrive.  bars: [nbatch][2] u32 (zero at allocation), slots: [nbatch][CTAs per chunk][32] float2.
__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const __half* __restrict__ x1, long long ld1, int C1, const __half* __restrict__ x2, long long ld2,
                int C, long long rows_per_batch, int rows_per_cta, int vw, int lanes, double* __restrict__ stats,
                unsigned int* __restrict__ bars, float2* __restrict__ slots) {
    __shared__ float s_sum[GN_GROUPS], s_sq[GN_GROUPS];
    __shared__ float2 scr[GN_THREADS * 8];
    __shared__ double s_tot[2 * GN_GROUPS];
    const int t = threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    const int batch = blockIdx.y;
    const int tx = t % vw, ty = t / vw;
    const int vec0 = blockIdx.z * vw;
    const int c0 = (vec0 + tx) * 8;
    const int cpg = C / GN_GROUPS;
    const bool active = ty < lanes && c0 < C;
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
    const long long r_begin = (long long)blockIdx.x * rows_per_cta;
    long long r_end = r_begin + rows_per_cta;
    if (r_end > rows_per_batch) r_end = rows_per_batch;
    if (active) gn_row_sums<true>(x1, ld1, C1, x2, ld2, c0, (long long)batch * rows_per_batch, r_begin, r_end, ty, lanes, s, q);
    gn_cta_group_sums<GN_THREADS>(s, q, active, c0, cpg, vec0, min(vw, C / 8 - vec0), lanes, gn_slots_per_thread(C), scr, s_sum, s_sq);
    const int ncta = gridDim.x * gridDim.z;
    const int cta = blockIdx.x + gridDim.x * blockIdx.z;
    float2* chunk = slots + (long long)batch * ncta * GN_GROUPS;
    if (t < GN_GROUPS) chunk[(long long)cta * GN_GROUPS + t] = make_float2(s_sum[t], s_sq[t]);   // zeros for groups outside my slab
    if (!gn_barrier(bars + 2 * batch, (unsigned)ncta, false)) return;
    gn_sum_slots<GN_THREADS>(chunk, ncta, s_tot);
    if (t < 2 * GN_GROUPS) stats[(long long)batch * 2 * GN_GROUPS + t] = s_tot[t];
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

// Single-launch GroupNorm: statistics, a per-chunk arrival barrier, then apply -- the second read of x comes from L2
// instead of HBM and the two launches of the split path collapse into one graph node.  Every CTA of the grid must be
// co-resident: the host wrapper checks the grid against the occupancy of this kernel and refuses otherwise (the caller
// then uses the split kernels).  bars: [nbatch][2] u32 (zeroed once at allocation, self-resetting), slots: per-CTA partials.
__global__ void __launch_bounds__(GN_THREADS)
gn_fused_kernel(const __half* __restrict__ x1, long long ld1, int C1, const __half* __restrict__ x2, long long ld2,
                int C, long long rows_per_batch, int rows_per_cta, int vw, int lanes, unsigned int* __restrict__ bars,
                float2* __restrict__ slots, const float* __restrict__ gamma, const float* __restrict__ beta,
                float eps, int silu, __half* __restrict__ out, long long ldo) {
    __shared__ float s_sum[GN_GROUPS], s_sq[GN_GROUPS], s_mean[GN_GROUPS], s_rstd[GN_GROUPS];
    __shared__ float2 scr[GN_THREADS * 8];
    __shared__ double s_tot[2 * GN_GROUPS];
    const int t = threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    const int batch = blockIdx.y;
    const int tx = t % vw, ty = t / vw;
    const int vec0 = blockIdx.z * vw;
    const int c0 = (vec0 + tx) * 8;
    const int cpg = C / GN_GROUPS;
    const bool active = ty < lanes && c0 < C;
    const long long r_begin = (long long)blockIdx.x * rows_per_cta;
    long long r_end = r_begin + rows_per_cta;
    if (r_end > rows_per_batch) r_end = rows_per_batch;
    const long long base = (long long)batch * rows_per_batch;
    // ---- phase 1: partial sums of this CTA's rows (plain loads: `out` may alias x)
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
    if (active) gn_row_sums<false>(x1, ld1, C1, x2, ld2, c0, base, r_begin, r_end, ty, lanes, s, q);
    gn_cta_group_sums<GN_THREADS>(s, q, active, c0, cpg, vec0, min(vw, C / 8 - vec0), lanes, gn_slots_per_thread(C), scr, s_sum, s_sq);
    const int ncta = gridDim.x * gridDim.z;
    if (ncta > 1) {
        const int cta = blockIdx.x + gridDim.x * blockIdx.z;
        float2* chunk = slots + (long long)batch * ncta * GN_GROUPS;
        if (t < GN_GROUPS) chunk[(long long)cta * GN_GROUPS + t] = make_float2(s_sum[t], s_sq[t]);
        gn_barrier(bars + 2 * batch, (unsigned)ncta, true);
        gn_sum_slots<GN_THREADS>(chunk, ncta, s_tot);
    } else {
        if (t < GN_GROUPS) { s_tot[2 * t] = (double)s_sum[t]; s_tot[2 * t + 1] = (double)s_sq[t]; }
        __syncthreads();
    }
    // ---- phase 2: normalise; x is re-read (L2)
    if (t < GN_GROUPS) {
        const double inv_cnt = 1.0 / ((double)rows_per_batch * cpg);
        const double mean = s_tot[2 * t] * inv_cnt;
        double var = s_tot[2 * t + 1] * inv_cnt - mean * mean;
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
// gn_fused_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int GNS_THREADS = 512;
constexpr int GNS_MAX_DYN_SMEM = 227 * 1024 - 3072;       // the kernel's static smem (barrier, 4 x 32 floats, 64 doubles: 1.1 KB) comes on top
constexpr int GNS_CHUNK_BYTES = 16384;

// Cross-GPU part of a 5-D GroupNorm in the pixel-sharded layout (multi-GPU frame sharding, csrc/peer.cu): every rank holds
// all frames of HW/P pixels, so the statistics of a sample are the sum over ranks.  The first CTA of a chunk publishes this
// rank's 64 partial sums into every rank's slot and raises an epoch flag there; every CTA of the chunk waits for the P epochs
// in the local flag array and sums the P slots in rank order.  world <= 1: single-GPU behaviour.
// Control words: one 64-byte line per chunk with the same layout as every other peer op (csrc/peer.cu): u32 flags[8] at +0,
// epoch at +32.
struct GnPeer {
    int world, rank;
    double* slots[VMV_PEER_MAX_RANKS];            // [world][nbatch][64] doubles in every rank's arena
    unsigned int* flags[VMV_PEER_MAX_RANKS];      // [nbatch] lines of 16 u32 in every rank's arena (flags: first 8)
    unsigned int* epoch;                          // local: word 8 of line 0 (line b: + 16*b)
    long long stat_rows;                          // rows per chunk summed over all ranks
};

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(GNS_THREADS, 1)
gn_smem_kernel(const __half* __restrict__ x1, long long ld1, int C1, const __half* __restrict__ x2, long long ld2, int C2,
               long long rows_per_batch, int rows_per_cta, unsigned int* __restrict__ bars, float2* __restrict__ slots,
               const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu,
               __half* __restrict__ out, long long ldo, const GnPeer pe) {
    extern __shared__ __align__(128) uint8_t gsm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ float s_sum[GN_GROUPS], s_sq[GN_GROUPS], s_mean[GN_GROUPS], s_rstd[GN_GROUPS];
    __shared__ double s_tot[2 * GN_GROUPS];
    const int t = threadIdx.x;
    const int C = C1 + C2;
    const int batch = blockIdx.y;
    const long long r_begin = (long long)blockIdx.x * rows_per_cta;
    long long left = rows_per_batch - r_begin;
    const int nrows = left <= 0 ? 0 : (left < rows_per_cta ? (int)left : rows_per_cta);
    const long long base_row = (long long)batch * rows_per_batch + r_begin;
    float2* scr = reinterpret_cast<float2*>(gsm + (((size_t)rows_per_cta * C * 2 + 15) & ~(size_t)15));   // after the slab
    if (t == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    pdl_launch_dependents();
    pdl_wait();
    // the epoch this launch will use: read before this CTA arrives anywhere, i.e. before the publisher can advance it
    unsigned int peer_epoch = 0;
    if (pe.world > 1 && t == 0) peer_epoch = *(pe.epoch + batch * 16) + 1;
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
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
    if (active) {
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
    }
    gn_cta_group_sums<GNS_THREADS>(s, q, active, c0, cpg, 0, vw, lanes, gn_slots_per_thread(C), scr, s_sum, s_sq);
    const int ncta = gridDim.x;
    if (ncta > 1) {
        float2* chunk = slots + (long long)batch * ncta * GN_GROUPS;
        if (t < GN_GROUPS) chunk[(long long)blockIdx.x * GN_GROUPS + t] = make_float2(s_sum[t], s_sq[t]);
        gn_barrier(bars + 2 * batch, (unsigned)ncta, true);
        gn_sum_slots<GNS_THREADS>(chunk, ncta, s_tot);
    } else {                                                 // the chunk is mine alone: no global round trip
        if (t < GN_GROUPS) { s_tot[2 * t] = (double)s_sum[t]; s_tot[2 * t + 1] = (double)s_sq[t]; }
        __syncthreads();
    }
    if (pe.world > 1) {
        const int nb = gridDim.y;
        if (blockIdx.x == 0) {                               // publisher of this chunk
            if (t < GN_GROUPS) {
                const long long o = (((long long)pe.rank * nb + batch) * GN_GROUPS + t) * 2;
                for (int qq = 0; qq < pe.world; ++qq) { pe.slots[qq][o] = s_tot[2 * t]; pe.slots[qq][o + 1] = s_tot[2 * t + 1]; }
            }
            __syncthreads();
            if (t == 0) peer_publish(pe.flags, pe.world, pe.rank, peer_epoch, batch * 16);
        }
        if (t == 0) {
            peer_wait_all(pe.flags, pe.world, pe.rank, peer_epoch, batch * 16);
            if (blockIdx.x == 0) *(pe.epoch + batch * 16) = peer_epoch;
        }
        __syncthreads();
    }
    if (t < GN_GROUPS) {
        const double inv_cnt = 1.0 / ((double)(pe.world > 1 ? pe.stat_rows : rows_per_batch) * cpg);
        double sum, sq;
        if (pe.world > 1) {                                  // sum of the ranks' partials, in rank order
            const int nb = gridDim.y;
            sum = 0.0; sq = 0.0;
            for (int qq = 0; qq < pe.world; ++qq) {
                const long long o = (((long long)qq * nb + batch) * GN_GROUPS + t) * 2;
                sum += __ldcg(pe.slots[pe.rank] + o);
                sq += __ldcg(pe.slots[pe.rank] + o + 1);
            }
        } else {
            sum = s_tot[2 * t];
            sq = s_tot[2 * t + 1];
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

// Upper bound of the per-CTA partial-statistics slots one GroupNorm call may use (any of the three kernels).
static long long gn_max_ctas_per_call(int C, long long rows_per_batch, int nbatch) {
    GnGeom g = gn_geom(C, 0, rows_per_batch, nbatch);
    long long split = (long long)((rows_per_batch + g.rows_per_cta - 1) / g.rows_per_cta) * g.slabs * nbatch;
    long long resident = 160LL * 8 + nbatch;                     // fused / smem kernels never exceed the co-resident capacity
    return split > resident ? split : resident;
}

extern "C" int64_t vmv_groupnorm_scratch_bytes(int32_t C, int64_t rows_per_batch, int32_t nbatch) {
    if (C <= 0 || rows_per_batch <= 0 || nbatch <= 0) return -1;
    return gn_max_ctas_per_call(C, rows_per_batch, nbatch) * GN_GROUPS * (int64_t)sizeof(float2);
}

extern "C" int vmv_groupnorm_stats(const void* x1, int64_t ldx1, int32_t C1, const void* x2, int64_t ldx2, int32_t C2,
                                   int64_t rows_per_batch, int32_t nbatch, double* stats, void* barriers, void* scratch,
                                   void* stream) {
    int rc = gn_check("vmv_groupnorm_stats", x1, ldx1, C1, x2, ldx2, C2, rows_per_batch, nbatch);
    if (rc) return rc;
    VMV_CHECK_ARG(stats != nullptr && barriers != nullptr && scratch != nullptr, "vmv_groupnorm_stats: null stats/barriers/scratch");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GnGeom g = gn_geom(C1, C2, rows_per_batch, nbatch);
    dim3 grid((unsigned)((rows_per_batch + g.rows_per_cta - 1) / g.rows_per_cta), nbatch, g.slabs);
    launch_kernel(gn_stats_kernel, grid, dim3(GN_THREADS), 0, st, static_cast<const __half*>(x1), ldx1, C1,
                                                 static_cast<const __half*>(x2), ldx2, g.C, rows_per_batch,
                                                 g.rows_per_cta, g.vw, g.lanes, stats, static_cast<unsigned int*>(barriers),
                                                 static_cast<float2*>(scratch));
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

// Launch plan of the smem-resident single-pass kernel: CTAs per chunk, rows per CTA, dynamic smem.  One CTA per SM at most
// (all CTAs of the grid co-resident for the arrival barrier); small chunks use FEWER CTAs than SMs -- at least
// `min_kb` KB of rows each (VMV_GN_MIN_KB, default 32; profiles/r2_gn_bench.md) -- because below that the kernel is bound by the arrival barrier
// and the slot sum, whose cost grows with the number of CTAs of the chunk, not by moving the rows.
struct GnSmemPlan { long long cpb, rpc, smem; };

static bool gn_smem_plan(int C, long long rows_per_batch, int nbatch, GnSmemPlan* pl) {
    static int use_smem = -1, num_sms = 0, min_kb = 32;
    if (use_smem < 0) {
        const char* e = getenv("VMV_GN_SMEM");            // VMV_GN_SMEM=0: always the re-read kernel
        use_smem = (e && e[0] == '0') ? 0 : 1;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
    }
    {
        const char* e = getenv("VMV_GN_MIN_KB");          // read per call: the micro-benchmark sweeps it
        min_kb = e ? atoi(e) : 32;
    }
    if (!use_smem || C > 8 * GNS_THREADS || nbatch > num_sms) return false;
    long long cpb = num_sms / nbatch;                      // CTAs per chunk; all CTAs co-resident at one per SM
    if (cpb > rows_per_batch) cpb = rows_per_batch;
    if (min_kb > 0) {
        long long by_bytes = rows_per_batch * C * 2 / ((long long)min_kb * 1024);
        if (by_bytes < 1) by_bytes = 1;
        if (cpb > by_bytes) cpb = by_bytes;
    }
    const long long rpc = (rows_per_batch + cpb - 1) / cpb;
    cpb = (rows_per_batch + rpc - 1) / rpc;                // drop CTAs that would own no rows
    // the slab + the reduction scratch (GNS_THREADS x slots-per-thread float2)
    const long long smem = ((rpc * C * 2 + 15) & ~15LL) + (long long)GNS_THREADS * gn_slots_per_thread(C) * 8;
    if (smem > GNS_MAX_DYN_SMEM) return false;
    pl->cpb = cpb; pl->rpc = rpc; pl->smem = smem;
    return true;
}

extern "C" int vmv_groupnorm_fused_fits_smem(int32_t C, int64_t rows_per_batch, int32_t nbatch) {
    GnSmemPlan pl;
    return (C > 0 && rows_per_batch > 0 && nbatch > 0 && gn_smem_plan(C, rows_per_batch, nbatch, &pl)) ? 1 : 0;
}

static int gn_fused_impl(const void* x1, int64_t ldx1, int32_t C1, const void* x2, int64_t ldx2, int32_t C2,
                         int64_t rows_per_batch, int32_t nbatch, void* barriers, void* scratch, const float* gamma,
                         const float* beta, float eps, int32_t silu, void* out, int64_t ldo, const vmv_gn_peer* peer,
                         void* stream) {
    int rc = gn_check("vmv_groupnorm_fused", x1, ldx1, C1, x2, ldx2, C2, rows_per_batch, nbatch);
    if (rc) return rc;
    VMV_CHECK_ARG(barriers && scratch && gamma && beta && out && ldo % 8 == 0 && ldo >= C1 + C2,
                  "vmv_groupnorm_fused: bad barriers/scratch/gamma/beta/out");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned int* bars = static_cast<unsigned int*>(barriers);
    float2* slots = static_cast<float2*>(scratch);
    {
        const int C = C1 + C2;
        const bool al16 = ((reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(x2) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
        GnSmemPlan sp;
        if (al16 && gn_smem_plan(C, rows_per_batch, nbatch, &sp)) {
            static bool attr_set = false;
            if (!attr_set) {
                cudaError_t e2 = cudaFuncSetAttribute(gn_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GNS_MAX_DYN_SMEM);
                if (e2 != cudaSuccess) { set_error("vmv_groupnorm_fused: smem attribute: %s", cudaGetErrorString(e2)); return VMV_ERR_CUDA; }
                attr_set = true;
            }
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
            launch_kernel(gn_smem_kernel, dim3((unsigned)sp.cpb, nbatch), dim3(GNS_THREADS), (size_t)sp.smem, st,
                          static_cast<const __half*>(x1), ldx1, C1, static_cast<const __half*>(x2), ldx2, C2, rows_per_batch,
                          (int)sp.rpc, bars, slots, gamma, beta, eps, silu, static_cast<__half*>(out), ldo, pe);
            count_launch();
            VMV_CUDA_LAUNCH_CHECK("vmv_groupnorm_fused (smem)");
            return VMV_OK;
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
        if (per_sm > 8) per_sm = 8;                            // vmv_groupnorm_scratch_bytes assumes <= 8 per SM
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
                                                 g.rows_per_cta, g.vw, g.lanes, bars, slots, gamma, beta, eps, silu,
                                                 static_cast<__half*>(out), ldo);
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_groupnorm_fused");
    return VMV_OK;
}

extern "C" int vmv_groupnorm_fused(const void* x1, int64_t ldx1, int32_t C1, const void* x2, int64_t ldx2, int32_t C2,
                                   int64_t rows_per_batch, int32_t nbatch, void* barriers, void* scratch, const float* gamma,
                                   const float* beta, float eps, int32_t silu, void* out, int64_t ldo, void* stream) {
    return gn_fused_impl(x1, ldx1, C1, x2, ldx2, C2, rows_per_batch, nbatch, barriers, scratch, gamma, beta, eps, silu, out, ldo, nullptr, stream);
}

extern "C" int vmv_groupnorm_fused_peer(const void* x1, int64_t ldx1, int32_t C1, const void* x2, int64_t ldx2, int32_t C2,
                                        int64_t rows_per_batch, int32_t nbatch, void* barriers, void* scratch, const float* gamma,
                                        const float* beta, float eps, int32_t silu, void* out, int64_t ldo,
                                        const vmv_gn_peer* peer, void* stream) {
    VMV_CHECK_ARG(peer && peer->world >= 1 && peer->world <= VMV_PEER_MAX_RANKS && peer->rank >= 0 && peer->rank < peer->world,
                  "vmv_groupnorm_fused_peer: bad world/rank");
    VMV_CHECK_ARG(peer->epoch && peer->stat_rows >= rows_per_batch, "vmv_groupnorm_fused_peer: bad epoch/stat_rows");
    for (int q = 0; q < peer->world; ++q)
        VMV_CHECK_ARG(peer->slots[q] && peer->flags[q], "vmv_groupnorm_fused_peer: null slots/flags for rank %d", q);
    return gn_fused_impl(x1, ldx1, C1, x2, ldx2, C2, rows_per_batch, nbatch, barriers, scratch, gamma, beta, eps, silu, out, ldo, peer, stream);
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
