// Fused attention  O = softmax(Q K^T * scale) V,  head_dim = 64, fp16 in/out, fp32 online softmax.
//
// One CTA = NWARPS*16 query rows of one (batch, head); K/V streamed in BC-key chunks through a
// double-buffered cp.async ring; S and O accumulators live in registers (mma.sync m16n8k16), P never
// leaves registers (the S accumulator layout is the A-fragment layout of the P.V MMA).
// Q/K/V/O are addressed with (outer, inner, row) strides so the kernel reads the fused-QKV GEMM output
// in place for spatial attention (rows = pixels of a frame), text cross-attention (K/V shared by the
// frames of a sample) and temporal attention (rows = frames of a pixel) -- no head split / merge or
// NCHW<->token transposes (reference util.py:237-244,262-267,1054-1083).
//
// This mma.sync kernel serves the strided / short-row cases (temporal attention over 24 frames, < 128 query rows);
// spatial self-attention and text cross-attention are dispatched to the persistent tcgen05/TMEM kernel in
// attention_tc.cu (profiles/r1_attention_tc_vs_mma.log).
#include "common.cuh"
#include "../../include/videomv_b200.h"

#include <stdlib.h>

namespace vmv {

void count_launch(int n = 1);

constexpr int HD = 64;   // head dim

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
    const int sz = pred ? 16 : 0;   // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// smem tile: rows of 64 halfs (128 B = 8 x 16B chunks); chunk index XOR-swizzled by (row & 7)
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) { return (uint32_t)((row * 8 + (chunk ^ (row & 7))) * 16); }

template <int NWARPS, int BC>
__global__ void __launch_bounds__(NWARPS * 32)
attention_kernel(const vmv_attn_params p) {
    constexpr int BR = NWARPS * 16;
    constexpr int NT = NWARPS * 32;
    __shared__ __align__(128) uint8_t sQ[BR * 128];
    __shared__ __align__(128) uint8_t sK[2][BC * 128];
    __shared__ __align__(128) uint8_t sV[2][BC * 128];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int bo = b / p.inner, bi = b % p.inner;
    const int kbo = bo / p.kv_group;
    const __half* qp = static_cast<const __half*>(p.q) + bo * p.q_bs_outer + bi * p.q_bs_inner + h * HD;
    const __half* kp = static_cast<const __half*>(p.k) + kbo * p.k_bs_outer + bi * p.k_bs_inner + h * HD;
    const __half* vp = static_cast<const __half*>(p.v) + kbo * p.v_bs_outer + bi * p.v_bs_inner + h * HD;
    __half* op = static_cast<__half*>(p.o) + bo * p.o_bs_outer + bi * p.o_bs_inner + h * HD;
    const int q0 = qt * BR;
    pdl_launch_dependents();
    pdl_wait();

    // ---- stage Q (once) and the first K/V chunk
    for (int i = tid; i < BR * 8; i += NT) {
        const int r = i >> 3, c = i & 7;
        const bool ok = (q0 + r) < p.nq;
        cp_async16(sQ + tile_off(r, c), qp + (long long)(ok ? q0 + r : 0) * p.q_rs + c * 8, ok);
    }
    auto load_kv = [&](int chunk, int buf) {
        const int k0 = chunk * BC;
        for (int i = tid; i < BC * 8; i += NT) {
            const int r = i >> 3, c = i & 7;
            const bool ok = (k0 + r) < p.nk;
            const long long rr = ok ? k0 + r : 0;
            cp_async16(sK[buf] + tile_off(r, c), kp + rr * p.k_rs + c * 8, ok);
            cp_async16(sV[buf] + tile_off(r, c), vp + rr * p.v_rs + c * 8, ok);
        }
    };
    load_kv(0, 0);
    cp_async_commit();

    const int nchunks = (p.nk + BC - 1) / BC;
    const float sl2 = p.scale * 1.4426950408889634f;   // scale * log2(e)

    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    uint32_t qf[4][4];

    const int r0 = warp * 16;
    const int g = lane >> 2, qd = lane & 3;

    for (int ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunks) {
            load_kv(ch + 1, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (ch == 0) {
            // Q fragments: A operand, 16 rows x (4 k-steps of 16)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int m = lane >> 3;
                const int row = r0 + (m & 1) * 8 + (lane & 7);
                const int chunk = ks * 2 + (m >> 1);
                ldsm_x4(smem_u32(sQ + tile_off(row, chunk)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
            }
        }
        // ---- S = Q K^T  (16 x BC per warp)
        float s[BC / 8][4];
#pragma unroll
        for (int j = 0; j < BC / 8; ++j) {
            s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                uint32_t b0, b1, b2, b3;
                const int m = lane >> 3;
                const int row = j * 8 + (lane & 7);
                ldsm_x4(smem_u32(sK[buf] + tile_off(row, 4 * t + m)), b0, b1, b2, b3);
                mma_16816(s[j], qf[2 * t], b0, b1);
                mma_16816(s[j], qf[2 * t + 1], b2, b3);
            }
        }
        // ---- mask + online softmax (rows g and g+8 of this warp's 16)
        const int kbase = ch * BC;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < BC / 8; ++j) {
            const int key = kbase + j * 8 + qd * 2;
            if (key >= p.nk) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
            if (key + 1 >= p.nk) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
            mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
            mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
        }
        float corr[2], msc[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float m_new = fmaxf(m_run[i], mx[i]);       // finite: chunk 0 always holds key 0
            corr[i] = ex2_approx_f((m_run[i] - m_new) * sl2);
            m_run[i] = m_new;
            msc[i] = m_new * sl2;
            l_run[i] *= corr[i];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o[j][0] *= corr[0]; o[j][1] *= corr[0];
            o[j][2] *= corr[1]; o[j][3] *= corr[1];
        }
        uint32_t pf[BC / 16][4];
#pragma unroll
        for (int j = 0; j < BC / 8; ++j) {
            const float p0 = ex2_approx_f(s[j][0] * sl2 - msc[0]);
            const float p1 = ex2_approx_f(s[j][1] * sl2 - msc[0]);
            const float p2 = ex2_approx_f(s[j][2] * sl2 - msc[1]);
            const float p3 = ex2_approx_f(s[j][3] * sl2 - msc[1]);
            l_run[0] += p0 + p1;
            l_run[1] += p2 + p3;
            pf[j >> 1][(j & 1) * 2] = pack_half2(p0, p1);
            pf[j >> 1][(j & 1) * 2 + 1] = pack_half2(p2, p3);
        }
        // ---- O += P V
#pragma unroll
        for (int kk = 0; kk < BC / 16; ++kk) {
#pragma unroll
            for (int jd = 0; jd < 8; jd += 2) {
                uint32_t b0, b1, b2, b3;
                const int m = lane >> 3;
                const int row = kk * 16 + (m & 1) * 8 + (lane & 7);
                ldsm_x4_t(smem_u32(sV[buf] + tile_off(row, jd + (m >> 1))), b0, b1, b2, b3);
                mma_16816(o[jd], pf[kk], b0, b1);
                mma_16816(o[jd + 1], pf[kk], b2, b3);
            }
        }
        __syncthreads();   // everyone done with sK/sV[buf] before it is refilled
    }

    // ---- finalise: O /= l, stage through sQ (this warp's own 16 rows), 16B coalesced stores
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        l_run[i] += __shfl_xor_sync(0xffffffffu, l_run[i], 1);
        l_run[i] += __shfl_xor_sync(0xffffffffu, l_run[i], 2);
    }
    const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int colb = (j * 8 + qd * 2) * 2;   // byte offset inside the 128B row
        const int chunk = colb >> 4, within = colb & 15;
        *reinterpret_cast<uint32_t*>(sQ + tile_off(r0 + g, chunk) + within) = pack_half2(o[j][0] * inv0, o[j][1] * inv0);
        *reinterpret_cast<uint32_t*>(sQ + tile_off(r0 + g + 8, chunk) + within) = pack_half2(o[j][2] * inv1, o[j][3] * inv1);
    }
    __syncwarp();
    for (int i = lane; i < 16 * 8; i += 32) {
        const int r = i >> 3, c = i & 7;
        const int qrow = q0 + r0 + r;
        if (qrow < p.nq) {
            uint4 u = *reinterpret_cast<const uint4*>(sQ + tile_off(r0 + r, c));
            *reinterpret_cast<uint4*>(op + (long long)qrow * p.o_rs + c * 8) = u;
        }
    }
}

int attention_tc_try(const vmv_attn_params* p, cudaStream_t st);   // attention_tc.cu

}  // namespace vmv

using namespace vmv;

extern "C" int vmv_attention(const vmv_attn_params* p, void* stream) {
    VMV_CHECK_ARG(p && p->q && p->k && p->v && p->o, "vmv_attention: null pointer");
    VMV_CHECK_ARG(p->outer > 0 && p->inner > 0 && p->heads > 0 && p->nq > 0 && p->nk > 0, "vmv_attention: bad sizes");
    VMV_CHECK_ARG(p->kv_group > 0, "vmv_attention: kv_group must be >= 1");
    const int64_t strides[] = {p->q_bs_outer, p->q_bs_inner, p->q_rs, p->k_bs_outer, p->k_bs_inner, p->k_rs,
                               p->v_bs_outer, p->v_bs_inner, p->v_rs, p->o_bs_outer, p->o_bs_inner, p->o_rs};
    for (int64_t s : strides) VMV_CHECK_ARG(s % 8 == 0, "vmv_attention: strides must be multiples of 8 elements");
    const long long nb = (long long)p->outer * p->inner;
    VMV_CHECK_ARG(nb <= 65535 * 1LL && p->heads <= 65535, "vmv_attention: batch %lld too large for one launch", nb);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        // tcgen05 kernel whenever the layout allows (long contiguous sequences, packed short ones, packed strided ones:
        // every attention of the UNet), or explicitly with impl == 2.  VMV_ATTN_TC=0 keeps auto on the mma.sync kernel.
        static int auto_tc = -1;
        if (auto_tc < 0) { const char* e = getenv("VMV_ATTN_TC"); auto_tc = (e && e[0] == '0') ? 0 : 1; }
        if (p->impl == 2 || (p->impl == 0 && auto_tc)) {
            const int rc = attention_tc_try(p, st);
            if (rc != VMV_ERR_UNSUPPORTED) return rc;
            if (p->impl == 2) {
                set_error("vmv_attention: impl=2 (tcgen05) needs contiguous batches (inner == 1, batch stride == n * row stride) of "
                          ">= 128 rows or of a power-of-two length that divides 128, or strided sequences of <= 128 rows "
                          "(inner > 1, nq == nk), and 32 B aligned output rows");
                return VMV_ERR_UNSUPPORTED;
            }
        }
    }
    if (p->nq <= 32 && p->nk <= 32) {
        dim3 grid((p->nq + 31) / 32, p->heads, (unsigned)nb);
        launch_kernel(attention_kernel<2, 32>, grid, dim3(64), 0, st, *p);
    } else {
        dim3 grid((p->nq + 63) / 64, p->heads, (unsigned)nb);
        launch_kernel(attention_kernel<4, 64>, grid, dim3(128), 0, st, *p);
    }
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_attention");
    return VMV_OK;
}
