// Small CUDA-core kernels around the tensor-core path: data movement (nearest upsample, stride-2 patch gather),
// the 4-channel input / output convolutions, embeddings, and the CFG + DDIM update.
#include "common.cuh"
#include "../../include/videomv_b200.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <atomic>

namespace vmv {

// ------------------------------------------------------------------------------------------------
// error / accounting plumbing shared by all translation units
// ------------------------------------------------------------------------------------------------
void count_launch(int n = 1);
static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VMV_PDL");       // VMV_PDL=0: plain stream-ordered launches (A/B switch)
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void upsample2x_kernel(const uint4* __restrict__ x, int H, int W, int cv, uint4* __restrict__ out,
                                  long long total) {
    pdl_launch_dependents();
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % cv);
    long long r = i / cv;
    const int ow = (int)(r % (2 * W)); r /= (2 * W);
    const int oh = (int)(r % (2 * H));
    const long long n = r / (2 * H);
    out[i] = __ldg(x + ((n * H + (oh >> 1)) * W + (ow >> 1)) * cv + c);
}

// out[(n,oh,ow), (ky,kx,c)] = x[n, 2*oh+ky-1, 2*ow+kx-1, c]  (zero outside)
__global__ void im2col_s2_kernel(const uint4* __restrict__ x, int H, int W, int cv, uint4* __restrict__ out,
                                 long long total) {
    pdl_launch_dependents();
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % cv);
    long long r = i / cv;
    const int tap = (int)(r % 9); r /= 9;
    const int OW = W / 2, OH = H / 2;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const long long n = r / OH;
    const int ih = 2 * oh + tap / 3 - 1, iw = 2 * ow + tap % 3 - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = __ldg(x + ((n * H + ih) * W + iw) * cv + c);
    out[i] = v;
}

// Input conv: tiny Cin (4 or 8), K = 9*Cin <= 72, fp32 in, fp32 accumulate.  A CTA owns IN_PIX consecutive pixels.
//   * weights: read coalesced in the reference layout [Cout][K] and scattered to smem transposed [k][Cout], so a thread
//     reads the 8 output channels it owns as two 16 B loads per k;
//   * the K input taps of the CTA's pixels are gathered once into smem as [k][pixel], so 4 consecutive pixels are one
//     16 B load;
//   * a thread computes 4 pixels x 8 channels per work item: 3 LDS.128 per 32 FMAs (the 1 x 8 version was shared-memory
//     bound: 200 us for the 24 x 32 x 32 x 4 -> 320 stem).
constexpr int IN_PIX = 128;
constexpr int IN_MAXK = 72;
__global__ void __launch_bounds__(256)
conv3x3_in_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2, int C2, int B, int F, int H, int W,
                  const float* __restrict__ w, const float* __restrict__ bias, int Cout, __half* __restrict__ out) {
    extern __shared__ __align__(16) float sm[];
    const int Cin = C1 + C2, K = 9 * Cin;
    float* sw = sm;                         // [K][Cout]
    float* sx = sm + K * Cout;              // [K][IN_PIX]
    const long long npix = (long long)B * F * H * W;
    const long long pix0 = (long long)blockIdx.x * IN_PIX;
    for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) {
        const int co = i / K, k = i - co * K;           // k = (ci*3 + ky)*3 + kx, matching w[co][ci][ky][kx]
        sw[k * Cout + co] = __ldg(w + i);
    }
    pdl_launch_dependents();                // the weights above are static: staged while the previous kernel drains
    pdl_wait();
    for (int i = threadIdx.x; i < IN_PIX * K; i += blockDim.x) {
        const int k = i / IN_PIX, p = i - k * IN_PIX;   // consecutive threads = consecutive pixels of one tap
        const long long pix = pix0 + p;
        float v = 0.f;
        if (pix < npix) {
            const int xw = (int)(pix % W), yh = (int)((pix / W) % H);
            const int f = (int)((pix / ((long long)W * H)) % F), b = (int)(pix / ((long long)W * H * F));
            const int ci = k / 9, ky = (k % 9) / 3, kx = k % 3;
            const int ih = yh + ky - 1, iw = xw + kx - 1;
            if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
                const float* src = ci < C1 ? x1 + (((long long)b * C1 + ci) * F + f) * H * W
                                           : x2 + (((long long)b * C2 + (ci - C1)) * F + f) * H * W;
                v = __ldg(src + ih * W + iw);
            }
        }
        sx[i] = v;
    }
    __syncthreads();
    const int cov = Cout / 8;
    for (int i = threadIdx.x; i < (IN_PIX / 4) * cov; i += blockDim.x) {
        // channels of a work item: 4 from each half of Cout, so that both weight loads are lane-contiguous 16 B
        // (conflict-free) instead of 16 B out of every 32 B
        const int pq = i / cov, ca = (i - pq * cov) * 4, cb = Cout / 2 + ca;
        float acc[4][8];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[q][j] = 0.f;
        for (int k = 0; k < K; ++k) {
            const float4 xv = *reinterpret_cast<const float4*>(sx + k * IN_PIX + pq * 4);
            const float4 w0 = *reinterpret_cast<const float4*>(sw + k * Cout + ca);
            const float4 w1 = *reinterpret_cast<const float4*>(sw + k * Cout + cb);
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
            const float ws[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[q][j] = fmaf(xs[q], ws[j], acc[q][j]);
        }
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + ca));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + cb));
        const float bs[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long long pix = pix0 + pq * 4 + q;
            if (pix >= npix) continue;
            *reinterpret_cast<uint2*>(out + pix * Cout + ca) =
                make_uint2(pack_half2(acc[q][0] + bs[0], acc[q][1] + bs[1]), pack_half2(acc[q][2] + bs[2], acc[q][3] + bs[3]));
            *reinterpret_cast<uint2*>(out + pix * Cout + cb) =
                make_uint2(pack_half2(acc[q][4] + bs[4], acc[q][5] + bs[5]), pack_half2(acc[q][6] + bs[6], acc[q][7] + bs[7]));
        }
    }
}

// Head conv: tiny Cout (4).  One warp per output pixel, lanes stride the (tap, 8-channel vector) reduction; the weights
// are staged once per CTA in smem as [tap][c][COUT] so each (tap, c) is one 16B broadcast-free load.
constexpr int OUT_PIX_PER_WARP = 8;
template <int COUT>
__global__ void __launch_bounds__(256)
conv3x3_out_kernel(const __half* __restrict__ x, int B, int F, int H, int W, int C, const float* __restrict__ w,
                   const float* __restrict__ bias, float* __restrict__ out) {
    extern __shared__ float sm[];                       // [9][C][COUT]
    for (int i = threadIdx.x; i < 9 * C * COUT; i += blockDim.x) {
        const int co = i % COUT, c = (i / COUT) % C, tap = i / (COUT * C);
        sm[i] = w[((long long)co * C + c) * 9 + tap];
    }
    pdl_launch_dependents();
    pdl_wait();
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long npix = (long long)B * F * H * W;
    const int cv = C / 8;
    for (int pp = 0; pp < OUT_PIX_PER_WARP; ++pp) {
        const long long pix = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * OUT_PIX_PER_WARP + pp;
        if (pix >= npix) break;
        const int xw = (int)(pix % W);
        const int yh = (int)((pix / W) % H);
        const long long n = pix / ((long long)W * H);     // b*F + f
        float acc[COUT];
#pragma unroll
        for (int j = 0; j < COUT; ++j) acc[j] = 0.f;
        for (int i = lane; i < 9 * cv; i += 32) {
            const int tap = i / cv, c0 = (i % cv) * 8;
            const int ih = yh + tap / 3 - 1, iw = xw + tap % 3 - 1;
            if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
            uint4 u = __ldg(reinterpret_cast<const uint4*>(x + ((n * H + ih) * W + iw) * C + c0));
            uint32_t ww[4] = {u.x, u.y, u.z, u.w};
            const float* wp = sm + ((long long)tap * C + c0) * COUT;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 f2 = unpack_half2(ww[j]);
#pragma unroll
                for (int co = 0; co < COUT; ++co) {
                    acc[co] += f2.x * wp[(2 * j) * COUT + co];
                    acc[co] += f2.y * wp[(2 * j + 1) * COUT + co];
                }
            }
        }
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[co] += __shfl_xor_sync(0xffffffffu, acc[co], o);
        }
        if (lane == 0) {
            const int f = (int)(n % F);
            const long long b = n / F;
#pragma unroll
            for (int co = 0; co < COUT; ++co)
                out[(((b * COUT + co) * F + f) * H + yh) * W + xw] = acc[co] + bias[co];
        }
    }
}

// [M, ld] fp16 rows (first Cout columns valid) -> fp32 NCFHW [B,Cout,F,H,W]   (head conv output, unet_t2v.py:368)
__global__ void rows_to_ncfhw_kernel(const __half* __restrict__ x, long long ld, int B, int F, long long HW, int Cout,
                                     float* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over B*Cout*F*HW outputs
    const long long total = (long long)B * Cout * F * HW;
    if (i >= total) return;
    const long long p = i % HW;
    const int f = (int)((i / HW) % F);
    const int co = (int)((i / (HW * F)) % Cout);
    const long long b = i / (HW * F * Cout);
    out[i] = __half2float(x[((b * F + f) * HW + p) * ld + co]);
}

__global__ void sinusoidal_kernel(const long long* __restrict__ t, int B, int dim, __half* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const int half_dim = dim / 2;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * half_dim) return;
    const int b = i / half_dim, k = i % half_dim;
    const float tv = (float)t[b];
    // torch.pow(10000, -arange(half)/half) evaluated in fp32 (util.py:184-186)
    const float freq = powf(10000.0f, -(float)k / (float)half_dim);
    const float ang = tv * freq;
    out[(long long)b * dim + k] = __float2half_rn(cosf(ang));
    out[(long long)b * dim + half_dim + k] = __float2half_rn(sinf(ang));
}

__global__ void embed_combine_silu_kernel(const __half* __restrict__ te, const __half* __restrict__ te2,
                                          const __half* __restrict__ cam, int B, int F, int E, __half* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * F * E) return;
    const int e = (int)(i % E);
    const long long bf = i / E;
    const long long b = bf / F;
    float v = __half2float(te[b * E + e]);
    if (te2) v += __half2float(te2[b * E + e]);
    if (cam) v += __half2float(cam[bf * E + e]);
    out[i] = __float2half_rn(silu_f(v));
}

__global__ void cfg_ddim_kernel(const float* __restrict__ xt, const float* __restrict__ y, const float* __restrict__ u,
                                const float* __restrict__ coef, long long n, float* __restrict__ xprev) {
    pdl_launch_dependents();
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float kx = coef[0], ko = coef[1], c_recip = coef[2], c_recipm1 = coef[3], sa_prev = coef[4],
                s1a_prev = coef[5], gs = coef[6];
    const float uu = u[i];
    const float out = uu + gs * (y[i] - uu);                       // diffusion_ddim.py:157-160
    const float x0 = kx * xt[i] - ko * out;                        // :193-199
    const float eps = (c_recip * xt[i] - x0) / c_recipm1;          // :233-234
    xprev[i] = sa_prev * x0 + s1a_prev * eps;                      // :240-243 (eta = 0)
}


// softmax over the last dimension of fp16 rows (fp32 maths): out[m, :] = softmax(scale * x[m, :]).  One warp per row, three
// passes over the row (max, sum of exponentials, normalised write) with 16 B accesses; the row stays in L1/L2 between them.
// Used by the single-head d=512 attention of the VAE decoder's middle block (autoencoder.py:419-441).
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const __half* __restrict__ x, long long ldx, long long M, int N, float scale, __half* __restrict__ out, long long ldo) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const uint4* xp = reinterpret_cast<const uint4*>(x + row * ldx);
    const int nvec = N / 8;
    const float sl2 = scale * 1.4426950408889634f;
    float mx = -INFINITY;
    for (int v = lane; v < nvec; v += 32) {
        const uint4 u = xp[v];
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = unpack_half2(w[j]); mx = fmaxf(mx, fmaxf(f.x, f.y)); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    // scale > 0: max of scale*x = scale*max
    const float off = mx * sl2;
    float sum = 0.f;
    for (int v = lane; v < nvec; v += 32) {
        const uint4 u = xp[v];
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = unpack_half2(w[j]); sum += exp2f(fmaf(f.x, sl2, -off)) + exp2f(fmaf(f.y, sl2, -off)); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    uint4* op = reinterpret_cast<uint4*>(out + row * ldo);
    for (int v = lane; v < nvec; v += 32) {
        const uint4 u = xp[v];
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_half2(w[j]);
            o[j] = pack_half2(exp2f(fmaf(f.x, sl2, -off)) * inv, exp2f(fmaf(f.y, sl2, -off)) * inv);
        }
        op[v] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

}  // namespace vmv

using namespace vmv;

extern "C" const char* vmv_last_error(void) { return g_err; }
extern "C" int vmv_abi_version(void) { return 6; }
extern "C" long long vmv_launch_count(void) { return g_launches.load(); }
extern "C" int vmv_sizeof_gemm_params(void) { return (int)sizeof(vmv_gemm_params); }
extern "C" int vmv_sizeof_attn_params(void) { return (int)sizeof(vmv_attn_params); }
extern "C" int vmv_sizeof_peer_exchange_params(void) { return (int)sizeof(vmv_peer_exchange_params); }
extern "C" int vmv_sizeof_peer_allreduce_params(void) { return (int)sizeof(vmv_peer_allreduce_params); }
extern "C" int vmv_sizeof_gemm_scatter(void) { return (int)sizeof(vmv_gemm_scatter); }
extern "C" int vmv_sizeof_gn_peer(void) { return (int)sizeof(vmv_gn_peer); }
extern "C" int vmv_sizeof_peer_allgather_params(void) { return (int)sizeof(vmv_peer_allgather_params); }

extern "C" int vmv_upsample_nearest2x(const void* x, int32_t n, int32_t H, int32_t W, int32_t C, void* out, void* stream) {
    VMV_CHECK_ARG(x && out && n > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "vmv_upsample_nearest2x: bad args");
    const long long total = (long long)n * 4 * H * W * (C / 8);
    launch_kernel(upsample2x_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
        static_cast<const uint4*>(x), H, W, C / 8, static_cast<uint4*>(out), total);
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_upsample_nearest2x");
    return VMV_OK;
}

extern "C" int vmv_im2col_3x3_s2(const void* x, int32_t n, int32_t H, int32_t W, int32_t C, void* out, void* stream) {
    VMV_CHECK_ARG(x && out && n > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "vmv_im2col_3x3_s2: bad args");
    const long long total = (long long)n * (H / 2) * (W / 2) * 9 * (C / 8);
    launch_kernel(im2col_s2_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
        static_cast<const uint4*>(x), H, W, C / 8, static_cast<uint4*>(out), total);
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_im2col_3x3_s2");
    return VMV_OK;
}

extern "C" int vmv_conv3x3_in(const float* x1, int32_t C1, const float* x2, int32_t C2, int32_t B, int32_t F, int32_t H,
                              int32_t W, const float* w, const float* bias, int32_t Cout, void* out, void* stream) {
    VMV_CHECK_ARG(x1 && w && bias && out && C1 > 0 && (C2 == 0 || x2) && Cout % 8 == 0, "vmv_conv3x3_in: bad args");
    const int K = 9 * (C1 + C2);
    VMV_CHECK_ARG(K <= IN_MAXK, "vmv_conv3x3_in: Cin=%d too large for the stem kernel (max %d)", C1 + C2, IN_MAXK / 9);
    const long long npix = (long long)B * F * H * W;
    const size_t smem = sizeof(float) * ((size_t)K * Cout + (size_t)IN_PIX * K);
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("vmv_conv3x3_in: smem attribute: %s", cudaGetErrorString(e)); return VMV_ERR_CUDA; }
        smem_set = smem;
    }
    launch_kernel(conv3x3_in_kernel, dim3((unsigned)((npix + IN_PIX - 1) / IN_PIX)), dim3(256), smem, static_cast<cudaStream_t>(stream),
        x1, C1, x2, C2, B, F, H, W, w, bias, Cout, static_cast<__half*>(out));
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_conv3x3_in");
    return VMV_OK;
}

extern "C" int vmv_conv3x3_out(const void* x, int32_t B, int32_t F, int32_t H, int32_t W, int32_t C, const float* w,
                               const float* bias, int32_t Cout, float* out, void* stream) {
    VMV_CHECK_ARG(x && w && bias && out && C % 8 == 0, "vmv_conv3x3_out: bad args");
    VMV_CHECK_ARG(Cout == 4, "vmv_conv3x3_out: only out_dim=4 is instantiated (got %d)", Cout);
    const long long npix = (long long)B * F * H * W;
    const size_t smem = sizeof(float) * 9 * (size_t)C * 4;
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_out_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("vmv_conv3x3_out: smem attribute: %s", cudaGetErrorString(e)); return VMV_ERR_CUDA; }
        smem_set = smem;
    }
    const int pix_per_cta = 8 * OUT_PIX_PER_WARP;
    launch_kernel(conv3x3_out_kernel<4>, dim3((unsigned)((npix + pix_per_cta - 1) / pix_per_cta)), dim3(256), smem, static_cast<cudaStream_t>(stream),
        static_cast<const __half*>(x), B, F, H, W, C, w, bias, out);
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_conv3x3_out");
    return VMV_OK;
}

extern "C" int vmv_softmax_rows(const void* x, int64_t ldx, int64_t M, int32_t N, float scale, void* out, int64_t ldo, void* stream) {
    VMV_CHECK_ARG(x && out && M > 0 && N > 0 && N % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && ldx >= N && ldo >= N && scale > 0.f,
                  "vmv_softmax_rows: bad args (N, ldx, ldo multiples of 8; scale > 0)");
    VMV_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "vmv_softmax_rows: x/out must be 16 B aligned");
    launch_kernel(softmax_rows_kernel, dim3((unsigned)((M + 7) / 8)), dim3(256), 0, static_cast<cudaStream_t>(stream),
                  static_cast<const __half*>(x), ldx, M, N, scale, static_cast<__half*>(out), ldo);
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_softmax_rows");
    return VMV_OK;
}

extern "C" int vmv_rows_to_ncfhw(const void* x, int64_t ldx, int32_t B, int32_t F, int32_t H, int32_t W, int32_t Cout,
                                 float* out, void* stream) {
    VMV_CHECK_ARG(x && out && B > 0 && F > 0 && H > 0 && W > 0 && Cout > 0 && ldx >= Cout, "vmv_rows_to_ncfhw: bad args");
    const long long total = (long long)B * Cout * F * H * W;
    launch_kernel(rows_to_ncfhw_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
        static_cast<const __half*>(x), ldx, B, F, (long long)H * W, Cout, out);
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_rows_to_ncfhw");
    return VMV_OK;
}

extern "C" int vmv_sinusoidal_embedding(const int64_t* t, int32_t B, int32_t dim, void* out, void* stream) {
    VMV_CHECK_ARG(t && out && B > 0 && dim > 0 && dim % 2 == 0, "vmv_sinusoidal_embedding: bad args");
    const int total = B * (dim / 2);
    launch_kernel(sinusoidal_kernel, dim3((total + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream),
        reinterpret_cast<const long long*>(t), B, dim, static_cast<__half*>(out));
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_sinusoidal_embedding");
    return VMV_OK;
}

extern "C" int vmv_embed_combine_silu(const void* t_emb, const void* t_emb2, const void* cam_emb, int32_t B, int32_t F,
                                      int32_t E, void* out, void* stream) {
    VMV_CHECK_ARG(t_emb && out && B > 0 && F > 0 && E > 0, "vmv_embed_combine_silu: bad args");
    const long long total = (long long)B * F * E;
    launch_kernel(embed_combine_silu_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
        static_cast<const __half*>(t_emb), static_cast<const __half*>(t_emb2), static_cast<const __half*>(cam_emb), B,
        F, E, static_cast<__half*>(out));
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_embed_combine_silu");
    return VMV_OK;
}

extern "C" int vmv_cfg_ddim_step(const float* xt, const float* y_out, const float* u_out, const float* coef7, int64_t n,
                                 float* x_prev, void* stream) {
    VMV_CHECK_ARG(xt && y_out && u_out && coef7 && x_prev && n > 0, "vmv_cfg_ddim_step: bad args");
    launch_kernel(cfg_ddim_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
        xt, y_out, u_out, coef7, n, x_prev);
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_cfg_ddim_step");
    return VMV_OK;
}
