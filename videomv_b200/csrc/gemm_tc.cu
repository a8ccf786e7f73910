// Tensor-core contraction kernel for the VideoMV UNet: Linear / 1x1 / 3x3 conv / (3,1,1) temporal conv.
//
//   D[M,N] = epilogue( A (*) W^T ),   A fp16 (K-major), W fp16 [N,Ktot] (K-major), fp32 accumulate in TMEM.
//
// One CTA = one 128 x BN output tile (or one K split of it).  Warp roles (192 threads):
//   warp 0      TMA producer: per K block (64 halfs = one 128B swizzle atom) one A box + one W box
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (4 x K=16 MMAs per K block)
//   warps 2..5  epilogue: tcgen05.ld (32 lanes x 16 cols) -> fp32 epilogue -> fp16 16B global stores
// smem ring of STAGES {A 128x64, W BNx64} tiles, full/empty mbarriers; tcgen05.commit frees a stage.
//
// The A operand is never materialised as im2col: for the 3x3 conv each K block is one 4-D TMA box
// (64 ch, bw, bh, bf) of the channels-last activation shifted by the tap offset, with the halo supplied
// by TMA out-of-bounds zero fill; the temporal conv shifts the box along the frame axis of a
// (C, HW, F, B) view.  Box rows land in smem in exactly the (f,h,w) row order of the output tile, so the
// 128B-swizzled tile is a legal K-major UMMA operand as is.
#include "common.cuh"
#include "../../include/videomv_b200.h"

#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace vmv {

void count_launch(int n = 1);

#ifndef VMV_EPI_SPLIT
#define VMV_EPI_SPLIT 2
#endif
#ifndef VMV_EPI_COLS
#define VMV_EPI_COLS 32
#endif
// VMV_GEMM_DEBUG knock-out experiments (skip stores / epilogue / operand traffic / MMAs) are compiled in only with
// -DVMV_GEMM_DEBUG_BUILD: their run-time tests sat in the hot loops of the production kernel
#ifdef VMV_GEMM_DEBUG_BUILD
#define VMV_DBG(a_) ((a_).dbg)
#else
#define VMV_DBG(a_) 0
#endif
constexpr int EPI_COLS = VMV_EPI_COLS;           // columns an epilogue thread drains per step: 32, or 16 (half the live registers, for more epilogue warps)
constexpr int EPI_SPLIT = VMV_EPI_SPLIT;         // epilogue warps per TMEM lane quarter: they share a quarter's rows and interleave the tile's 32-column blocks
constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KiB

struct GemmArgs {
    int mode, M, N, n_out;
    int nkb;          // total K blocks
    int nkb1;         // linear: K blocks in source 1
    int cblocks;      // conv: K blocks per tap (= Cin/64)
    int kb_per_split; // K blocks per split (== nkb when no split)
    // conv geometry (tile decode)
    int F, H, W, HW;
    int bw, bh, bf, bp;
    int tiles_w, tiles_h;          // CONV3X3 family
    int tiles_per_sample, tiles_p; // TCONV3
    // CONV3X3_S2: the tile geometry above is that of the OUTPUT (H/2 x W/2); TMA coordinates are 2 * (output coordinate) + tap.
    // UPCONV3X3:  the tile geometry is that of the INPUT grid; m tiles come in 4 phases (py, px) of `mt_phase` tiles each,
    //             output pixel (2y+py, 2x+px); 4 taps (ty, tx) at input offset (ty - 1 + py, tx - 1 + px); W rows phase*N + n.
    int mt_phase;                  // m tiles per phase (== all m tiles when the mode has no phases)
    // A-resident schedule (small K: all K blocks of an m tile fit the ring).  The ring depth is set to the number of K
    // blocks, so stage s always holds K block s of the CURRENT m tile; every cluster works through a contiguous run of
    // tiles (all N tiles of an m pair back to back) and re-loads only W for the 2nd, 3rd, ... N tile of an m pair: the A
    // rows cross the L2 -> SM path once instead of once per N tile (the K <= 512 transformer GEMMs at the 32x32 level are
    // bound by exactly that path).
    int a_res;
    // epilogue
    __half* D;
    long long ldd;
    const float* bias;
    const __half* rowbias;
    long long ld_rowbias;
    int rows_per_group;
    const __half* residual;
    long long ldr;
    int act;
    float* partial;   // split-K: fp32 [splits, M, N]
    const float2* ln_stats;   // folded LayerNorm: per-row {mean, rstd}, or [M][ln_nslots] partial {mean_k, M2_k} slots
    const float* ln_colsum;   // folded LayerNorm: per-column sum of the gamma-scaled weights
    int dbg;          // profiling experiments (VMV_GEMM_DEBUG): 1 = skip the TMA stores, 2 = skip the whole epilogue body
    int fast_epi;     // v2: register epilogue with 256-bit global accesses (needs 32 B aligned D / residual / rowbias rows)
    int w_static;     // W may be fetched ahead of the programmatic-dependent-launch wait (model weights)
    // ln_nslots != 0: ln_stats holds, per row, ln_nslots partial statistics {mean_k, M2_k} written by the epilogue of the
    // upstream GEMM (its `rowstats`): slot (nt, hh) covers the 32-column blocks hh, hh+2, ... of that GEMM's N tile nt
    // (ln_src_n columns in tiles of ln_src_bn).  They are merged in slot order (Chan et al.) -- no atomics, bit-reproducible,
    // and free of the E[x^2]-E[x]^2 cancellation for rows with a large mean.
    int ln_nslots, ln_src_n, ln_src_bn;
    float ln_eps;
    float2* rowstats; // fp32 [M][2*n_tiles][2]: partial {mean, M2} of every output row per (N tile, epilogue warp half)
    int rowstats_nslots;
    // multi-GPU: output rows go to the peers' tensors of the other sharding layout + epoch-flag rendezvous at the end (csrc/peer.cu)
    int sc_world, sc_rank, sc_dir, sc_nowait;
    int sc_B, sc_Fl, sc_HWl;
    __half* sc_dst[VMV_PEER_MAX_RANKS];
    unsigned int* sc_flags[VMV_PEER_MAX_RANKS];
    unsigned int* sc_epoch;
    unsigned int* sc_done;
};

// destination of output row `grow` when the epilogue scatters into the other sharding layout (32-bit index arithmetic:
// M < 2^31)
__device__ __forceinline__ __half* scatter_row(const GemmArgs& a, long long grow_) {
    const unsigned P = (unsigned)a.sc_world, HWl = (unsigned)a.sc_HWl, Fl = (unsigned)a.sc_Fl, grow = (unsigned)grow_;
    unsigned q;
    unsigned long long row;
    if (a.sc_dir == 0) {                               // rows (b, f, pixel), pixel < P*HWl
        const unsigned HW = HWl * P;
        const unsigned bf = grow / HW, pix = grow - bf * HW;
        const unsigned b = bf / Fl, f = bf - b * Fl;
        q = pix / HWl;
        row = ((unsigned long long)(b * P + (unsigned)a.sc_rank) * Fl + f) * HWl + (pix - q * HWl);
    } else {                                           // rows (b, fg, pl), fg < P*Fl
        const unsigned F = Fl * P;
        const unsigned bf = grow / HWl, pl = grow - bf * HWl;
        const unsigned b = bf / F, fg = bf - b * F;
        q = fg / Fl;
        row = ((unsigned long long)(b * Fl + (fg - q * Fl)) * P + (unsigned)a.sc_rank) * HWl + pl;
    }
    return a.sc_dst[q] + row * a.ldd;
}

constexpr int LN_MAX_SLOTS = 16;      // EPI_SPLIT x ceil(1280 / 256) for EPI_SPLIT <= 3

// number of columns slot s of a row covers, from the producing GEMM's tiling
__device__ __forceinline__ int ln_slot_count(const GemmArgs& a, int s) {
    const int nt = s / EPI_SPLIT, hh = s - nt * EPI_SPLIT;
    int nvalid = min(a.ln_src_bn / 32, (a.ln_src_n - nt * a.ln_src_bn + 31) / 32);
    nvalid = max(nvalid, 0);
    return 32 * max((nvalid - hh + EPI_SPLIT - 1) / EPI_SPLIT, 0);
}

// Folded-LayerNorm row statistics {mean, rstd}: stored as such, or merged from the slots an upstream GEMM wrote -- in slot
// order with the parallel-variance formula, four independent loads in flight at a time.
__device__ __forceinline__ float2 ln_row_stats(const GemmArgs& a, long long row) {
    if (a.ln_nslots == 0) return a.ln_stats[row];
    const float2* sp = a.ln_stats + row * a.ln_nslots;
    float n = 0.f, mean = 0.f, m2 = 0.f;
#pragma unroll 1
    for (int s0 = 0; s0 < a.ln_nslots; s0 += 4) {
        float2 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (s0 + j < a.ln_nslots) ? sp[s0 + j] : make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float nk = (s0 + j < a.ln_nslots) ? (float)ln_slot_count(a, s0 + j) : 0.f;
            if (nk > 0.f) {
                const float nn = n + nk;
                const float d = v[j].x - mean;
                const float w = __fdividef(nk, nn);
                mean = fmaf(d, w, mean);
                m2 += fmaf(d * d, n * w, v[j].y);
                n = nn;
            }
        }
    }
    return make_float2(mean, rsqrtf(__fdividef(m2, n) + a.ln_eps));
}

// Decode (m tile, row in tile) -> global output row; returns -1 when the row is padding.
__device__ __forceinline__ long long tile_row_to_global(const GemmArgs& a, int mt, int r, int phase = 0) {
    if (a.mode == VMV_GEMM_LINEAR) {
        long long g = (long long)mt * BM + r;
        return g < a.M ? g : -1;
    } else if (a.mode == VMV_GEMM_CONV3X3 || a.mode == VMV_GEMM_CONV3X3_S2) {
        int tw = mt % a.tiles_w;
        int th = (mt / a.tiles_w) % a.tiles_h;
        int tf = mt / (a.tiles_w * a.tiles_h);
        int rw = r % a.bw;
        int rh = (r / a.bw) % a.bh;
        int rf = r / (a.bw * a.bh);
        long long f = (long long)tf * a.bf + rf;
        long long g = (f * a.H + (th * a.bh + rh)) * a.W + (tw * a.bw + rw);
        return g < a.M ? g : -1;
    } else if (a.mode == VMV_GEMM_UPCONV3X3) {
        if (mt >= a.mt_phase) return -1;                   // phantom second tile of an odd pair
        int tw = mt % a.tiles_w;
        int th = (mt / a.tiles_w) % a.tiles_h;
        int tf = mt / (a.tiles_w * a.tiles_h);
        int rw = r % a.bw;
        int rh = (r / a.bw) % a.bh;
        int rf = r / (a.bw * a.bh);
        long long f = (long long)tf * a.bf + rf;
        const int y = th * a.bh + rh, x = tw * a.bw + rw;
        long long g = (f * (2 * a.H) + (2 * y + (phase >> 1))) * (2 * a.W) + (2 * x + (phase & 1));
        return (f < a.F && g < a.M) ? g : -1;
    } else {
        int b = mt / a.tiles_per_sample;
        int ts = mt % a.tiles_per_sample;
        int tp = ts % a.tiles_p;
        int tf = ts / a.tiles_p;
        int rp = r % a.bp;
        int rf = r / a.bp;
        int f = tf * a.bf + rf;
        if (f >= a.F) return -1;
        long long g = ((long long)b * a.F + f) * a.HW + (tp * a.bp + rp);
        return g < a.M ? g : -1;          // b == B for the phantom second tile of an odd CTA pair
    }
}

// First TMA row coordinate of tile rows [r, r+32) (r multiple of 32): rows of a tile are contiguous in memory for every
// tiling make_plan accepts.  For the temporal conv the coordinate is sample-local and *b is the sample index.
__device__ __forceinline__ long long tile_row0(const GemmArgs& a, int mt, int r, int* b) {
    *b = 0;
    if (a.mode == VMV_GEMM_LINEAR) return (long long)mt * BM + r;
    if (a.mode == VMV_GEMM_CONV3X3) {
        int tw = mt % a.tiles_w;
        int th = (mt / a.tiles_w) % a.tiles_h;
        int tf = mt / (a.tiles_w * a.tiles_h);
        int rw = r % a.bw;
        int rh = (r / a.bw) % a.bh;
        int rf = r / (a.bw * a.bh);
        return (((long long)tf * a.bf + rf) * a.H + (th * a.bh + rh)) * a.W + (tw * a.bw + rw);
    }
    *b = mt / a.tiles_per_sample;
    int ts = mt % a.tiles_per_sample;
    int tp = ts % a.tiles_p;
    int tf = ts / a.tiles_p;
    return ((long long)tf * a.bf + r / a.bp) * a.HW + (tp * a.bp + r % a.bp);
}

// Drain one 128 x BN fp32 accumulator tile from TMEM (this thread: one row, `trow` = TMEM address of its lane,
// column 0 of the tile) through the fused epilogue into global memory.  Warp-uniform control flow around tcgen05.ld.
template <int BN>
__device__ __forceinline__ void epilogue_store(const GemmArgs& a, int nt, int split, long long grow, bool valid,
                                               uint32_t trow) {
    const int n0 = nt * BN;
    if (a.partial != nullptr) {
        // split-K: raw fp32 partial sums, reduced by splitk_finish_kernel
        float* prow = a.partial + ((long long)split * a.M + (valid ? grow : 0)) * a.N;
#pragma unroll 1
        for (int c = 0; c < BN; c += 16) {
            if (n0 + c >= a.N) break;
            uint32_t v[16];
            tmem_ld_32x32b_x16(trow + c, v);
            tmem_ld_wait();
            if (valid) {
                float4* dst = reinterpret_cast<float4*>(prow + n0 + c);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                         __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
            }
        }
    } else if (a.act == VMV_ACT_GEGLU) {
        constexpr int HB = BN / 2;
        const int o0 = nt * HB;
#pragma unroll 1
        for (int c = 0; c < HB; c += 16) {
            if (n0 + c >= a.N) break;
            uint32_t v[16], g[16];
            tmem_ld_32x32b_x16(trow + c, v);
            tmem_ld_32x32b_x16(trow + HB + c, g);
            tmem_ld_wait();
            if (valid) {
                float x[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float val = __uint_as_float(v[j]), gate = __uint_as_float(g[j]);
                    if (a.ln_stats) {
                        const float2 ms = ln_row_stats(a, grow);
                        val = ms.y * (val - ms.x * __ldg(a.ln_colsum + n0 + c + j));
                        gate = ms.y * (gate - ms.x * __ldg(a.ln_colsum + n0 + HB + c + j));
                    }
                    if (a.bias) {
                        val += __ldg(a.bias + n0 + c + j);
                        gate += __ldg(a.bias + n0 + HB + c + j);
                    }
                    x[j] = val * gelu_erf_f(gate);
                }
                if (a.residual) {
                    const uint4* rp = reinterpret_cast<const uint4*>(a.residual + grow * a.ldr + o0 + c);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint4 u = __ldg(rp + h);
                        uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float2 f = unpack_half2(w[j]);
                            x[8 * h + 2 * j] += f.x;
                            x[8 * h + 2 * j + 1] += f.y;
                        }
                    }
                }
                uint4* dp = reinterpret_cast<uint4*>(a.D + grow * a.ldd + o0 + c);
                dp[0] = make_uint4(pack_half2(x[0], x[1]), pack_half2(x[2], x[3]), pack_half2(x[4], x[5]),
                                   pack_half2(x[6], x[7]));
                dp[1] = make_uint4(pack_half2(x[8], x[9]), pack_half2(x[10], x[11]), pack_half2(x[12], x[13]),
                                   pack_half2(x[14], x[15]));
            }
        }
    } else {
        const __half* rb = nullptr;
        if (a.rowbias && valid) rb = a.rowbias + (grow / a.rows_per_group) * a.ld_rowbias;
#pragma unroll 1
        for (int c = 0; c < BN; c += 16) {
            if (n0 + c >= a.N) break;
            uint32_t v[16];
            tmem_ld_32x32b_x16(trow + c, v);
            tmem_ld_wait();
            if (valid) {
                const int n = n0 + c;
                float x[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(v[j]);
                if (a.ln_stats) {
                    const float2 ms = ln_row_stats(a, grow);
#pragma unroll
                    for (int j = 0; j < 16; ++j) x[j] = ms.y * (x[j] - ms.x * __ldg(a.ln_colsum + n + j));
                }
                if (a.bias) {
                    const float4* bp = reinterpret_cast<const float4*>(a.bias + n);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 b4 = __ldg(bp + j);
                        x[4 * j] += b4.x; x[4 * j + 1] += b4.y; x[4 * j + 2] += b4.z; x[4 * j + 3] += b4.w;
                    }
                }
                if (rb) {
                    const uint4* rp = reinterpret_cast<const uint4*>(rb + n);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint4 u = __ldg(rp + h);
                        uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float2 f = unpack_half2(w[j]);
                            x[8 * h + 2 * j] += f.x;
                            x[8 * h + 2 * j + 1] += f.y;
                        }
                    }
                }
                if (a.act == VMV_ACT_SILU) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) x[j] = silu_f(x[j]);
                }
                if (a.residual) {
                    const uint4* rp = reinterpret_cast<const uint4*>(a.residual + grow * a.ldr + n);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint4 u = __ldg(rp + h);
                        uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float2 f = unpack_half2(w[j]);
                            x[8 * h + 2 * j] += f.x;
                            x[8 * h + 2 * j + 1] += f.y;
                        }
                    }
                }
                uint4* dp = reinterpret_cast<uint4*>(a.D + grow * a.ldd + n);
                dp[0] = make_uint4(pack_half2(x[0], x[1]), pack_half2(x[2], x[3]), pack_half2(x[4], x[5]),
                                   pack_half2(x[6], x[7]));
                dp[1] = make_uint4(pack_half2(x[8], x[9]), pack_half2(x[10], x[11]), pack_half2(x[12], x[13]),
                                   pack_half2(x[14], x[15]));
            }
        }
    }
}

template <int BN, int STAGES>
struct SmemLayout {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int A_OFF = 0;
    static constexpr int B_OFF = STAGES * A_STAGE_BYTES;
    static constexpr int BAR_OFF = B_OFF + STAGES * B_STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16;
    static constexpr int DYN_BYTES = TOTAL + 1024;   // slack for manual 1024B alignment
};

template <int BN>
struct TmemCols {
    static constexpr int value = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : BN <= 256 ? 256 : 512;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmW, const GemmArgs a) {
    using L = SmemLayout<BN, STAGES>;
    constexpr int TCOLS = TmemCols<BN>::value;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nt = blockIdx.x;      // N tile
    const int mt = blockIdx.y;      // M tile
    const int split = blockIdx.z;
    const int kb_begin = split * a.kb_per_split;
    const int kb_end = min(a.nkb, kb_begin + a.kb_per_split);
    const int n0 = nt * BN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA1);
        tma_prefetch_desc(&tmW);
        if (a.nkb1 < a.nkb && a.mode == VMV_GEMM_LINEAR) tma_prefetch_desc(&tmA2);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TCOLS>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------ TMA producer ------------------------------
            // tile origin in the coordinates of the A tensor map
            int c1 = 0, c2 = 0, c3 = 0;
            if (a.mode == VMV_GEMM_LINEAR) {
                c1 = mt * BM;
            } else if (a.mode == VMV_GEMM_CONV3X3) {
                c1 = (mt % a.tiles_w) * a.bw;
                c2 = ((mt / a.tiles_w) % a.tiles_h) * a.bh;
                c3 = (mt / (a.tiles_w * a.tiles_h)) * a.bf;
            } else {
                int ts = mt % a.tiles_per_sample;
                c1 = (ts % a.tiles_p) * a.bp;
                c2 = (ts / a.tiles_p) * a.bf;
                c3 = mt / a.tiles_per_sample;
            }
            constexpr uint32_t tx_bytes = A_STAGE_BYTES + L::B_STAGE_BYTES;
            int it = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
                void* sa = smem + L::A_OFF + s * A_STAGE_BYTES;
                void* sb = smem + L::B_OFF + s * L::B_STAGE_BYTES;
                if (a.mode == VMV_GEMM_LINEAR) {
                    if (kb < a.nkb1) tma_load_2d(sa, &tmA1, &full_bar[s], kb * BK, c1);
                    else tma_load_2d(sa, &tmA2, &full_bar[s], (kb - a.nkb1) * BK, c1);
                } else if (a.mode == VMV_GEMM_CONV3X3) {
                    const int tap = kb / a.cblocks, cb = kb - tap * a.cblocks;
                    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                    tma_load_4d(sa, &tmA1, &full_bar[s], cb * BK, c1 + dx, c2 + dy, c3);
                } else {
                    const int tap = kb / a.cblocks, cb = kb - tap * a.cblocks;
                    tma_load_4d(sa, &tmA1, &full_bar[s], cb * BK, c1, c2 + tap - 1, c3);
                }
                tma_load_2d(sb, &tmW, &full_bar[s], kb * BK, n0);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------ MMA issuer --------------------------------
            constexpr uint32_t idesc = umma_idesc_f16_f32(BM, BN);
            int it = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(smem + L::A_OFF + s * A_STAGE_BYTES));
                const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(smem + L::B_OFF + s * L::B_STAGE_BYTES));
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    // advance 16 halfs = 32 bytes inside the 128B swizzle atom: +2 in the (addr>>4) field
                    umma_f16_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);     // frees the smem stage when these MMAs retire
            }
            umma_commit(tmem_full_bar);         // accumulator complete
        }
        __syncwarp();
    } else {
        // ---------------------------------- epilogue ----------------------------------
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;
        const long long grow = tile_row_to_global(a, mt, r);
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        epilogue_store<BN>(a, nt, split, grow, grow >= 0, tmem_base + (static_cast<uint32_t>(q * 32) << 16));
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TCOLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// v2: persistent CTA-pair kernel.  A cluster of 2 CTAs (one SM each) owns 256 x BN output tiles:
//   * tcgen05.mma.cta_group::2, M = 256: each CTA stages its own 128 A rows and HALF of the W rows, so per K block the
//     pair moves (256 + BN) x 128 B through L2 for 2*256*BN*64 FLOP -- 1.4-1.6x the arithmetic intensity of the 1-CTA
//     kernel, which is L2-bandwidth bound on this part;
//   * persistent: grid = one cluster per SM pair, static round-robin tile scheduler, so barrier init / TMEM allocation
//     / descriptor prefetch are paid once per launch instead of once per tile;
//   * two TMEM accumulator buffers (2 x BN columns): the 8 epilogue warps of the pair drain tile i while the MMA
//     thread already accumulates tile i+1 (tmem_full / tmem_empty mbarriers).
// Barrier topology (DeepGEMM-style): full[s] lives in the leader CTA (count 2: leader arrive.expect_tx for both CTAs'
// bytes + the peer's remote arrive; both CTAs' TMA complete_tx are routed to it), empty[s] and tmem_full[b] exist in
// both CTAs and are signalled by multicast tcgen05.commit, tmem_empty[b] lives in the leader (count 8 = epilogue
// warps of both CTAs).
// ------------------------------------------------------------------------------------------------
constexpr int EPI_BLK_COLS = 32;                 // epilogue column block: 32 fp16 = 64 B rows, SWIZZLE_64B boxes
constexpr int EPI_BLK_BYTES = 32 * 64;            // 32 rows x 64 B per warp per block

constexpr int V2_EPI_WARPS = 4 * EPI_SPLIT;
constexpr int V2_THREADS = 64 + 32 * V2_EPI_WARPS;   // warp 0 TMA, warp 1 MMA, then the epilogue warps

template <int BN, int STAGES, int NBLK_>
struct SmemLayout2 {
    static constexpr int B_STAGE_BYTES = (BN / 2) * BK * 2;
    static constexpr int NBLK = NBLK_;
    static constexpr int A_OFF = 0;
    static constexpr int B_OFF = STAGES * A_STAGE_BYTES;
    static constexpr int BAR_OFF = B_OFF + STAGES * B_STAGE_BYTES;
    static constexpr int NBARS = 2 * STAGES + 4;
    static constexpr int TOTAL = BAR_OFF + NBARS * 8 + 16;
    static constexpr int DYN_BYTES = TOTAL + 1024;
    // + the kernel's static smem: the per-epilogue-warp bias / column-sum scratch (V2_EPI_WARPS x 1 KiB)
    static_assert(DYN_BYTES + V2_EPI_WARPS * 1024 + 1024 <= 232448, "shared memory budget exceeded");
};

// FLAVOR specialises the epilogue at compile time (the one kernel with every path behind run-time branches was 9.9 k SASS
// instructions = 159 KB of code; ncu showed instruction-fetch stalls in the epilogue warps):
//   1 = register epilogue with GEGLU, 2 = generic epilogue (split-K partials, unaligned rows),
//   0 / 4 / 8 / 12 = register epilogue without GEGLU; bit 2: a LayerNorm is folded in, bit 3: a residual is added
template <int BN, int STAGES, int FLAVOR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(V2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                const __grid_constant__ CUtensorMap tmW, const GemmArgs a, const int m_pairs, const int n_tiles,
                const int splits) {
    using L = SmemLayout2<BN, STAGES, FLAVOR>;
    constexpr int TCOLS = TmemCols<2 * BN>::value;
    static_assert(2 * BN <= 512, "two accumulator buffers must fit TMEM");
    // Per-epilogue-warp scratch: 128 bias floats + 128 LayerNorm column sums of the columns the warp owns in the current
    // tile.  A static __shared__ array (not carved out of the manually aligned dynamic buffer) so that the compiler knows
    // the address space and alignment: plain C++ float4 reads become LDS.128 that it is free to schedule among the math.
    __shared__ __align__(16) float s_epi_scr[V2_EPI_WARPS][256];
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2], used in the leader only
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int tiles_mn = m_pairs * n_tiles;
    const int total_tiles = tiles_mn * splits;
    const int mp_phase = (a.mt_phase + 1) / 2;          // CTA pairs per output phase (all of them when the mode has one phase)
    // tile schedule: round-robin over clusters, or (A-resident) one contiguous run of tiles per cluster
    const int depth = a.a_res ? a.nkb : STAGES;         // ring depth in use
    const int run = a.a_res ? (total_tiles + num_clusters - 1) / num_clusters : 0;
    const int t_begin = a.a_res ? min(total_tiles, cluster_id * run) : cluster_id;
    const int t_end = a.a_res ? min(total_tiles, (cluster_id + 1) * run) : total_tiles;
    const int t_step = a.a_res ? 1 : num_clusters;

    cluster_sync_all();                                 // both CTAs resident before the pair-wide TMEM allocation
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA1);
        tma_prefetch_desc(&tmW);
        if (a.nkb1 < a.nkb && a.mode == VMV_GEMM_LINEAR) tma_prefetch_desc(&tmA2);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 2);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full_bar[b], 1);
            mbar_init(&tmem_empty_bar[b], 2 * V2_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm<TCOLS>(tmem_ptr_smem);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // Programmatic dependent launch: the prologue above ran while the previous kernel was still draining.  Everything
    // that touches activations waits here; only the producer thread goes on first, to put the W tiles of its first
    // pipeline stages in flight (weights are not produced by the previous kernel) before it waits as well.
    pdl_launch_dependents();
    const bool is_producer = warp == 0 && lane == 0;
    if (!is_producer) pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------ TMA producer (both CTAs) ------------------------------
            constexpr uint32_t tx_bytes = 2u * (A_STAGE_BYTES + L::B_STAGE_BYTES);
            int pre = 0;                                    // stages whose barrier arrival + W load were issued pre-wait
            if (a.w_static && !(VMV_DBG(a) & 4) && t_begin < t_end) {
                const int split = t_begin / tiles_mn;
                const int rem0 = t_begin - split * tiles_mn;
                const int nt = rem0 % n_tiles;
                const int nt_cols = min(BN, a.N - nt * BN);
                const int nrow = ((rem0 / n_tiles) / mp_phase) * a.N + nt * BN + (int)rank * (nt_cols / 2);
                const int kb_begin = split * a.kb_per_split;
                const int kb_end = min(a.nkb, kb_begin + a.kb_per_split);
                pre = min(depth, kb_end - kb_begin);
                for (int i = 0; i < pre; ++i) {
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[i], tx_bytes);
                    else mbar_arrive_remote(&full_bar[i], 0);
                    tma_load_2d_2sm(smem + L::B_OFF + i * L::B_STAGE_BYTES, &tmW, &full_bar[i], (kb_begin + i) * BK, nrow);
                }
            }
            pdl_wait();
            int it = 0;
            for (int t = t_begin; t < t_end; t += t_step) {
                const int split = t / tiles_mn;
                const int rem = t - split * tiles_mn;
                const int mp = rem / n_tiles, nt = rem - mp * n_tiles;
                // A-resident: the A rows of this m pair are already in the ring unless this is the first tile of my run or of the m pair
                const bool load_a = !a.a_res || t == t_begin || nt == 0;
                const int phase = mp / mp_phase;                // UPCONV3X3 only (0 otherwise)
                const int mt = 2 * (mp - phase * mp_phase) + (int)rank;
                // the pair splits the tile's W rows; a ragged last tile (N % BN != 0) is nt_cols wide and each CTA
                // contributes nt_cols/2 rows (its box still loads BN/2 rows; the surplus is never read by the MMA)
                const int nt_cols = min(BN, a.N - nt * BN);
                const int nrow = phase * a.N + nt * BN + (int)rank * (nt_cols / 2);
                int c1 = 0, c2 = 0, c3 = 0;
                if (a.mode == VMV_GEMM_LINEAR) {
                    c1 = mt * BM;
                } else if (a.mode == VMV_GEMM_CONV3X3 || a.mode == VMV_GEMM_CONV3X3_S2 || a.mode == VMV_GEMM_UPCONV3X3) {
                    c1 = (mt % a.tiles_w) * a.bw;
                    c2 = ((mt / a.tiles_w) % a.tiles_h) * a.bh;
                    c3 = (mt / (a.tiles_w * a.tiles_h)) * a.bf;
                    if (a.mode == VMV_GEMM_CONV3X3_S2) { c1 *= 2; c2 *= 2; }     // input coordinates of the stride-2 window
                } else {
                    int ts = mt % a.tiles_per_sample;
                    c1 = (ts % a.tiles_p) * a.bp;
                    c2 = (ts / a.tiles_p) * a.bf;
                    c3 = mt / a.tiles_per_sample;
                }
                const int kb_begin = split * a.kb_per_split;
                const int kb_end = min(a.nkb, kb_begin + a.kb_per_split);
                for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
                    const int s = it % depth;
                    const uint32_t ph = (it / depth) & 1;
                    const bool prefetched = it < pre;      // first ring pass: arrival + W tile already issued
                    if (!prefetched) mbar_wait(&empty_bar[s], ph ^ 1);
                    if (VMV_DBG(a) & 4) {                       // experiment: no operand traffic, barrier protocol only
                        if (rank == 0) mbar_arrive(&full_bar[s]);
                        else mbar_arrive_remote(&full_bar[s], 0);
                        continue;
                    }
                    if (!prefetched) {
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], load_a ? tx_bytes : 2u * L::B_STAGE_BYTES);
                        else mbar_arrive_remote(&full_bar[s], 0);
                    }
                    void* sa = smem + L::A_OFF + s * A_STAGE_BYTES;
                    void* sb = smem + L::B_OFF + s * L::B_STAGE_BYTES;
                    if (!load_a) {
                        // stage s still holds K block kb of this m pair's A rows
                    } else if (a.mode == VMV_GEMM_LINEAR) {
                        if (kb < a.nkb1) tma_load_2d_2sm(sa, &tmA1, &full_bar[s], kb * BK, c1);
                        else tma_load_2d_2sm(sa, &tmA2, &full_bar[s], (kb - a.nkb1) * BK, c1);
                    } else if (a.mode == VMV_GEMM_CONV3X3 || a.mode == VMV_GEMM_CONV3X3_S2) {
                        const int tap = kb / a.cblocks, cb = kb - tap * a.cblocks;
                        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                        tma_load_4d_2sm(sa, &tmA1, &full_bar[s], cb * BK, c1 + dx, c2 + dy, c3);
                    } else if (a.mode == VMV_GEMM_UPCONV3X3) {
                        const int tap = kb / a.cblocks, cb = kb - tap * a.cblocks;      // 4 taps (ty, tx)
                        const int dy = (tap >> 1) - 1 + (phase >> 1), dx = (tap & 1) - 1 + (phase & 1);
                        tma_load_4d_2sm(sa, &tmA1, &full_bar[s], cb * BK, c1 + dx, c2 + dy, c3);
                    } else {
                        const int tap = kb / a.cblocks, cb = kb - tap * a.cblocks;
                        tma_load_4d_2sm(sa, &tmA1, &full_bar[s], cb * BK, c1, c2 + tap - 1, c3);
                    }
                    if (!prefetched) tma_load_2d_2sm(sb, &tmW, &full_bar[s], kb * BK, nrow);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // ------------------------------ MMA issuer (leader CTA only) ------------------------------
            int it = 0, acc_it = 0;
            for (int t = t_begin; t < t_end; t += t_step, ++acc_it) {
                const int split = t / tiles_mn;
                const int nt_mma = (t - split * tiles_mn) % n_tiles;
                const uint32_t idesc = umma_idesc_f16_f32(2 * BM, min(BN, a.N - nt_mma * BN));   // ragged last N tile
                const int kb_begin = split * a.kb_per_split;
                const int kb_end = min(a.nkb, kb_begin + a.kb_per_split);
                const int buf = acc_it & 1;
                const uint32_t aph = (acc_it >> 1) & 1;
                mbar_wait(&tmem_empty_bar[buf], aph ^ 1);      // epilogues of both CTAs drained this buffer
                tc_fence_after();
                const uint32_t dcol = tmem_base + buf * BN;
                for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
                    const int s = it % depth;
                    const uint32_t ph = (it / depth) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(smem + L::A_OFF + s * A_STAGE_BYTES));
                    const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(smem + L::B_OFF + s * L::B_STAGE_BYTES));
                    if (!(VMV_DBG(a) & 8)) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_f16_ss_2sm(dcol, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
                    }
                    umma_commit_2sm(&empty_bar[s]);
                }
                umma_commit_2sm(&tmem_full_bar[buf]);
            }
        }
        __syncwarp();
    } else {
        // ---------------------------------- epilogue (both CTAs, 8 warps) ----------------------------------
        // Warp w may only touch TMEM lanes [32*(w%4), +32): two warps share each lane quarter and split the tile's
        // 32-column output blocks (even / odd) -- twice the issue slots and latency hiding of a 4-warp epilogue.
        // Each lane owns one output row: per block it drains 32 fp32 accumulators (2 x tcgen05.ld.x16), applies the
        // fused epilogue and writes 64 contiguous bytes with two 256-bit stores (full 32 B sectors, no staging).
        // Global-latency loads are taken off the accumulator critical path: bias / LayerNorm column sums are staged in
        // smem before the tile's MMAs finish, and the residual of block i+1 is requested before block i is processed.
        const int q = warp & 3;
        const int hh = (warp - 2) >> 2;                         // which of the quarter's EPI_SPLIT warps: blocks hh, hh + EPI_SPLIT, ...
        const int r = q * 32 + lane;
        float* sbias = s_epi_scr[warp - 2];
        float* scol = sbias + 128;
        constexpr bool geglu = FLAVOR == 1;
        const int out_bn = geglu ? BN / 2 : BN;                 // output columns per tile
        // Bias / LayerNorm column sums of this warp's columns are fetched ONE TILE AHEAD into registers (<= 128 values per
        // array per warp = 4 per lane) and only copied to the smem scratch at the top of their tile: when the epilogue is
        // the bottleneck (K <= 640) the accumulator is already waiting, and a global-load -> smem-store chain at that point
        // was the largest stall of the kernel (ncu source view, profiles/r1_s2_ncu_hot_kernels_summary.md).
        // (always staged -- zeros when a pointer is null -- so the math below never reads an unwritten slot)
        const int per = geglu ? 64 : 32;
        float nb[4] = {0.f, 0.f, 0.f, 0.f}, nc[4] = {0.f, 0.f, 0.f, 0.f};
        auto fetch_cols = [&](int tt) {
            if (tt >= t_end) return;
            const int rem_ = tt % tiles_mn;
            const int nt_ = rem_ % n_tiles;
            const int col0_ = nt_ * out_bn;
            int nvalid_ = min(out_bn / EPI_BLK_COLS, (a.n_out - col0_ + EPI_BLK_COLS - 1) / EPI_BLK_COLS);
            if (nvalid_ < 0) nvalid_ = 0;
            const int cnt = max((nvalid_ - hh + EPI_SPLIT - 1) / EPI_SPLIT, 0) * per;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
                const int i = lane + 32 * qq;
                if (i < cnt) {
                    const int j = i / per, w = i - j * per;
                    const int blk = hh + EPI_SPLIT * j;
                    const int n = geglu ? nt_ * BN + blk * 32 + (w & 31) + (w >= 32 ? BN / 2 : 0) : col0_ + blk * 32 + w;
                    nb[qq] = a.bias ? __ldg(a.bias + n) : 0.f;
                    nc[qq] = a.ln_colsum ? __ldg(a.ln_colsum + n) : 0.f;
                }
            }
        };
        if (FLAVOR != 2) fetch_cols(t_begin);
        int acc_it = 0;
        for (int t = t_begin; t < t_end; t += t_step, ++acc_it) {
            const int split = t / tiles_mn;
            const int rem = t - split * tiles_mn;
            const int mp = rem / n_tiles, nt = rem - mp * n_tiles;
            const int phase = mp / mp_phase;
            const int mt = 2 * (mp - phase * mp_phase) + (int)rank;
            const int buf = acc_it & 1;
            const uint32_t aph = (acc_it >> 1) & 1;
            const long long grow = tile_row_to_global(a, mt, r, phase);
            const bool valid = grow >= 0;
            const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN;
            const int col0 = nt * out_bn;
            int nvalid = min(out_bn / EPI_BLK_COLS, (a.n_out - col0 + EPI_BLK_COLS - 1) / EPI_BLK_COLS);
            if (nvalid < 0) nvalid = 0;
            // (compile-time for the non-GEGLU register flavors, run-time otherwise)
            constexpr bool kFixed = FLAVOR != 1 && FLAVOR != 2;
            const bool use_res = kFixed ? (FLAVOR & 8) != 0 : a.residual != nullptr;
            const bool use_ln = kFixed ? (FLAVOR & 4) != 0 : a.ln_stats != nullptr;
            const __half* resrow = (use_res && valid) ? a.residual + grow * a.ldr + col0 : nullptr;
            __half* drow = nullptr;                              // my output row (possibly in a peer's memory)
            if (valid) drow = a.sc_world ? scatter_row(a, grow) : a.D + grow * a.ldd;
            uint32_t rcur[16];
            float2 ln_ms = make_float2(0.f, 1.f);
            if (FLAVOR != 2) {
                // (1) this warp's bias / column-sum slices (fetched one tile ago) go to the smem scratch.  Slot j*32+i holds
                //     column (hh+2j)*32+i of the tile; GEGLU keeps value and gate columns in slots j*64+i and j*64+32+i.
                __syncwarp();                                   // previous tile's readers are done with the scratch
                {
                    const int cnt = max((nvalid - hh + EPI_SPLIT - 1) / EPI_SPLIT, 0) * per;
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        const int i = lane + 32 * qq;
                        if (i < cnt) { sbias[i] = nb[qq]; scol[i] = nc[qq]; }
                    }
                }
                fetch_cols(t + t_step);                         // next tile's values: a whole tile of latency hiding
                // (2) request the residual of my first block and my row's LayerNorm statistics
                if (resrow && hh < nvalid) {
                    ldg256(resrow + hh * 32, *reinterpret_cast<uint32_t(*)[8]>(&rcur[0]));
                    if (EPI_COLS == 32) ldg256(resrow + hh * 32 + 16, *reinterpret_cast<uint32_t(*)[8]>(&rcur[8]));
                }
                // (slot statistics are merged right here: the loads are independent, so their latency is that of the single
                // {mean, rstd} load, and only two registers stay live across the accumulator wait)
                if (use_ln && valid) ln_ms = ln_row_stats(a, grow);
                __syncwarp();
            }
            mbar_wait(&tmem_full_bar[buf], aph);
            tc_fence_after();
            if (FLAVOR == 2) {
                if (hh == 0) epilogue_store<BN>(a, nt, split, grow, valid, trow);     // split-K partials / unaligned outputs
            } else if (EPI_COLS == 16 && !(VMV_DBG(a) & 2)) {
                // ---- 16-column steps: the same fused epilogue with half the live registers per thread (one tcgen05.ld.x16, one
                // 32 B residual / row-bias load and one 32 B store per step), so that more epilogue warps fit the register file
                const __half* rb = (a.rowbias && valid) ? a.rowbias + (grow / a.rows_per_group) * a.ld_rowbias : nullptr;
                const bool ln = use_ln;
                float ln_a = 1.f, ln_b = 0.f;
                if (ln && valid) { ln_a = ln_ms.y; ln_b = -ln_ms.y * ln_ms.x; }
                const bool has_b = a.bias != nullptr;
                const int nblk = max((nvalid - hh + EPI_SPLIT - 1) / EPI_SPLIT, 0);     // my 32-column blocks
                const int nsteps = 2 * nblk;
                float row_s = 0.f, row_q = 0.f, row_sh = 0.f;
                int row_cnt = 0;
                uint32_t v[16], g[16];
                auto col_of = [&](int st_) { return (hh + EPI_SPLIT * (st_ >> 1)) * EPI_BLK_COLS + (st_ & 1) * 16; };
                auto request = [&](int st_) {                   // TMEM -> registers for step st_ (value columns, + gate columns for GEGLU)
                    const int c_ = col_of(st_);
                    tmem_ld_32x32b_x16(trow + c_, v);
                    if (geglu) tmem_ld_32x32b_x16(trow + BN / 2 + c_, g);
                };
                if (nsteps > 0) request(0);
#pragma unroll 1
                for (int st_ = 0; st_ < nsteps; ++st_) {
                    const int j = st_ >> 1, half = st_ & 1;
                    const int c = col_of(st_);
                    uint32_t rbv[8];
                    if (rb) ldg256(rb + col0 + c, rbv);
                    tmem_ld_wait();
                    float x[16];
                    if (geglu) {
                        const float4* bv4 = reinterpret_cast<const float4*>(sbias + j * 64 + half * 16);    // value; gate at +32
                        const float4* cv4 = reinterpret_cast<const float4*>(scol + j * 64 + half * 16);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 b4 = bv4[i], g4 = bv4[8 + i], c4 = cv4[i], d4 = cv4[8 + i];
                            const float bv[4] = {b4.x, b4.y, b4.z, b4.w}, bg[4] = {g4.x, g4.y, g4.z, g4.w};
                            const float cv[4] = {c4.x, c4.y, c4.z, c4.w}, cg[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float val = fmaf(__uint_as_float(v[4 * i + e]), ln_a, fmaf(ln_b, cv[e], bv[e]));
                                const float gate = fmaf(__uint_as_float(g[4 * i + e]), ln_a, fmaf(ln_b, cg[e], bg[e]));
                                x[4 * i + e] = geglu_f(val, gate);
                            }
                        }
                    } else {
                        const float4* bv4 = reinterpret_cast<const float4*>(sbias + j * 32 + half * 16);
                        const float4* cv4 = reinterpret_cast<const float4*>(scol + j * 32 + half * 16);
                        if (ln) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float4 b4 = bv4[i], c4 = cv4[i];
                                const float bv[4] = {b4.x, b4.y, b4.z, b4.w}, cv[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) x[4 * i + e] = fmaf(__uint_as_float(v[4 * i + e]), ln_a, fmaf(ln_b, cv[e], bv[e]));
                            }
                        } else if (has_b) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float4 b4 = bv4[i];
                                const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) x[4 * i + e] = __uint_as_float(v[4 * i + e]) + bv[e];
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(v[i]);
                        }
                    }
                    if (st_ + 1 < nsteps) request(st_ + 1);     // the next step's accumulators travel while this one is finished
                    if (!geglu) {
                        if (rb) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float2 f = unpack_half2(rbv[i]);
                                x[2 * i] += f.x;
                                x[2 * i + 1] += f.y;
                            }
                        }

                    }
                    if (resrow) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float2 f = unpack_half2(rcur[i]);
                            x[2 * i] += f.x;
                            x[2 * i + 1] += f.y;
                        }
                        if (st_ + 1 < nsteps) ldg256(resrow + col_of(st_ + 1), *reinterpret_cast<uint32_t(*)[8]>(&rcur[0]));
                    }
                    if (a.rowstats) {
                        if (row_cnt == 0) row_sh = x[0];
                        row_cnt += 16;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float d = x[i] - row_sh;
                            row_s += d;
                            row_q = fmaf(d, d, row_q);
                        }
                    }
                    if (valid && !(VMV_DBG(a) & 1)) {
                        uint32_t o[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] = pack_half2(x[2 * i], x[2 * i + 1]);
                        stg256(drow + col0 + c, o);
                    }
                }
                if (a.rowstats && valid) {
                    float2 st = make_float2(0.f, 0.f);
                    if (row_cnt > 0) {
                        const float inv = 1.f / (float)row_cnt;
                        st.x = fmaf(row_s, inv, row_sh);
                        st.y = fmaxf(fmaf(-row_s * inv, row_s, row_q), 0.f);
                    }
                    a.rowstats[grow * a.rowstats_nslots + EPI_SPLIT * nt + hh] = st;
                }
            } else if (!(VMV_DBG(a) & 2)) {
                const __half* rb = (a.rowbias && valid) ? a.rowbias + (grow / a.rows_per_group) * a.ld_rowbias : nullptr;
                // folded LayerNorm as two FMAs per accumulator:  rstd*(acc - mean*colsum) + bias = acc*ln_a + (ln_b*colsum + bias)
                const bool ln = use_ln;
                float ln_a = 1.f, ln_b = 0.f;
                if (ln && valid) {
                    const float2 ms = ln_ms;
                    ln_a = ms.y;
                    ln_b = -ms.y * ms.x;
                }
                const bool has_b = a.bias != nullptr;
                int j = 0;
                // The accumulators of block i+2 are requested from TMEM as soon as block i's have been consumed into x[],
                // so the tcgen05.ld latency overlaps the residual / activation / pack / store part (VMV_GEMM_DEBUG 16: off).
                const bool ld_ahead = !(VMV_DBG(a) & 16);
                bool requested = false;
                uint32_t v[32];
                // LayerNorm partial statistics of my row over my blocks (rowstats): sums of (x - row_sh), (x - row_sh)^2 with
                // row_sh = my first value, so nothing cancels when the row's mean is large compared with its spread
                float row_s = 0.f, row_q = 0.f, row_sh = 0.f;
                int row_cnt = 0;
#pragma unroll 1
                for (int blk = hh; blk < nvalid; blk += EPI_SPLIT, ++j) {
                    const int c = blk * EPI_BLK_COLS;           // column inside the tile's output range
                    float x[32];
                    if (geglu) {
                        uint32_t g[32];
                        tmem_ld_32x32b_x16(trow + c, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
                        tmem_ld_32x32b_x16(trow + c + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
                        tmem_ld_32x32b_x16(trow + BN / 2 + c, *reinterpret_cast<uint32_t(*)[16]>(&g[0]));
                        tmem_ld_32x32b_x16(trow + BN / 2 + c + 16, *reinterpret_cast<uint32_t(*)[16]>(&g[16]));
                        tmem_ld_wait();
                        const float4* bv4 = reinterpret_cast<const float4*>(sbias + j * 64);    // [0,8) value, [8,16) gate
                        const float4* cv4 = reinterpret_cast<const float4*>(scol + j * 64);
                        if (ln) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 b4 = bv4[i], g4 = bv4[8 + i], c4 = cv4[i], d4 = cv4[8 + i];
                                const float bv[4] = {b4.x, b4.y, b4.z, b4.w}, bg[4] = {g4.x, g4.y, g4.z, g4.w};
                                const float cv[4] = {c4.x, c4.y, c4.z, c4.w}, cg[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float val = fmaf(__uint_as_float(v[4 * i + e]), ln_a, fmaf(ln_b, cv[e], bv[e]));
                                    const float gate = fmaf(__uint_as_float(g[4 * i + e]), ln_a, fmaf(ln_b, cg[e], bg[e]));
                                    x[4 * i + e] = geglu_f(val, gate);
                                }
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 b4 = bv4[i], g4 = bv4[8 + i];                  // zeros when there is no bias
                                const float bv[4] = {b4.x, b4.y, b4.z, b4.w}, bg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    x[4 * i + e] = geglu_f(__uint_as_float(v[4 * i + e]) + bv[e], __uint_as_float(g[4 * i + e]) + bg[e]);
                            }
                        }
                    } else {
                        if (!requested) {
                            tmem_ld_32x32b_x16(trow + c, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
                            tmem_ld_32x32b_x16(trow + c + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
                        }
                        uint32_t rbv[16];
                        if (rb) {
                            ldg256(rb + col0 + c, *reinterpret_cast<uint32_t(*)[8]>(&rbv[0]));
                            ldg256(rb + col0 + c + 16, *reinterpret_cast<uint32_t(*)[8]>(&rbv[8]));
                        }
                        tmem_ld_wait();
                        const float4* bv4 = reinterpret_cast<const float4*>(sbias + j * 32);
                        const float4* cv4 = reinterpret_cast<const float4*>(scol + j * 32);
                        if (ln) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 b4 = bv4[i], c4 = cv4[i];
                                const float bv[4] = {b4.x, b4.y, b4.z, b4.w}, cv[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    x[4 * i + e] = fmaf(__uint_as_float(v[4 * i + e]), ln_a, fmaf(ln_b, cv[e], bv[e]));
                            }
                        } else if (has_b) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 b4 = bv4[i];
                                const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) x[4 * i + e] = __uint_as_float(v[4 * i + e]) + bv[e];
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(v[i]);
                        }
                        requested = ld_ahead && blk + EPI_SPLIT < nvalid;
                        if (requested) {
                            tmem_ld_32x32b_x16(trow + c + EPI_SPLIT * EPI_BLK_COLS, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
                            tmem_ld_32x32b_x16(trow + c + EPI_SPLIT * EPI_BLK_COLS + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
                        }
                        if (rb) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float2 f = unpack_half2(rbv[i]);
                                x[2 * i] += f.x;
                                x[2 * i + 1] += f.y;
                            }
                        }
                        // (SiLU outputs -- only the tiny embedding MLPs -- take the generic epilogue: see the flavor dispatch)
                    }
                    if (resrow) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float2 f = unpack_half2(rcur[i]);
                            x[2 * i] += f.x;
                            x[2 * i + 1] += f.y;
                        }
                        if (blk + EPI_SPLIT < nvalid) {         // request the next block's residual now
                            ldg256(resrow + (blk + EPI_SPLIT) * 32, *reinterpret_cast<uint32_t(*)[8]>(&rcur[0]));
                            ldg256(resrow + (blk + EPI_SPLIT) * 32 + 16, *reinterpret_cast<uint32_t(*)[8]>(&rcur[8]));
                        }
                    }
                    if (a.rowstats) {
                        if (row_cnt == 0) row_sh = x[0];
                        row_cnt += 32;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float d = x[i] - row_sh;
                            row_s += d;
                            row_q = fmaf(d, d, row_q);
                        }
                    }
                    if (valid && !(VMV_DBG(a) & 1)) {
                        uint32_t o[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = pack_half2(x[2 * i], x[2 * i + 1]);
                        __half* dst = drow + col0 + c;
                        stg256(dst, *reinterpret_cast<uint32_t(*)[8]>(&o[0]));
                        stg256(dst + 16, *reinterpret_cast<uint32_t(*)[8]>(&o[8]));
                    }
                }
                if (a.rowstats && valid) {                       // one writer per slot: no atomics, no zero-initialised buffer
                    float2 st = make_float2(0.f, 0.f);
                    if (row_cnt > 0) {
                        const float inv = 1.f / (float)row_cnt;
                        st.x = fmaf(row_s, inv, row_sh);
                        st.y = fmaxf(fmaf(-row_s * inv, row_s, row_q), 0.f);
                    }
                    a.rowstats[grow * a.rowstats_nslots + EPI_SPLIT * nt + hh] = st;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(&tmem_empty_bar[buf]);
                else mbar_arrive_remote(&tmem_empty_bar[buf], 0);
            }
        }
    }

    if (a.sc_world) {
        // fused layout exchange: every CTA's remote stores are performed (system scope), the last CTA of this rank publishes
        // the epoch in every peer's flag line and waits for the peers' epochs -- same protocol as peer_exchange_kernel
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            const unsigned int prev = atomicAdd(a.sc_done, 1u);
            if (prev == gridDim.x - 1) {
                __threadfence();
                *a.sc_done = 0;
                const unsigned int e = *a.sc_epoch + 1;
                *a.sc_epoch = e;
                peer_publish(a.sc_flags, a.sc_world, a.sc_rank, e);
                if (!a.sc_nowait) peer_wait_all(a.sc_flags, a.sc_world, a.sc_rank, e);
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_2sm<TCOLS>(tmem_base);
}

// Reduce split-K partials and apply the (non-GEGLU) epilogue.  One thread per 8 output columns.  With `scatter` the rows go
// to the peers' tensors of the other sharding layout (coalesced 16 B stores) and the kernel ends with the flag rendezvous.
__global__ void splitk_finish_kernel(const float* __restrict__ partial, int splits, int M, int N, GemmArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const int nvec = N / 8;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (long long)M * nvec) {
        const long long row = idx / nvec;
        const int n = (int)(idx % nvec) * 8;
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = 0.f;
        for (int s = 0; s < splits; ++s) {
            const float4* p = reinterpret_cast<const float4*>(partial + ((long long)s * M + row) * N + n);
            float4 u = p[0], w = p[1];
            x[0] += u.x; x[1] += u.y; x[2] += u.z; x[3] += u.w;
            x[4] += w.x; x[5] += w.y; x[6] += w.z; x[7] += w.w;
        }
        if (a.ln_stats) {
            const float2 ms = ln_row_stats(a, row);
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = ms.y * (x[j] - ms.x * a.ln_colsum[n + j]);
        }
        if (a.bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] += a.bias[n + j];
        }
        if (a.rowbias) {
            const __half* rb = a.rowbias + (row / a.rows_per_group) * a.ld_rowbias + n;
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] += __half2float(rb[j]);
        }
        if (a.act == VMV_ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = silu_f(x[j]);
        }
        if (a.residual) {
            const __half* rp = a.residual + row * a.ldr + n;
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] += __half2float(rp[j]);
        }
        uint4 o = make_uint4(pack_half2(x[0], x[1]), pack_half2(x[2], x[3]), pack_half2(x[4], x[5]),
                             pack_half2(x[6], x[7]));
        __half* drow = a.sc_world ? scatter_row(a, row) : a.D + row * a.ldd;
        *reinterpret_cast<uint4*>(drow + n) = o;
    }
    if (a.sc_world) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            const unsigned int prev = atomicAdd(a.sc_done, 1u);
            if (prev == gridDim.x - 1) {
                __threadfence();
                *a.sc_done = 0;
                const unsigned int e = *a.sc_epoch + 1;
                *a.sc_epoch = e;
                peer_publish(a.sc_flags, a.sc_world, a.sc_rank, e);
                if (!a.sc_nowait) peer_wait_all(a.sc_flags, a.sc_world, a.sc_rank, e);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// fp16 tensor map, 128B swizzle, inner box = 64 elements.  dims/strides innermost first; strides in bytes
// for dims 1..rank-1.
static int make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                    const cuuint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B,
                    const cuuint32_t* elem_strides = nullptr) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return VMV_ERR_CUDA;
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (elem_strides)
        for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] "
                  "stride0 %llu base %p",
                  (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                  (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                  rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0,
                  (unsigned long long)(rank > 1 ? strides[0] : 0), base);
        return VMV_ERR_CUDA;
    }
    return VMV_OK;
}

// exported to the other translation units (attention_tc.cu): fp16 tiled tensor map, inner box = 64 elements
int make_map_generic(CUtensorMap* m, const void* base, int rank, const unsigned long long* dims,
                     const unsigned long long* strides_bytes, const unsigned* box, int swizzle_bytes) {
    cuuint64_t d[5], st[4];
    cuuint32_t b[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
    return make_map(m, base, rank, d, st, b, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int BN, int STAGES>
static int launch_instance(const CUtensorMap& tA1, const CUtensorMap& tA2, const CUtensorMap& tW, const GemmArgs& a,
                           dim3 grid, cudaStream_t st) {
    using L = SmemLayout<BN, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             L::DYN_BYTES);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(smem=%d) failed: %s", L::DYN_BYTES, cudaGetErrorString(e));
            return VMV_ERR_CUDA;
        }
        attr_set = true;
    }
    launch_kernel(gemm_tc_kernel<BN, STAGES>, grid, dim3(192), L::DYN_BYTES, st, tA1, tA2, tW, a);
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_gemm");
    return VMV_OK;
}

template <int BN, int STAGES, int FLAVOR>
static int launch_instance2(const CUtensorMap& tA1, const CUtensorMap& tA2, const CUtensorMap& tW, const GemmArgs& a,
                            int m_pairs, int n_tiles, int splits, cudaStream_t st) {
    using L = SmemLayout2<BN, STAGES, FLAVOR>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<BN, STAGES, FLAVOR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             L::DYN_BYTES);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(v2 smem=%d) failed: %s", L::DYN_BYTES, cudaGetErrorString(e));
            return VMV_ERR_CUDA;
        }
        attr_set = true;
    }
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
    }
    const long long total = (long long)m_pairs * n_tiles * splits;
    int clusters = num_sms / 2;
    if (total < clusters) clusters = (int)total;
    launch_kernel(gemm_tc2_kernel<BN, STAGES, FLAVOR>, dim3(2 * clusters), dim3(V2_THREADS), L::DYN_BYTES, st, tA1, tA2, tW, a, m_pairs,
                  n_tiles, splits);
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_gemm (cta_group::2)");
    return VMV_OK;
}

static int pick_block_n(const vmv_gemm_params* p, int variant) {
    if (p->block_n) return p->block_n;
    const int N = p->N;
    if (variant == 2) {
        if (p->act == VMV_ACT_GEGLU) return (N % 256 == 0) ? 256 : 128;   // BN/2 must be a multiple of 32
        // tcgen05.mma in SS mode is operand-bandwidth bound: wider N amortises the A reads (measured 759 / 874 / 1254
        // TFLOP/s MMA-only for N = 128 / 160 / 256, profiles/r1_mainloop_isolation.log), and fewer N tiles re-read A
        // less often.  256-wide tiles with a ragged last tile for everything wider than 320 columns.
        if (N > 320) return 256;
        if (N % 160 == 0) return 160;
        return 128;
    }
    if (p->act == VMV_ACT_GEGLU) return (N % 160 == 0) ? 160 : 128;
    if (N % 160 == 0) return 160;
    if (N % 128 == 0) return 128;
    if (N <= 64) return 64;
    return 128;
}

struct Plan {
    GemmArgs a;
    int bn, stages, splits, m_tiles, n_tiles, variant;
};

static int default_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VMV_GEMM_VARIANT");      // 1 = 1-CTA per-tile kernel, 2 = persistent CTA-pair kernel
        v = (e && e[0] == '1') ? 1 : 2;
    }
    return v;
}

static int make_plan(const vmv_gemm_params* p, Plan* pl) {
    GemmArgs& a = pl->a;
    memset(&a, 0, sizeof(a));
    VMV_CHECK_ARG(p->M > 0 && p->N > 0, "vmv_gemm: M, N must be positive (M=%d N=%d)", p->M, p->N);
    VMV_CHECK_ARG(p->N % 16 == 0, "vmv_gemm: N=%d must be a multiple of 16", p->N);
    VMV_CHECK_ARG(p->A1 && p->W && (p->D || p->scatter), "vmv_gemm: null A1/W/D");
    VMV_CHECK_ARG(p->K1 > 0 && p->K1 % BK == 0 && p->K2 % BK == 0,
                  "vmv_gemm: K1=%d, K2=%d must be multiples of %d", p->K1, p->K2, BK);
    VMV_CHECK_ARG(p->lda1 % 8 == 0 && p->ldw % 8 == 0 && p->ldd % 8 == 0, "vmv_gemm: leading dims must be multiples of 8");
    a.mode = p->mode;
    a.M = p->M;
    a.N = p->N;
    a.act = p->act;
    a.n_out = p->act == VMV_ACT_GEGLU ? p->N / 2 : p->N;
    a.D = static_cast<__half*>(p->D);
    a.ldd = p->ldd;
    a.bias = p->bias;
    a.rowbias = static_cast<const __half*>(p->rowbias);
    a.ld_rowbias = p->ld_rowbias;
    a.rows_per_group = p->rows_per_group > 0 ? p->rows_per_group : 1;
    a.residual = static_cast<const __half*>(p->residual);
    a.ldr = p->ldr;
    a.w_static = p->w_static;
    {
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("VMV_GEMM_DEBUG"); dbg = e ? atoi(e) : 0; }
        a.dbg = dbg;
    }
    a.ln_stats = static_cast<const float2*>(p->ln_stats);
    a.ln_colsum = p->ln_colsum;
    a.ln_eps = p->ln_eps;
    if (p->ln_stats && p->ln_stats_src_n != 0) {
        VMV_CHECK_ARG(p->ln_stats_src_n > 0 && p->ln_stats_src_bn >= 32 && p->ln_stats_src_bn % 32 == 0,
                      "vmv_gemm: ln_stats_src_n / ln_stats_src_bn must describe the producing GEMM's tiling");
        a.ln_src_n = p->ln_stats_src_n;
        a.ln_src_bn = p->ln_stats_src_bn;
        a.ln_nslots = EPI_SPLIT * ((a.ln_src_n + a.ln_src_bn - 1) / a.ln_src_bn);
        VMV_CHECK_ARG(a.ln_nslots <= LN_MAX_SLOTS, "vmv_gemm: %d LayerNorm statistic slots (max %d)", a.ln_nslots, LN_MAX_SLOTS);
    }
    a.rowstats = static_cast<float2*>(p->rowstats_out);
    if (p->scatter != nullptr) {
        const vmv_gemm_scatter* sc = p->scatter;
        VMV_CHECK_ARG(sc->world >= 1 && sc->world <= VMV_PEER_MAX_RANKS && sc->rank >= 0 && sc->rank < sc->world, "vmv_gemm scatter: bad world/rank");
        VMV_CHECK_ARG(sc->direction == 0 || sc->direction == 1, "vmv_gemm scatter: direction must be 0 or 1");
        VMV_CHECK_ARG(sc->B > 0 && sc->Fl > 0 && sc->HWl > 0 && (long long)sc->B * sc->Fl * sc->HWl * sc->world == p->M,
                      "vmv_gemm scatter: B*Fl*HWl*world must equal M");
        VMV_CHECK_ARG(sc->epoch && sc->done && p->ldd % 16 == 0, "vmv_gemm scatter: null epoch/done, or ldd not a multiple of 16");
        a.sc_world = sc->world; a.sc_rank = sc->rank; a.sc_dir = sc->direction; a.sc_nowait = sc->nowait;
        a.sc_B = sc->B; a.sc_Fl = sc->Fl; a.sc_HWl = sc->HWl;
        for (int q = 0; q < sc->world; ++q) {
            VMV_CHECK_ARG(sc->dst[q] && sc->flags[q] && (reinterpret_cast<uintptr_t>(sc->dst[q]) & 31) == 0, "vmv_gemm scatter: bad dst/flags for rank %d", q);
            a.sc_dst[q] = static_cast<__half*>(sc->dst[q]);
            a.sc_flags[q] = static_cast<unsigned int*>(sc->flags[q]);
        }
        a.sc_epoch = static_cast<unsigned int*>(sc->epoch);
        a.sc_done = static_cast<unsigned int*>(sc->done);
    }
    VMV_CHECK_ARG((p->ln_stats == nullptr) == (p->ln_colsum == nullptr), "vmv_gemm: ln_stats and ln_colsum go together");
    if (p->rowbias) VMV_CHECK_ARG(p->ld_rowbias % 8 == 0, "vmv_gemm: ld_rowbias must be a multiple of 8");
    if (p->residual) VMV_CHECK_ARG(p->ldr % 8 == 0, "vmv_gemm: ldr must be a multiple of 8");

    if (p->mode == VMV_GEMM_LINEAR) {
        if (p->K2 > 0) VMV_CHECK_ARG(p->A2 && p->lda2 % 8 == 0, "vmv_gemm: K2>0 needs A2 with lda2 %% 8 == 0");
        a.nkb1 = p->K1 / BK;
        a.nkb = (p->K1 + p->K2) / BK;
        a.cblocks = a.nkb;
        pl->m_tiles = (p->M + BM - 1) / BM;
    } else if (p->mode == VMV_GEMM_CONV3X3) {
        VMV_CHECK_ARG(p->K2 == 0, "vmv_gemm: conv modes take a single source");
        const int H = p->H, W = p->Wd, NF = p->B * p->F;
        VMV_CHECK_ARG(H > 0 && W > 0 && NF > 0 && (long long)NF * H * W == p->M, "vmv_gemm conv3x3: M != B*F*H*W");
        VMV_CHECK_ARG(W >= 128 ? (W % 128 == 0) : (128 % W == 0), "vmv_gemm conv3x3: W=%d must divide or be a multiple of 128", W);
        a.bw = W >= 128 ? 128 : W;
        const int rem = 128 / a.bw;
        VMV_CHECK_ARG(H >= rem ? (H % rem == 0) : (rem % H == 0), "vmv_gemm conv3x3: H=%d incompatible with 128-row tiles", H);
        a.bh = H >= rem ? rem : H;
        a.bf = 128 / (a.bw * a.bh);
        a.tiles_w = W / a.bw;
        a.tiles_h = H / a.bh;
        a.F = NF; a.H = H; a.W = W; a.HW = H * W;
        a.cblocks = p->K1 / BK;
        a.nkb = 9 * a.cblocks;
        a.nkb1 = a.nkb;
        pl->m_tiles = a.tiles_w * a.tiles_h * ((NF + a.bf - 1) / a.bf);
    } else if (p->mode == VMV_GEMM_CONV3X3_S2 || p->mode == VMV_GEMM_UPCONV3X3) {
        // geometry arguments describe the INPUT images; tiles are laid over the output grid (stride 2: H/2 x W/2) or over the
        // input grid (nearest-x2 upsample + conv: every input position produces the 4 output phases)
        VMV_CHECK_ARG(p->K2 == 0, "vmv_gemm: conv modes take a single source");
        const bool s2 = p->mode == VMV_GEMM_CONV3X3_S2;
        const int Hin = p->H, Win = p->Wd, NF = p->B * p->F;
        VMV_CHECK_ARG(Hin > 0 && Win > 0 && NF > 0 && (!s2 || (Hin % 2 == 0 && Win % 2 == 0)), "vmv_gemm strided / upsampling conv: bad H, W");
        const int H = s2 ? Hin / 2 : Hin, W = s2 ? Win / 2 : Win;             // tile grid
        VMV_CHECK_ARG((long long)NF * H * W * (s2 ? 1 : 4) == p->M, "vmv_gemm strided / upsampling conv: M does not match the output size");
        VMV_CHECK_ARG(W >= 128 ? (W % 128 == 0) : (128 % W == 0), "vmv_gemm conv: tile-grid W=%d must divide or be a multiple of 128", W);
        a.bw = W >= 128 ? 128 : W;
        const int rem = 128 / a.bw;
        VMV_CHECK_ARG(H >= rem ? (H % rem == 0) : (rem % H == 0), "vmv_gemm conv: tile-grid H=%d incompatible with 128-row tiles", H);
        a.bh = H >= rem ? rem : H;
        a.bf = 128 / (a.bw * a.bh);
        a.tiles_w = W / a.bw;
        a.tiles_h = H / a.bh;
        a.F = NF; a.H = H; a.W = W; a.HW = H * W;
        a.cblocks = p->K1 / BK;
        a.nkb = (s2 ? 9 : 4) * a.cblocks;
        a.nkb1 = a.nkb;
        pl->m_tiles = a.tiles_w * a.tiles_h * ((NF + a.bf - 1) / a.bf);
    } else if (p->mode == VMV_GEMM_TCONV3) {
        VMV_CHECK_ARG(p->K2 == 0, "vmv_gemm: conv modes take a single source");
        const int HW = p->H * p->Wd;
        VMV_CHECK_ARG(HW > 0 && p->B > 0 && p->F > 0 && (long long)p->B * p->F * HW == p->M, "vmv_gemm tconv3: M != B*F*H*W");
        VMV_CHECK_ARG(HW >= 128 ? (HW % 128 == 0) : (128 % HW == 0), "vmv_gemm tconv3: H*W=%d must divide or be a multiple of 128", HW);
        a.bp = HW >= 128 ? 128 : HW;
        a.bf = 128 / a.bp;
        a.tiles_p = HW / a.bp;
        a.tiles_per_sample = a.tiles_p * ((p->F + a.bf - 1) / a.bf);
        a.F = p->F; a.HW = HW; a.H = p->H; a.W = p->Wd;
        a.cblocks = p->K1 / BK;
        a.nkb = 3 * a.cblocks;
        a.nkb1 = a.nkb;
        pl->m_tiles = p->B * a.tiles_per_sample;
    } else {
        set_error("vmv_gemm: unknown mode %d", p->mode);
        return VMV_ERR_INVALID;
    }
    a.mt_phase = pl->m_tiles;
    pl->variant = p->variant ? p->variant : default_variant();
    VMV_CHECK_ARG(pl->variant == 1 || pl->variant == 2, "vmv_gemm: variant=%d unsupported", pl->variant);
    if (p->mode == VMV_GEMM_CONV3X3_S2 || p->mode == VMV_GEMM_UPCONV3X3)
        VMV_CHECK_ARG(pl->variant == 2, "vmv_gemm: the strided / upsampling conv modes need the CTA-pair kernel");
    if (p->mode == VMV_GEMM_UPCONV3X3)
        VMV_CHECK_ARG(p->split_k <= 1, "vmv_gemm: the upsampling conv mode does not take split-K");
    if (pl->variant == 2 && p->block_n == 64) pl->variant = 1;      // 64-wide tiles exist only in the 1-CTA kernel
    pl->bn = pick_block_n(p, pl->variant);
    VMV_CHECK_ARG(pl->bn == 64 || pl->bn == 128 || pl->bn == 160 || pl->bn == 256, "vmv_gemm: block_n=%d unsupported", pl->bn);
    if (p->act == VMV_ACT_GEGLU)
        VMV_CHECK_ARG(p->N % pl->bn == 0 && (pl->bn / 2) % (pl->variant == 2 ? 32 : 16) == 0,
                      "vmv_gemm: GEGLU needs N %% block_n == 0 and block_n in {128, 256} for the CTA-pair kernel");
    pl->n_tiles = (p->N + pl->bn - 1) / pl->bn;
    // (a ragged last tile of the CTA-pair kernel is (N % bn) wide: a multiple of 16, so each CTA's half is 8-row aligned)
    pl->splits = p->split_k > 1 ? p->split_k : 1;
    if (pl->splits > a.nkb) pl->splits = a.nkb;
    if (p->act == VMV_ACT_GEGLU) pl->splits = 1;
    a.kb_per_split = (a.nkb + pl->splits - 1) / pl->splits;
    pl->splits = (a.nkb + a.kb_per_split - 1) / a.kb_per_split;
    pl->stages = p->stages;
    return VMV_OK;
}

}  // namespace vmv

using namespace vmv;

extern "C" int vmv_gemm_epilogue_split(void) { return EPI_SPLIT; }

extern "C" int vmv_gemm_block_n(const vmv_gemm_params* p) {
    if (!p) return -1;
    return pick_block_n(p, p->variant ? p->variant : default_variant());
}

extern "C" int64_t vmv_gemm_workspace_bytes(const vmv_gemm_params* p) {
    Plan pl;
    if (make_plan(p, &pl) != VMV_OK) return -1;
    if (pl.splits <= 1) return 0;
    return (int64_t)pl.splits * p->M * p->N * (int64_t)sizeof(float);
}

extern "C" int vmv_gemm(const vmv_gemm_params* p, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Plan pl;
    int rc = make_plan(p, &pl);
    if (rc != VMV_OK) return rc;
    GemmArgs& a = pl.a;
    const int BN = pl.bn;

    CUtensorMap tA1, tA2, tW;
    const long long Ktot = p->mode == VMV_GEMM_LINEAR ? (long long)p->K1 + p->K2
                           : (p->mode == VMV_GEMM_CONV3X3 || p->mode == VMV_GEMM_CONV3X3_S2) ? 9LL * p->K1
                           : p->mode == VMV_GEMM_UPCONV3X3 ? 4LL * p->K1 : 3LL * p->K1;
    const long long w_rows = p->mode == VMV_GEMM_UPCONV3X3 ? 4LL * p->N : p->N;      // one weight set per output phase
    VMV_CHECK_ARG(p->ldw >= Ktot, "vmv_gemm: ldw=%lld < Ktot=%lld", (long long)p->ldw, Ktot);
    {
        cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)w_rows};
        cuuint64_t strides[1] = {(cuuint64_t)p->ldw * 2};
        cuuint32_t box[2] = {BK, (cuuint32_t)(pl.variant == 2 ? BN / 2 : BN)};   // a CTA pair splits the W rows
        if ((rc = make_map(&tW, p->W, 2, dims, strides, box)) != VMV_OK) return rc;
    }
    if (p->mode == VMV_GEMM_LINEAR) {
        cuuint64_t dims[2] = {(cuuint64_t)p->K1, (cuuint64_t)p->M};
        cuuint64_t strides[1] = {(cuuint64_t)p->lda1 * 2};
        cuuint32_t box[2] = {BK, BM};
        if ((rc = make_map(&tA1, p->A1, 2, dims, strides, box)) != VMV_OK) return rc;
        if (p->K2 > 0) {
            cuuint64_t dims2[2] = {(cuuint64_t)p->K2, (cuuint64_t)p->M};
            cuuint64_t strides2[1] = {(cuuint64_t)p->lda2 * 2};
            if ((rc = make_map(&tA2, p->A2, 2, dims2, strides2, box)) != VMV_OK) return rc;
        } else {
            tA2 = tA1;
        }
    } else if (p->mode == VMV_GEMM_CONV3X3 || p->mode == VMV_GEMM_UPCONV3X3) {
        const cuuint64_t ld = (cuuint64_t)p->lda1 * 2;
        cuuint64_t dims[4] = {(cuuint64_t)p->K1, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.F};
        cuuint64_t strides[3] = {ld, ld * a.W, ld * a.W * a.H};
        cuuint32_t box[4] = {BK, (cuuint32_t)a.bw, (cuuint32_t)a.bh, (cuuint32_t)a.bf};
        if ((rc = make_map(&tA1, p->A1, 4, dims, strides, box)) != VMV_OK) return rc;
        tA2 = tA1;
    } else if (p->mode == VMV_GEMM_CONV3X3_S2) {
        // stride-2 window straight from the input image: the box spans 2*bw x 2*bh input pixels and is traversed with
        // element strides {1, 2, 2, 1}, so bw x bh pixels land in smem -- in the (h, w) row order of the output tile
        const cuuint64_t ld = (cuuint64_t)p->lda1 * 2;
        const cuuint64_t Win = 2ull * a.W, Hin = 2ull * a.H;
        cuuint64_t dims[4] = {(cuuint64_t)p->K1, Win, Hin, (cuuint64_t)a.F};
        cuuint64_t strides[3] = {ld, ld * Win, ld * Win * Hin};
        cuuint32_t box[4] = {BK, 2u * (cuuint32_t)a.bw, 2u * (cuuint32_t)a.bh, (cuuint32_t)a.bf};
        cuuint32_t estr[4] = {1, 2, 2, 1};
        if ((rc = make_map(&tA1, p->A1, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, estr)) != VMV_OK) return rc;
        tA2 = tA1;
    } else {
        const cuuint64_t ld = (cuuint64_t)p->lda1 * 2;
        cuuint64_t dims[4] = {(cuuint64_t)p->K1, (cuuint64_t)a.HW, (cuuint64_t)a.F, (cuuint64_t)p->B};
        cuuint64_t strides[3] = {ld, ld * a.HW, ld * a.HW * a.F};
        cuuint32_t box[4] = {BK, (cuuint32_t)a.bp, (cuuint32_t)a.bf, 1};
        if ((rc = make_map(&tA1, p->A1, 4, dims, strides, box)) != VMV_OK) return rc;
        tA2 = tA1;
    }

    if (pl.splits > 1) {
        const int64_t need = (int64_t)pl.splits * p->M * p->N * (int64_t)sizeof(float);
        VMV_CHECK_ARG(p->workspace && p->workspace_bytes >= need,
                      "vmv_gemm: split_k=%d needs %lld workspace bytes (have %lld)", pl.splits, (long long)need,
                      (long long)p->workspace_bytes);
        a.partial = static_cast<float*>(p->workspace);
    }

    if (pl.variant == 2) {
        const int m_pairs = ((pl.m_tiles + 1) / 2) * (p->mode == VMV_GEMM_UPCONV3X3 ? 4 : 1);
        if (pl.splits <= 1) {
            VMV_CHECK_ARG(!(p->act == VMV_ACT_GEGLU && p->residual), "vmv_gemm: GEGLU with residual is not supported");
            auto al32 = [](const void* ptr, long long ld) { return ptr == nullptr || ((reinterpret_cast<uintptr_t>(ptr) & 31) == 0 && ld % 16 == 0); };
            a.fast_epi = (a.n_out % EPI_BLK_COLS == 0) && al32(p->D, p->ldd) && al32(p->residual, p->ldr) &&
                                 al32(p->rowbias, p->ld_rowbias)
                             ? 1 : 0;
        }
        a.rowstats_nslots = EPI_SPLIT * pl.n_tiles;
        if (a.sc_world && pl.splits <= 1 && !(a.fast_epi && p->act != VMV_ACT_GEGLU)) {
            set_error("vmv_gemm: scatter needs the CTA-pair kernel's register epilogue (no GEGLU, N %% 32 == 0, 32 B aligned rows) or split-K");
            return VMV_ERR_UNSUPPORTED;
        }
        if (a.rowstats && !(a.fast_epi && p->act != VMV_ACT_GEGLU)) {
            set_error("vmv_gemm: rowstats_out needs the CTA-pair kernel's register epilogue (no split-K, no GEGLU, N %% 32 == 0, "
                      "32 B aligned rows)");
            return VMV_ERR_UNSUPPORTED;
        }
        GemmArgs ak = a;
        if (pl.splits > 1) ak.sc_world = 0;             // split-K: the finish kernel scatters (and runs the rendezvous), not this one
        {
            // A-resident schedule: small K (every K block of an m tile fits the ring), several N tiles per m tile, and enough
            // tiles that contiguous runs per cluster balance as well as round-robin does.  OFF by default (VMV_GEMM_ARES=1):
            // measured neutral on B200 (19.70 vs 19.77 frames/s, profiles/r2_ab_ares.txt) -- the K <= 512 GEMMs are bound by
            // the latency of the 8-warp epilogue (ncu: 26 % issue slots busy, 2.5 warps per scheduler), not by operand traffic.
            static int ares = -1;
            if (ares < 0) { const char* e = getenv("VMV_GEMM_ARES"); ares = (e && e[0] == '1') ? 1 : 0; }
            const int ring = BN == 256 ? 6 : 8;
            const long long total = (long long)m_pairs * pl.n_tiles;
            if (ares && p->mode == VMV_GEMM_LINEAR && pl.splits <= 1 && a.nkb <= ring && pl.n_tiles >= 2 && total >= 2 * 74) ak.a_res = 1;
        }
        if (p->act == VMV_ACT_SILU && !ak.rowstats && !ak.sc_world) ak.fast_epi = 0;      // rare (embedding MLPs): generic epilogue
        if (p->act == VMV_ACT_SILU && ak.fast_epi) { set_error("vmv_gemm: SiLU cannot be combined with rowstats_out / scatter"); return VMV_ERR_UNSUPPORTED; }
        const int flavor = !ak.fast_epi ? 2 : (p->act == VMV_ACT_GEGLU ? 1 : ((ak.ln_stats ? 4 : 0) | (ak.residual ? 8 : 0)));
#define VMV_L2(BN_, ST_, FL_) rc = launch_instance2<BN_, ST_, FL_>(tA1, tA2, tW, ak, m_pairs, pl.n_tiles, pl.splits, st)
#define VMV_L2_FLAVORS(BN_, ST_)                                                                      \
        switch (flavor) {                                                                             \
            case 0: VMV_L2(BN_, ST_, 0); break;                                                       \
            case 4: VMV_L2(BN_, ST_, 4); break;                                                       \
            case 8: VMV_L2(BN_, ST_, 8); break;                                                       \
            case 12: VMV_L2(BN_, ST_, 12); break;                                                     \
            case 1: VMV_L2(BN_, ST_, 1); break;                                                       \
            default: VMV_L2(BN_, ST_, 2); break;                                                      \
        }
        if (BN == 160 && flavor == 1) { set_error("vmv_gemm: GEGLU needs block_n 128 or 256"); return VMV_ERR_UNSUPPORTED; }
        if (BN == 128) { VMV_L2_FLAVORS(128, 8) }
        else if (BN == 160) { VMV_L2_FLAVORS(160, 8) }
        else { VMV_L2_FLAVORS(256, 6) }
#undef VMV_L2_FLAVORS
#undef VMV_L2
    } else {
        if (a.rowstats || a.sc_world) {
            set_error("vmv_gemm: rowstats_out / scatter are not available in the one-tile-per-CTA kernel (variant 1)");
            return VMV_ERR_UNSUPPORTED;
        }
        dim3 grid(pl.n_tiles, pl.m_tiles, pl.splits);
        // stage count: deep ring for one CTA/SM; the shallow ring leaves room for two co-resident CTAs so one
        // CTA's epilogue overlaps the other's main loop.
        int stages = pl.stages;
        if (stages == 0) stages = 3;
#define VMV_LAUNCH(BN_, ST_) rc = launch_instance<BN_, ST_>(tA1, tA2, tW, a, grid, st)
        if (BN == 64) {
            if (stages <= 4) VMV_LAUNCH(64, 4); else VMV_LAUNCH(64, 8);
        } else if (BN == 128) {
            if (stages <= 3) VMV_LAUNCH(128, 3); else VMV_LAUNCH(128, 6);
        } else if (BN == 160) {
            if (stages <= 3) VMV_LAUNCH(160, 3); else VMV_LAUNCH(160, 6);
        } else {
            VMV_LAUNCH(256, 4);
        }
#undef VMV_LAUNCH
    }
    if (rc != VMV_OK) return rc;

    if (pl.splits > 1) {
        const long long nthreads = (long long)p->M * (p->N / 8);
        const int tb = 256;
        launch_kernel(splitk_finish_kernel, dim3((unsigned)((nthreads + tb - 1) / tb)), dim3(tb), 0, st, (const float*)a.partial, pl.splits,
                      p->M, p->N, a);
        count_launch();
        VMV_CUDA_LAUNCH_CHECK("vmv_gemm split-K finish");
    }
    return VMV_OK;
}
