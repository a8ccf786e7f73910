// tcgen05 flash attention:  O = softmax(Q K^T * scale) V,  head_dim 64, fp16 in/out, fp32 softmax.
//   mode 0  long sequences (spatial self-attention 256..4096 tokens, text cross-attention), batches contiguous in memory;
//   mode 1  SHORT contiguous sequences (spatial / cross attention at the 8x8 and 4x4 levels: 64 or 16 tokens): 128 / n
//           consecutive sequences are packed into one 128-row tile; self-attention masks the score tile block-diagonally,
//           cross-attention needs no mask (the packed frames share the text K/V);
//   mode 2  STRIDED short sequences (temporal attention: a sequence = the F frames of one pixel, frame stride = HW rows):
//           G = 128 / F pixels are packed into one tile by a 4-D TMA box (64 d, G pixels, F frames, 1 sample); tile row
//           r = f * G + p, so query r may attend key j iff j % G == r % G (and j < G * F).
// The packed modes spend up to 128 / n times the minimum MMA / exp work, which is free here: these shapes are bound by the
// HBM traffic of Q, K, V, O, and the tile now arrives in a handful of TMA boxes instead of per-thread strided loads.
//
// Persistent: one CTA per SM walks tiles of 128 query rows of one (batch, head); K/V are streamed in 128-key chunks
// through a 4-stage TMA ring.  All pipeline counters run across tiles, so the next tile's Q / first K/V chunks are
// loaded and its first two score tiles computed while the softmax warps finish and store the current tile.
//   warp 0      TMA producer (Q per tile, then K/V chunks; 128B-swizzled boxes straight out of the fused QKV GEMM output)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer
//                 S_g = Q K_g^T      M128 x N128 x K64   (both operands K-major)            -> TMEM cols [128 (g&1), +128)
//                 O_g = P_g V_g      M128 x N64  x K128  (P K-major from smem, V MN-major)  -> TMEM cols [256,320)
//   warps 2..9  softmax: two warps per TMEM lane quarter; a query row is shared by one thread of each, which takes 64 of
//               the chunk's 128 keys (single TMEM pass, scores stay in registers) and 32 of the 64 output columns.  The
//               pair exchanges its partial row max through smem (named barrier per quarter); P is written to smem in the
//               UMMA K-major 128B-swizzle layout (one 64-key atom per warp); the running output lives in registers and is
//               corrected per chunk (O_j is read back from TMEM, never accumulated there).
// Two score buffers: S_{g+1} (and S_{g+2} as soon as softmax_g releases its buffer) are computed while the softmax warps
// work on chunk g.  The kernel is exponent-bound (16 MUFU/clk/SM: 1024 cycles per 128x128 chunk vs 512 cycles of MMA),
// hence two warps per scheduler on the softmax side to hide the TMEM / barrier latencies.
#include "common.cuh"
#include "../../include/videomv_b200.h"

#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace vmv {

void count_launch(int n = 1);
int make_map_generic(CUtensorMap* m, const void* base, int rank, const unsigned long long* dims,
                     const unsigned long long* strides_bytes, const unsigned* box, int swizzle_bytes);

constexpr int AT_BM = 128, AT_BN = 128, AT_D = 64;
constexpr int AT_TILE = 128 * 64 * 2;                     // 16 KiB: one [128][64] fp16 tile
constexpr int AT_KV = 4;                                  // K/V ring depth (a 32 KiB chunk is ~1 chunk period of TMA latency)
constexpr int AT_OFF_Q = 0;
constexpr int AT_OFF_K = AT_TILE;
constexpr int AT_OFF_V = AT_OFF_K + AT_KV * AT_TILE;
constexpr int AT_OFF_P = AT_OFF_V + AT_KV * AT_TILE;      // [128][128] fp16 = two K atoms of 64 keys
constexpr int AT_OFF_X = AT_OFF_P + 2 * AT_TILE;          // float [2 parity][2 halves][128 rows] max exchange + [2][128] sums
constexpr int AT_OFF_BAR = AT_OFF_X + 6 * 128 * 4;
constexpr int AT_SMEM = AT_OFF_BAR + 16 * 8 + 16;
constexpr int AT_THREADS = 320;
constexpr int AT_TMEM_COLS = 512;                         // S0: [0,128)  S1: [128,256)  O: [256,320)
constexpr int AT_OCOL = 256;

struct AttTcArgs {
    int nq, nk, kv_group;
    int nqt, heads, ntiles;                               // tile = (batch, head, 128-row block), q block fastest
    int mode;                                             // 0 long sequences, 1 packed contiguous, 2 packed strided (see the header)
    int G, seq_shift;                                     // mode 2: pixels per tile; mode 1: log2(tokens per sequence)
    int self_mask;                                        // mode 1: block-diagonal mask (self-attention) or none (cross)
    int rows_total;                                       // mode 1: outer * nq rows in the flattened row space
    int HW;                                               // mode 2: pixels (sequences) per sample
    long long o_bs_inner;                                 // mode 2: output stride between pixels
    float scale_log2;                                     // scale * log2(e)
    __half* o;
    long long o_bs, o_rs;                                 // elements
};

// MN-major (N contiguous) 128B-swizzled B operand: V chunk [128 keys][64 d].  One 64-wide MN atom; 8-key groups 1024 B
// apart (SBO); a K step of 16 keys advances the start address by 2048 B.  (cute::UMMA::SmemDescriptor, DeepGEMM
// make_umma_desc<Major::MN>: SBO = 8 * 64 * 2, LBO = BLOCK_K * 64 * 2.)
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr, uint32_t block_k) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((block_k * 128u) >> 4) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttTcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];     // 128B-swizzled tiles need 1024 B alignment (checked below)
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT_OFF_BAR);
    uint64_t* q_full = bars;            // [1]
    uint64_t* kv_full = bars + 1;       // [AT_KV]
    uint64_t* kv_empty = bars + 5;      // [AT_KV]
    uint64_t* s_full = bars + 9;        // [2]  scores of chunk g are in TMEM buffer g&1
    uint64_t* p_full = bars + 11;       // [1]  P_g is in smem and S buffer g&1 is free (8 warp arrivals)
    uint64_t* o_full = bars + 12;       // [1]  P_g V_g is in TMEM, the P buffer is free
    uint64_t* q_empty = bars + 13;      // [1]  every S MMA of the tile has read Q
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // key chunks per tile: packed self-attention tiles hold their keys in one chunk; cross-attention streams the context
    const int nchunks = (a.mode == 0 || (a.mode == 1 && !a.self_mask)) ? (a.nk + AT_BN - 1) / AT_BN : 1;
    const int tstep = gridDim.x;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int s = 0; s < AT_KV; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        mbar_init(&s_full[0], 1);
        mbar_init(&s_full[1], 1);
        mbar_init(p_full, 8);
        mbar_init(o_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<AT_TMEM_COLS>(tmem_ptr_smem);
    if (a.mode == 2) {
        // the packed box brings G * F < 128 rows: the rows behind it are never written by TMA.  Their keys are masked (P = 0
        // exactly), but 0 x NaN would still poison P V, so the tail rows of every V stage are zeroed once.
        const int first = a.G * a.nq * 128;                   // bytes
        for (int i = first + (int)threadIdx.x * 16; i < AT_TILE; i += AT_THREADS * 16)
            for (int st = 0; st < AT_KV; ++st) *reinterpret_cast<uint4*>(smem + AT_OFF_V + st * AT_TILE + i) = make_uint4(0, 0, 0, 0);
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_launch_dependents();                                  // prologue done: let the next kernel start its own
    pdl_wait();                                               // Q/K/V come from the previous kernel

    if (warp == 0) {
        if (lane == 0) {
            int g = 0, tc = 0;                                // chunk / tile counters of this CTA
            const uint32_t box_bytes = a.mode == 2 ? (uint32_t)(a.G * a.nq) * 128u : (uint32_t)AT_TILE;   // rows a box brings x 128 B
            for (int T = blockIdx.x; T < a.ntiles; T += tstep, ++tc) {
                const int qt = T % a.nqt, h = (T / a.nqt) % a.heads, bo = T / (a.nqt * a.heads);
                mbar_wait(q_empty, (tc & 1) ^ 1);
                mbar_arrive_expect_tx(q_full, box_bytes);
                if (a.mode == 0) {
                    const int kbo = bo / a.kv_group;
                    tma_load_2d(smem + AT_OFF_Q, &tmQ, q_full, h * AT_D, bo * a.nq + qt * AT_BM);
                    for (int j = 0; j < nchunks; ++j, ++g) {
                        const int st = g % AT_KV;
                        mbar_wait(&kv_empty[st], ((g / AT_KV) & 1) ^ 1);
                        mbar_arrive_expect_tx(&kv_full[st], 2 * AT_TILE);
                        tma_load_2d(smem + AT_OFF_K + st * AT_TILE, &tmK, &kv_full[st], h * AT_D, kbo * a.nk + j * AT_BN);
                        tma_load_2d(smem + AT_OFF_V + st * AT_TILE, &tmV, &kv_full[st], h * AT_D, kbo * a.nk + j * AT_BN);
                    }
                } else {
                    // packed.  mode 1: qt = 128-row block of the flattened row space (bo == 0); mode 2: qt = group of G pixels,
                    // bo = sample
                    int krow = qt * AT_BM;                                            // self-attention: the same rows
                    if (a.mode == 1 && !a.self_mask) krow = ((qt * AT_BM) >> a.seq_shift) / a.kv_group * a.nk;   // the frames' shared context
                    if (a.mode == 1) tma_load_2d(smem + AT_OFF_Q, &tmQ, q_full, h * AT_D, qt * AT_BM);
                    else tma_load_4d(smem + AT_OFF_Q, &tmQ, q_full, h * AT_D, qt * a.G, 0, bo);
                    for (int j = 0; j < nchunks; ++j, ++g) {
                        const int st = g % AT_KV;
                        mbar_wait(&kv_empty[st], ((g / AT_KV) & 1) ^ 1);
                        mbar_arrive_expect_tx(&kv_full[st], 2 * box_bytes);
                        if (a.mode == 1) {
                            tma_load_2d(smem + AT_OFF_K + st * AT_TILE, &tmK, &kv_full[st], h * AT_D, krow + j * AT_BN);
                            tma_load_2d(smem + AT_OFF_V + st * AT_TILE, &tmV, &kv_full[st], h * AT_D, krow + j * AT_BN);
                        } else {
                            tma_load_4d(smem + AT_OFF_K + st * AT_TILE, &tmK, &kv_full[st], h * AT_D, qt * a.G, 0, bo);
                            tma_load_4d(smem + AT_OFF_V + st * AT_TILE, &tmV, &kv_full[st], h * AT_D, qt * a.G, 0, bo);
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = umma_idesc_f16_f32(AT_BM, AT_BN);                  // K-major A and B
            constexpr uint32_t idesc_o = umma_idesc_f16_f32(AT_BM, AT_D) | (1u << 16);     // B (= V) is MN-major
            const uint64_t qdesc = umma_desc_sw128_kmajor(smem_u32(smem + AT_OFF_Q));
            const uint64_t pdesc0 = umma_desc_sw128_kmajor(smem_u32(smem + AT_OFF_P));
            const uint64_t pdesc1 = umma_desc_sw128_kmajor(smem_u32(smem + AT_OFF_P + AT_TILE));
            // score cursor: runs up to two chunks ahead of the P V cursor, across tile boundaries
            int sT = blockIdx.x, sj = 0, sg = 0, stc = 0;
            auto issue_next_s = [&]() {                       // S_sg -> TMEM buffer sg&1
                if (sT >= a.ntiles) return;
                if (sj == 0) mbar_wait(q_full, stc & 1);
                const int st = sg % AT_KV;
                mbar_wait(&kv_full[st], (sg / AT_KV) & 1);
                tc_fence_after();
                const uint64_t kdesc = umma_desc_sw128_kmajor(smem_u32(smem + AT_OFF_K + st * AT_TILE));
#pragma unroll
                for (int k = 0; k < AT_D / 16; ++k)
                    umma_f16_ss(tmem_base + (sg & 1) * AT_BN, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
                umma_commit(&s_full[sg & 1]);
                ++sg;
                if (++sj == nchunks) { umma_commit(q_empty); sj = 0; sT += tstep; ++stc; }
            };
            issue_next_s();
            issue_next_s();
            int g = 0;
            for (int T = blockIdx.x; T < a.ntiles; T += tstep) {
                for (int j = 0; j < nchunks; ++j, ++g) {
                    const int st = g % AT_KV;
                    mbar_wait(p_full, g & 1);                 // P_g in smem; the softmax warps are done with S buffer g&1
                    tc_fence_after();
                    const uint64_t vdesc = umma_desc_sw128_mnmajor(smem_u32(smem + AT_OFF_V + st * AT_TILE), AT_BN);
#pragma unroll
                    for (int k = 0; k < AT_BN / 16; ++k) {
                        const uint64_t pd = (k < 4 ? pdesc0 : pdesc1) + 2 * (k & 3);
                        umma_f16_ss(tmem_base + AT_OCOL, pd, vdesc + (2048 >> 4) * k, idesc_o, k > 0 ? 1u : 0u);
                    }
                    umma_commit(o_full);
                    umma_commit(&kv_empty[st]);
                    issue_next_s();                           // refill the score buffer softmax_g just released
                }
            }
        }
        __syncwarp();
    } else {
        // ---------------- softmax / output: a query row = one thread of each of the two warps of its lane quarter ----------------
        const int q = warp & 3;
        const int hh = (warp - 2) >> 2;                       // key half [64 hh, +64) and output columns [32 hh, +32)
        const int r = q * 32 + lane;
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const int swz = r & 7;
        uint8_t* prow = smem + AT_OFF_P + hh * AT_TILE + r * 128;      // my 64-key atom, my row
        float* xmax = reinterpret_cast<float*>(smem + AT_OFF_X);       // [2][2][128]
        float* xsum = xmax + 4 * 128;                                  // [2][128]
        const float sl2 = a.scale_log2;
        auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); };
        int g = 0;
        for (int T = blockIdx.x; T < a.ntiles; T += tstep) {
            const int qt = T % a.nqt, h = (T / a.nqt) % a.heads, bo = T / (a.nqt * a.heads);
            float m_run = -INFINITY, l_run = 0.f;
            float o[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = 0.f;
            auto add_o = [&]() {                              // o += my 32 columns of O_g from TMEM
                uint32_t v[32];
                tmem_ld_32x32b_x16(trow + AT_OCOL + hh * 32, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
                tmem_ld_32x32b_x16(trow + AT_OCOL + hh * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] += __uint_as_float(v[i]);
            };
            for (int j = 0; j < nchunks; ++j, ++g) {
                const int kbase = j * AT_BN + hh * 64;
                // warp-uniform: only the last chunk can be ragged; the packed modes always take the masked path
                const bool full = a.mode == 0 ? kbase + 64 <= a.nk : (a.mode == 1 && !a.self_mask && kbase + 64 <= a.nk);
                mbar_wait(&s_full[g & 1], (g >> 1) & 1);
                tc_fence_after();
                uint32_t sv[64];
#pragma unroll
                for (int c = 0; c < 64; c += 16)
                    tmem_ld_32x32b_x16(trow + (g & 1) * AT_BN + hh * 64 + c, *reinterpret_cast<uint32_t(*)[16]>(&sv[c]));
                tmem_ld_wait();
                float mx = -INFINITY;
                if (full) {
#pragma unroll
                    for (int i = 0; i < 64; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
                } else if (a.mode == 2) {
                    // tile row = f * G + pixel: key jj belongs to my sequence iff jj % G == r % G (and it is a loaded row)
                    const int G = a.G, lim = a.G * a.nq;
                    int jm = kbase % G;
                    const int rm = r % G;
#pragma unroll
                    for (int i = 0; i < 64; ++i) {
                        if (jm != rm || kbase + i >= lim) sv[i] = 0xff800000u;
                        jm = (jm + 1 == G) ? 0 : jm + 1;
                        mx = fmaxf(mx, __uint_as_float(sv[i]));
                    }
                } else if (a.mode == 1 && a.self_mask) {
                    // packed contiguous sequences of 2^seq_shift tokens: block-diagonal
                    const int rb = r >> a.seq_shift;
#pragma unroll
                    for (int i = 0; i < 64; ++i) {
                        if (((kbase + i) >> a.seq_shift) != rb) sv[i] = 0xff800000u;
                        mx = fmaxf(mx, __uint_as_float(sv[i]));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 64; ++i) {
                        if (kbase + i >= a.nk) sv[i] = 0xff800000u;             // -inf: exp2 gives exactly 0
                        mx = fmaxf(mx, __uint_as_float(sv[i]));
                    }
                }
                float* xm = xmax + (g & 1) * 256;
                xm[hh * 128 + r] = mx;
                pair_sync();
                const float m_new = fmaxf(m_run, fmaxf(mx, xm[(hh ^ 1) * 128 + r]));   // finite: chunk 0 holds key 0
                const float corr = ex2_approx_f((m_run - m_new) * sl2);
                m_run = m_new;
                const float msc = m_new * sl2;
                // Exponentials first, into registers: the MUFU phase overlaps P_{g-1} V_{g-1} on the tensor core.  Only
                // then wait for that MMA (it frees the P buffer and holds the O tile to fold into the running output).
                uint32_t pk[32];
                float lsum = 0.f;
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {              // 8 x 16-byte chunks = my 64 keys
                    float p[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        p[i] = ex2_approx_f(fmaf(__uint_as_float(sv[c8 * 8 + i]), sl2, -msc));
                    lsum += ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
#pragma unroll
                    for (int i = 0; i < 4; ++i) pk[c8 * 4 + i] = pack_half2(p[2 * i], p[2 * i + 1]);
                }
                l_run = fmaf(l_run, corr, lsum);
                if (j > 0) {
                    mbar_wait(o_full, (g - 1) & 1);           // P_{g-1} V_{g-1} landed; the P buffer is free again
                    tc_fence_after();
                    add_o();
                }
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8)
                    *reinterpret_cast<uint4*>(prow + ((c8 ^ swz) << 4)) =
                        make_uint4(pk[c8 * 4], pk[c8 * 4 + 1], pk[c8 * 4 + 2], pk[c8 * 4 + 3]);
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] *= corr;
                fence_proxy_async();                          // generic-proxy smem writes -> visible to the MMA
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full);
            }
            mbar_wait(o_full, (g - 1) & 1);
            tc_fence_after();
            add_o();
            xsum[hh * 128 + r] = l_run;
            pair_sync();
            const float inv = 1.f / (l_run + xsum[(hh ^ 1) * 128 + r]);
            const int qrow = qt * AT_BM + r;
            bool st_ok;
            long long o_off;
            if (a.mode == 0) {
                st_ok = qrow < a.nq;
                o_off = (long long)bo * a.o_bs + (long long)qrow * a.o_rs;
            } else if (a.mode == 1) {                             // flattened row space, sequences contiguous
                st_ok = qrow < a.rows_total;
                o_off = (long long)qrow * a.o_rs;
            } else {                                              // row r = frame f, pixel qt * G + p
                const int f = r / a.G, pix = qt * a.G + (r - f * a.G);
                st_ok = f < a.nq && pix < a.HW;
                o_off = (long long)bo * a.o_bs + (long long)pix * a.o_bs_inner + (long long)f * a.o_rs;
            }
            if (st_ok) {
                __half* dst = a.o + o_off + h * AT_D + hh * 32;
#pragma unroll
                for (int c = 0; c < 32; c += 16) {
                    uint32_t w[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) w[i] = pack_half2(o[c + 2 * i] * inv, o[c + 2 * i + 1] * inv);
                    stg256(dst + c, w);
                }
            }
            pair_sync();                                      // xsum is reused by the next tile
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<AT_TMEM_COLS>(tmem_base);
}

static bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; }

// Returns VMV_OK if launched, VMV_ERR_UNSUPPORTED if the problem does not fit this kernel (caller falls back to the
// generic strided kernel), another code on error.
int attention_tc_try(const vmv_attn_params* p, cudaStream_t st) {
    // The packed forms are correct for every short-sequence attention of the UNet but measured SLOWER than the mma.sync
    // kernel on them (temporal 24 x 24 at the 32x32 level, CFG batch: 53.5 vs 27 us; 64-token spatial: 13.9 vs ~9 us --
    // profiles/r2_attention_packed.md): one key chunk per tile leaves the per-tile latency chain (TMA -> S -> softmax -> P ->
    // PV -> O) unamortised, while the mma.sync kernel already streams Q/K/V/O at 3.4-4.6 TB/s.  So `impl = 0` (auto) takes
    // them only with VMV_ATTN_TC_PACKED=1; `impl = 2` always does.
    static int packed_auto = -1;
    if (packed_auto < 0) { const char* e = getenv("VMV_ATTN_TC_PACKED"); packed_auto = (e && e[0] == '1') ? 1 : 0; }
    if (p->impl == 0 && !packed_auto && !(p->inner == 1 && p->nq >= 128)) return VMV_ERR_UNSUPPORTED;
    // ---- which form of the kernel takes this problem?
    int mode = -1, G = 0, seq_shift = 0, self_mask = 0;
    if (p->inner == 1 && p->nq >= 128) {
        mode = 0;
    } else if (p->inner == 1 && p->nq >= 8 && (p->nq & (p->nq - 1)) == 0) {
        // short contiguous sequences: 128 / nq of them per tile
        while ((1 << seq_shift) < p->nq) ++seq_shift;
        G = 128 / p->nq;
        if (p->kv_group == 1 && p->nk == p->nq) { mode = 1; self_mask = 1; }
        else if (p->kv_group > 1 && p->kv_group % G == 0) { mode = 1; self_mask = 0; }
        if (mode == 1 && p->o_bs_outer != (int64_t)p->nq * p->o_rs) mode = -1;       // the output rows must be contiguous too
    } else if (p->inner > 1 && p->kv_group == 1 && p->nq == p->nk && p->nq >= 2 && p->nq <= 128) {
        mode = 2;                                                 // strided short sequences (temporal attention)
        G = 128 / p->nq;
    }
    if (mode < 0) return VMV_ERR_UNSUPPORTED;
    if (mode != 2 && (p->q_bs_outer != (int64_t)p->nq * p->q_rs || p->k_bs_outer != (int64_t)p->nk * p->k_rs ||
                      p->v_bs_outer != (int64_t)p->nk * p->v_rs))
        return VMV_ERR_UNSUPPORTED;                               // batches must be contiguous row blocks
    if (p->outer % p->kv_group != 0) return VMV_ERR_UNSUPPORTED;
    if (!aligned32(p->o) || p->o_rs % 16 != 0 || p->o_bs_outer % 16 != 0 || (mode == 2 && p->o_bs_inner % 16 != 0)) return VMV_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(p->q) | reinterpret_cast<uintptr_t>(p->k) | reinterpret_cast<uintptr_t>(p->v)) & 15)
        return VMV_ERR_UNSUPPORTED;
    CUtensorMap tq, tk, tv;
    int rc;
    if (mode != 2) {
        const unsigned box[2] = {AT_D, 128};
        auto mk = [&](CUtensorMap* m, const void* base, long long rows, long long rs) {
            const unsigned long long dims[2] = {(unsigned long long)p->heads * AT_D, (unsigned long long)rows};
            const unsigned long long strides[1] = {(unsigned long long)rs * 2};
            return make_map_generic(m, base, 2, dims, strides, box, 128);
        };
        if ((rc = mk(&tq, p->q, (long long)p->outer * p->nq, p->q_rs)) != VMV_OK) return rc;
        if ((rc = mk(&tk, p->k, (long long)(p->outer / p->kv_group) * p->nk, p->k_rs)) != VMV_OK) return rc;
        if ((rc = mk(&tv, p->v, (long long)(p->outer / p->kv_group) * p->nk, p->v_rs)) != VMV_OK) return rc;
    } else {
        // (d, pixel, frame, sample): one box = 64 d x G pixels x all F frames of one sample; rows land as (frame, pixel)
        const unsigned box[4] = {AT_D, (unsigned)G, (unsigned)p->nq, 1};
        auto mk4 = [&](CUtensorMap* m, const void* base, long long bs_o, long long bs_i, long long rs) {
            if ((bs_o * 2) % 16 || (bs_i * 2) % 16 || (rs * 2) % 16) return (int)VMV_ERR_UNSUPPORTED;
            const unsigned long long dims[4] = {(unsigned long long)p->heads * AT_D, (unsigned long long)p->inner,
                                                (unsigned long long)p->nq, (unsigned long long)p->outer};
            const unsigned long long strides[3] = {(unsigned long long)bs_i * 2, (unsigned long long)rs * 2, (unsigned long long)bs_o * 2};
            return make_map_generic(m, base, 4, dims, strides, box, 128);
        };
        if ((rc = mk4(&tq, p->q, p->q_bs_outer, p->q_bs_inner, p->q_rs)) != VMV_OK) return rc;
        if ((rc = mk4(&tk, p->k, p->k_bs_outer, p->k_bs_inner, p->k_rs)) != VMV_OK) return rc;
        if ((rc = mk4(&tv, p->v, p->v_bs_outer, p->v_bs_inner, p->v_rs)) != VMV_OK) return rc;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
        if (e != cudaSuccess) { set_error("attention_tc: smem attribute: %s", cudaGetErrorString(e)); return VMV_ERR_CUDA; }
        attr_set = true;
    }
    AttTcArgs a;
    memset(&a, 0, sizeof(a));
    a.nq = p->nq; a.nk = p->nk; a.kv_group = p->kv_group;
    a.scale_log2 = p->scale * 1.4426950408889634f;
    a.o = static_cast<__half*>(p->o);
    a.o_bs = p->o_bs_outer; a.o_rs = p->o_rs; a.o_bs_inner = p->o_bs_inner;
    a.heads = p->heads;
    a.mode = mode; a.G = G; a.seq_shift = seq_shift; a.self_mask = self_mask;
    a.rows_total = (int)((long long)p->outer * p->nq);
    a.HW = p->inner;
    long long ntiles;
    if (mode == 0) {
        a.nqt = (p->nq + AT_BM - 1) / AT_BM;
        ntiles = (long long)a.nqt * p->heads * p->outer;
    } else if (mode == 1) {
        a.nqt = (int)(((long long)p->outer * p->nq + AT_BM - 1) / AT_BM);       // 128-row blocks of the flattened row space
        ntiles = (long long)a.nqt * p->heads;
    } else {
        a.nqt = (p->inner + G - 1) / G;                                          // groups of G pixels
        ntiles = (long long)a.nqt * p->heads * p->outer;
    }
    if (ntiles > 0x7fffffffLL) return VMV_ERR_UNSUPPORTED;
    a.ntiles = (int)ntiles;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int grid = a.ntiles < num_sms ? a.ntiles : num_sms;   // persistent: one CTA per SM (176 KiB smem, 512 TMEM columns)
    launch_kernel(attention_tc_kernel, dim3(grid), dim3(AT_THREADS), AT_SMEM, st, tq, tk, tv, a);
    count_launch();
    VMV_CUDA_LAUNCH_CHECK("vmv_attention (tcgen05)");
    return VMV_OK;
}

}  // namespace vmv
