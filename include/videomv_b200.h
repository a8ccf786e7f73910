/*
 * videomv_b200 -- C ABI of the B200-native (sm_100a) kernels behind the VideoMV video-UNet forward.
 *
 * The reference (alibaba/VideoMV) is 100% Python: its "FFI" for this path is the set of torch library
 * calls made from tools/modules/unet/{unet_t2v.py,unet_i2vgen.py,util.py}.  Each entry point below
 * names the reference call sites it replaces.  Signatures are plain C: raw device pointers, sizes,
 * strides (in ELEMENTS unless stated), and a CUDA stream handle passed as void*.  No torch types.
 *
 * Conventions
 *   - activations are channels-last fp16:  x[(b*F + f), h, w, c]  ==  row-major [M = B*F*H*W, C]
 *   - accumulation, normalisation statistics and softmax are fp32 (stats are reduced in fp64)
 *   - every function returns 0 on success; otherwise vmv_last_error() describes the failure.
 *     Nothing falls back to a CPU path.
 *   - all launches go to `stream` and are CUDA-graph capturable.
 */
#ifndef VIDEOMV_B200_H_
#define VIDEOMV_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* vmv_last_error(void);
/* ABI version of this header; bump on any struct change. */
int vmv_abi_version(void);
/* Number of kernels launched through this library since process start (bench.py `gpu_launches`). */
long long vmv_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core contraction:  D[M,N] = epilogue( A (*) W^T )      (tcgen05.mma, TMA-staged, TMEM acc)
 *
 * mode VMV_GEMM_LINEAR   A = [A1 | A2]  ([M,K1] and [M,K2] fp16, K-concatenated, K2 may be 0)
 *      replaces nn.Linear (util.py:223-227,337,351,546,573,667; unet_t2v.py:141-151), nn.Conv1d k=1
 *      (util.py:1016,1032), nn.Conv2d 1x1 skip (util.py:688) incl. the torch.cat of unet_t2v.py:361.
 * mode VMV_GEMM_CONV3X3  A = implicit im2col of x[BF,H,W,Cin], 3x3, stride 1, zero pad 1
 *      (9 shifted 4-D TMA box loads per K block; halo by TMA out-of-bounds zero fill)
 *      replaces nn.Conv2d 3x3 (util.py:651,677,595; unet_t2v.py:264).
 * mode VMV_GEMM_TCONV3   A = implicit 3-tap unfold along frames of x[B,F,HW,Cin], zero pad 1 on F
 *      replaces nn.Conv3d (3,1,1) (util.py:1360-1375) without the NCHW<->NCFHW rearranges.
 *
 * mode VMV_GEMM_CONV3X3_S2  3x3 conv, STRIDE 2, zero pad 1 of x[BF,H,W,Cin] -> [BF,H/2,W/2,N]: the stride-2 windows are
 *      read by TMA boxes with element strides {1,2,2,1} (no patch gather); replaces Downsample.op (util.py:749).
 *      B,F,H,Wd describe the INPUT; M = BF*(H/2)*(Wd/2).
 * mode VMV_GEMM_UPCONV3X3  nearest-x2 upsample followed by a 3x3 conv (Upsample, util.py:604-606) WITHOUT the 4x tensor:
 *      output pixel (2y+py, 2x+px) is a 2x2 conv of the input around (y, x) with the 3x3 taps that fall on the same input
 *      pixel pre-summed (packing.pack_upconv3x3) -- 4 phases x 4 taps x Cin, 2.25x fewer FLOPs than the 9-tap conv on the
 *      upsampled image.  B,F,H,Wd describe the INPUT; M = BF*4*H*Wd; W is [4*N, 4*Cin] (phase-major rows, K = (ty,tx,c)).
 *
 * W is fp16 [N, Ktot] row-major (K contiguous): Ktot = K1+K2 (linear), 9*Cin ordered (ky,kx,c),
 * or 3*Cin ordered (kt,c).  See videomv_b200/packing.py for the repack from reference layouts.
 *
 * epilogue, in fp32, in this order:  + bias[n]  + rowbias[row / rows_per_group, n]
 *                                    act (none | SiLU | GEGLU)  + residual[row, n]   -> fp16
 * GEGLU (util.py:543-550): W rows are packed per N tile as [BN/2 value rows | BN/2 gate rows];
 * output has N/2 columns: value * gelu_erf(gate).
 * ---------------------------------------------------------------------------------------------- */
enum { VMV_GEMM_LINEAR = 0, VMV_GEMM_CONV3X3 = 1, VMV_GEMM_TCONV3 = 2, VMV_GEMM_CONV3X3_S2 = 3, VMV_GEMM_UPCONV3X3 = 4 };
enum { VMV_ACT_NONE = 0, VMV_ACT_SILU = 1, VMV_ACT_GEGLU = 2 };

#ifndef VMV_PEER_MAX_RANKS
#define VMV_PEER_MAX_RANKS 8
#endif
/* Multi-GPU sharding: vmv_peer_exchange (below) fused into this GEMM's epilogue.  The output rows are NOT written to D but
 * straight into the tensors of the OTHER sharding layout in every rank's arena -- compute and collective as one kernel, tile by
 * tile: the destination (rank q, row) of an output row is a pure function of its index
 *   direction 0 (frame shard -> pixel shard): row (b, f, pixel)  -> rank q = pixel / HWl, row ((b*world + rank)*Fl + f)*HWl + pixel % HWl
 *   direction 1 (pixel shard -> frame shard): row (b, fg, pl)    -> rank q = fg / Fl,      row ((b*Fl + fg % Fl)*world + rank)*HWl + pl
 * -- and the kernel ends with the epoch-flag rendezvous of vmv_peer_exchange (same control-line layout), so when it completes
 * this rank's whole destination tensor is in its memory.  dst[q] have leading dimension ldd.  CTA-pair kernel, register
 * epilogue, no split-K (else VMV_ERR_UNSUPPORTED). */
typedef struct vmv_gemm_scatter {
    int32_t world, rank, direction, nowait;
    int32_t B, Fl, HWl, pad_;
    void* dst[VMV_PEER_MAX_RANKS];
    void* flags[VMV_PEER_MAX_RANKS];
    void* epoch; void* done;
} vmv_gemm_scatter;

typedef struct vmv_gemm_params {
    int32_t mode;
    int32_t M, N;               /* output rows / GEMM columns (GEGLU writes N/2 columns) */
    int32_t K1, K2;             /* linear: K of A1, A2.  conv modes: K1 = Cin, K2 = 0 */
    const void* A1; int64_t lda1;   /* fp16; lda = row (pixel) stride in elements */
    const void* A2; int64_t lda2;
    const void* W;  int64_t ldw;    /* fp16 [N, Ktot] */
    void* D;        int64_t ldd;    /* fp16 [M, N or N/2] */
    /* geometry for the conv modes */
    int32_t B, F, H, Wd;        /* CONV3X3: images = B*F of H x Wd.  TCONV3: B samples, F frames, H*Wd pixels */
    /* epilogue */
    const float* bias;          /* fp32 [N] or NULL */
    const void* rowbias; int64_t ld_rowbias; int32_t rows_per_group;  /* fp16 [M/rows_per_group, N] or NULL */
    const void* residual; int64_t ldr;                                /* fp16 [M, N_out] or NULL */
    int32_t act;
    /* LayerNorm folded into this GEMM (nn.LayerNorm util.py:528-530 feeding nn.Linear): A holds the RAW rows, W holds
     * W*gamma, bias holds bias + W@beta, and the epilogue computes  rstd[m]*(acc - mean[m]*ln_colsum[n]) + bias[n]
     * with ln_stats[m] = {mean, rstd} (vmv_layernorm_stats) and ln_colsum[n] = sum_k W'[n,k].  NULL = off. */
    const void* ln_stats; const float* ln_colsum;
    /* ln_stats_src_n != 0: ln_stats is the `rowstats_out` of the upstream vmv_gemm that produced A (N = ln_stats_src_n,
     * block_n = ln_stats_src_bn = vmv_gemm_block_n of that call): per row vmv_gemm_epilogue_split()*ceil(N/block_n) partial {mean_k, M2_k} slots,
     * merged by the epilogue in slot order (parallel-variance merge) and finished with ln_eps.  0 = ln_stats is {mean, rstd}. */
    int32_t ln_stats_src_n; int32_t ln_stats_src_bn; float ln_eps;
    /* rowstats_out != NULL: fp32 [M][vmv_gemm_epilogue_split()*ceil(N/block_n)][2]; the epilogue writes, per row and per (N tile,
     * epilogue warp of the row's TMEM lane quarter), the partial LayerNorm statistics {mean_k, M2_k = sum (d - mean_k)^2} of the final (pre-rounding) output
     * values it holds -- the LayerNorm statistics of the tensor this GEMM produces, for the next GEMM's folded LayerNorm.
     * One writer per slot (no atomics, no initialisation needed, bit-reproducible).  CTA-pair kernel without split-K /
     * GEGLU only (else VMV_ERR_UNSUPPORTED). */
    void* rowstats_out;
    const vmv_gemm_scatter* scatter;   /* NULL = plain output to D */
    /* tuning (0 = auto) */
    int32_t block_n;            /* 64 (variant 1 only), 128, 160 or 256 */
    int32_t stages;             /* smem pipeline depth */
    int32_t split_k;            /* >1: K split across CTAs, fp32 partials in workspace, reduced by a 2nd kernel */
    int32_t variant;            /* 0 auto | 1 one tile per CTA (cta_group::1) | 2 persistent CTA pairs (cta_group::2) */
    int32_t w_static;           /* != 0: W is not written by any earlier work of this stream that may still be in flight
                                 * (model weights), so the kernel may fetch its first W tiles BEFORE it waits for the
                                 * previous kernel (programmatic dependent launch); 0 = W is ordered like A */
    void* workspace; int64_t workspace_bytes;
} vmv_gemm_params;

int vmv_gemm(const vmv_gemm_params* p, void* stream);
/* the N-tile width vmv_gemm will use for p (sizes rowstats_out; becomes the consumer's ln_stats_src_bn) */
int vmv_gemm_block_n(const vmv_gemm_params* p);
/* epilogue warps per TMEM lane quarter of the CTA-pair kernel = statistic slots per (row, N tile) of rowstats_out */
int vmv_gemm_epilogue_split(void);
/* bytes of workspace vmv_gemm needs for p (0 when split_k <= 1) */
int64_t vmv_gemm_workspace_bytes(const vmv_gemm_params* p);

/* ------------------------------------------------------------------------------------------------
 * GroupNorm(32 groups) over channels-last rows; replaces nn.GroupNorm at util.py:329,649,673,1014,
 * 1358-1372 and unet_t2v.py:262, plus the following nn.SiLU, plus the torch.cat feeding it.
 *
 * The logical input is [x1 | x2] along channels (C2 may be 0).  Rows are split into `nbatch`
 * consecutive chunks of `rows_per_batch` rows; statistics are per (chunk, group): a chunk is one frame
 * (4-D GroupNorm) or one whole sample (5-D GroupNorm: statistics span frames, util.py:1358,1014).
 *   vmv_groupnorm_stats: stats[nbatch*32*2] (fp64 sum, sum of squares), written by the call.
 *   vmv_groupnorm_apply: out[rows, C1+C2] fp16 = (x-mean)*rstd*gamma+beta, optional SiLU.  `stat_rows` (0 = rows_per_batch)
 *                        is the number of rows the statistics cover: larger than rows_per_batch when the caller summed the
 *                        partial statistics of several row shards (multi-GPU pixel sharding) between the two calls.
 * All reductions are fixed-order (per-CTA slots summed in CTA order; no floating-point atomics): results are bit-identical
 * from run to run.  Every statistics-producing call takes
 *   barriers: nbatch x 2 uint32 {arrival count, generation}, ZERO when first used and never touched by the caller again
 *             (self-resetting; may be reused by any later call of the same stream);
 *   scratch:  vmv_groupnorm_scratch_bytes(C1+C2, rows_per_batch, nbatch) bytes, uninitialised (per-CTA partial sums).
 * ---------------------------------------------------------------------------------------------- */
int64_t vmv_groupnorm_scratch_bytes(int32_t C, int64_t rows_per_batch, int32_t nbatch);
int vmv_groupnorm_stats(const void* x1, int64_t ldx1, int32_t C1, const void* x2, int64_t ldx2, int32_t C2,
                        int64_t rows_per_batch, int32_t nbatch, double* stats, void* barriers, void* scratch, void* stream);
int vmv_groupnorm_apply(const void* x1, int64_t ldx1, int32_t C1, const void* x2, int64_t ldx2, int32_t C2,
                        int64_t rows_per_batch, int32_t nbatch, const double* stats, int64_t stat_rows,
                        const float* gamma, const float* beta, float eps, int32_t silu,
                        void* out, int64_t ldo, void* stream);
/* Single-launch form of the two calls above (single-GPU path: no reduction between statistics and apply): the rows are
 * staged in shared memory once (or re-read from L2 when they do not fit), statistics, an in-kernel arrival barrier per
 * chunk, apply.  Returns VMV_ERR_UNSUPPORTED when the grid cannot be made co-resident (never spins then). */
/* 1 when vmv_groupnorm_fused takes the smem-resident single-pass kernel for this shape on the current device (the only form
 * vmv_groupnorm_fused_peer supports), 0 when it re-reads the rows from L2. */
int vmv_groupnorm_fused_fits_smem(int32_t C, int64_t rows_per_batch, int32_t nbatch);
int vmv_groupnorm_fused(const void* x1, int64_t ldx1, int32_t C1, const void* x2, int64_t ldx2, int32_t C2,
                        int64_t rows_per_batch, int32_t nbatch, void* barriers, void* scratch, const float* gamma,
                        const float* beta, float eps, int32_t silu, void* out, int64_t ldo, void* stream);

/* vmv_groupnorm_fused for a chunk whose rows are spread over `world` GPUs (5-D GroupNorm in the pixel-sharded layout of
 * multi-GPU frame sharding, see vmv_peer_exchange below): the cross-GPU sum of the statistics happens INSIDE the kernel
 * over NVLink peer memory (the first CTA of a chunk publishes this rank's partial sums into every rank's slot + an epoch
 * flag; all CTAs wait for the `world` epochs and sum the slots in rank order).  slots[q]: [world][nbatch][64] doubles;
 * control words: one 64-byte line per chunk at the same arena offset of every rank q, laid out like every other peer
 * op's line (uint32 flags[8] at +0, epoch at +32, done at +36; zero at start): flags[q] = line 0 of rank q, epoch = this
 * rank's line 0 + 32; stat_rows = rows of a chunk over all ranks.  VMV_ERR_UNSUPPORTED when the tensor does not fit the
 * smem-resident kernel (callers then use stats + vmv_peer_allreduce_f64 + apply). */
#ifndef VMV_PEER_MAX_RANKS
#define VMV_PEER_MAX_RANKS 8
#endif
typedef struct vmv_gn_peer {
    int32_t world, rank;
    double* slots[VMV_PEER_MAX_RANKS];
    void* flags[VMV_PEER_MAX_RANKS];
    void* epoch;
    int64_t stat_rows;
} vmv_gn_peer;
int vmv_groupnorm_fused_peer(const void* x1, int64_t ldx1, int32_t C1, const void* x2, int64_t ldx2, int32_t C2,
                             int64_t rows_per_batch, int32_t nbatch, void* barriers, void* scratch, const float* gamma,
                             const float* beta, float eps, int32_t silu, void* out, int64_t ldo, const vmv_gn_peer* peer,
                             void* stream);

/* Per-row LayerNorm statistics only: stats[m] = {mean, 1/sqrt(var+eps)} fp32 (for the folded form in vmv_gemm). */
int vmv_layernorm_stats(const void* x, int64_t ldx, int64_t M, int32_t C, float eps, void* stats, void* stream);
/* LayerNorm over the last dim (eps 1e-5): nn.LayerNorm util.py:528-530.  x,out fp16 [M,C]. */
int vmv_layernorm(const void* x, int64_t ldx, int64_t M, int32_t C, const float* gamma, const float* beta,
                  float eps, void* out, int64_t ldo, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused softmax(q k^T * scale) v, head_dim 64, fp16 in/out, fp32 softmax (online, flash style).
 * Replaces xformers.ops.memory_efficient_attention + the head split/merge reshapes of util.py:237-268.
 * Batch index = bo*inner + bi (bo < outer, bi < inner); element offset of token n, head h:
 *     base + bo*bs_outer + bi*bs_inner + n*row_stride + h*64
 * K/V batch = (q batch) / kv_group  (kv_group = frames for text cross-attention, else 1).
 *   spatial self  (util.py:536 attn1 in SpatialTransformer): outer=B*F frames, inner=1, n over H*W
 *   text cross    (attn2, context [B,77|145,1024] projected once per sample)
 *   temporal self (attn1+attn2 in TemporalTransformer): outer=B, inner=H*W pixels, n over F frames
 * ---------------------------------------------------------------------------------------------- */
typedef struct vmv_attn_params {
    const void* q; const void* k; const void* v; void* o;     /* fp16 */
    int32_t outer, inner, heads, nq, nk;
    int64_t q_bs_outer, q_bs_inner, q_rs;
    int64_t k_bs_outer, k_bs_inner, k_rs;
    int64_t v_bs_outer, v_bs_inner, v_rs;
    int64_t o_bs_outer, o_bs_inner, o_rs;
    int32_t kv_group;          /* kv outer index = bo / kv_group */
    float scale;               /* head_dim^-0.5 */
    int32_t impl;              /* 0 auto | 1 strided mma.sync kernel (any layout) | 2 tcgen05/TMEM kernel: contiguous batches
                                  of >= 128 rows; short contiguous sequences (power-of-two length, packed 128/n per tile,
                                  block-diagonal mask); strided sequences of <= 128 rows (temporal attention: 128/F pixels
                                  packed per tile by a 4-D TMA box); VMV_ERR_UNSUPPORTED otherwise.  Auto picks 2 for the long
                                  sequences (and, with VMV_ATTN_TC_PACKED=1, for the packed forms, which measure slower
                                  than kernel 1 on B200), else 1. */
} vmv_attn_params;
int vmv_attention(const vmv_attn_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Small data-movement / CUDA-core kernels on the path
 * ---------------------------------------------------------------------------------------------- */
/* F.interpolate(scale 2, nearest) util.py:604: x[n,H,W,C] -> out[n,2H,2W,C] fp16 */
int vmv_upsample_nearest2x(const void* x, int32_t n, int32_t H, int32_t W, int32_t C, void* out, void* stream);
/* patch gather for the stride-2 3x3 Downsample conv (util.py:749): out[n*(H/2)*(W/2), 9*C] ordered (ky,kx,c) */
int vmv_im2col_3x3_s2(const void* x, int32_t n, int32_t H, int32_t W, int32_t C, void* out, void* stream);
/* input conv (unet_t2v.py:169): x fp32 NCFHW [B,C1,F,H,W] (+ optional second tensor [B,C2,F,H,W], I2V concat
 * unet_i2vgen.py:384), w fp32 [Cout, C1+C2, 3, 3], bias fp32 -> out fp16 [B*F,H,W,Cout] */
int vmv_conv3x3_in(const float* x1, int32_t C1, const float* x2, int32_t C2, int32_t B, int32_t F, int32_t H,
                   int32_t W, const float* w, const float* bias, int32_t Cout, void* out, void* stream);
/* head conv (unet_t2v.py:264-265,368): x fp16 [B*F,H,W,C] (already GN+SiLU), w fp32 [Cout,C,3,3]
 * -> out fp32 NCFHW [B,Cout,F,H,W] */
int vmv_conv3x3_out(const void* x, int32_t B, int32_t F, int32_t H, int32_t W, int32_t C, const float* w,
                    const float* bias, int32_t Cout, float* out, void* stream);
/* channels-last fp16 rows [B*F*H*W, ldx] (first Cout columns) -> fp32 NCFHW [B,Cout,F,H,W]: the final rearrange of
 * unet_t2v.py:368 when the head conv runs on the tensor cores (vmv_gemm CONV3X3 with Cout zero-padded to 16 columns) */
int vmv_rows_to_ncfhw(const void* x, int64_t ldx, int32_t B, int32_t F, int32_t H, int32_t W, int32_t Cout, float* out,
                      void* stream);
/* out[m, :] = softmax(scale * x[m, :]) over fp16 rows (fp32 maths), scale > 0: the softmax of the single-head, d = C
 * attention of the VAE decoder's middle block (autoencoder.py:419-441), whose QK^T and PV are plain vmv_gemm calls. */
int vmv_softmax_rows(const void* x, int64_t ldx, int64_t M, int32_t N, float scale, void* out, int64_t ldo, void* stream);
/* sinusoidal_embedding util.py:177-189: t int64 [B] -> out fp16 [B, dim] = [cos | sin] */
int vmv_sinusoidal_embedding(const int64_t* t, int32_t B, int32_t dim, void* out, void* stream);
/* e[b*F+f, :] = silu( t_emb[b,:] (+ t_emb2[b,:]) (+ cam_emb[b*F+f,:]) )   (unet_t2v.py:326-335 + the nn.SiLU
 * that opens every ResBlock.emb_layers, util.py:665).  fp16 in/out. */
int vmv_embed_combine_silu(const void* t_emb, const void* t_emb2, const void* cam_emb, int32_t B, int32_t F,
                           int32_t E, void* out, void* stream);
/* classifier-free guidance + DDIM update (diffusion_ddim.py:157-160,193-199,233-243), eta = 0, one launch:
 *   out    = u + s*(y-u)                         (y == u and s == 1 when there is no guidance)
 *   x0     = kx*xt - ko*out                      (eps-pred: kx=sqrt(1/ac), ko=sqrt(1/ac-1); v-pred: kx=sqrt(ac), ko=sqrt(1-ac))
 *   eps    = (sqrt(1/ac)*xt - x0) / sqrt(1/ac-1)
 *   x_prev = sqrt(ac_prev)*x0 + sqrt(1-ac_prev)*eps
 * all tensors fp32, n elements; coef7 (device) = {kx, ko, sqrt(1/ac), sqrt(1/ac-1), sqrt(ac_prev), sqrt(1-ac_prev), s} */
int vmv_cfg_ddim_step(const float* xt, const float* y_out, const float* u_out, const float* coef7, int64_t n,
                      float* x_prev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU frame sharding of ONE sample (BASELINE config 5; new design, no reference counterpart -- the reference's
 * multi-GPU mode is independent replicas, inference_text2video_entrance.py:79,152-170).  Around every temporal segment
 * (util.py:1043-1089 TemporalTransformer, :1381-1392 TemporalConvBlock_v2) the activation switches between
 *   layout A "frame shard"  rows (b, f_local, pixel)   [B*Fl*HW, C]      and
 *   layout B "pixel shard"  rows (b, f, pixel_local)   [B*F*HWl, C]      (F = Fl*world, HW = HWl*world).
 * vmv_peer_exchange does that switch as ONE kernel per rank over NVLink peer memory: remote 16 B stores of the slices
 * the other ranks need into THEIR output tensors, then an epoch flag barrier (st.release.sys / ld.acquire.sys); it
 * returns (stream-ordered) only when this rank's whole output is in its memory.  vmv_peer_allreduce_f64 sums the 5-D
 * GroupNorm partial statistics (util.py:1014,1358-1372) the same way, in rank order (bit-identical on all ranks).
 * dst[q] / flags[q] / slots[q] are the SAME arena offsets in every rank q's memory, mapped here through CUDA IPC
 * (vmv_ipc_export on the owner, vmv_ipc_import on the peers); [rank] entries are local pointers.  flags: `world`
 * uint32 per call site, zero at start; epoch: local uint32 per call site (zero at start); done: local uint32, zero.
 * All ranks must issue the same sequence of calls.  nowait != 0 skips the wait (single-process tests of the data
 * movement only).
 * ---------------------------------------------------------------------------------------------- */
#ifndef VMV_PEER_MAX_RANKS
#define VMV_PEER_MAX_RANKS 8
#endif
typedef struct vmv_peer_exchange_params {
    const void* src;                      /* local fp16 rows in the source layout, contiguous */
    void* dst[VMV_PEER_MAX_RANKS];        /* output tensor (destination layout) in every rank's arena */
    void* flags[VMV_PEER_MAX_RANKS];
    void* epoch; void* done;
    int32_t world, rank;
    int32_t direction;                    /* 0: A -> B (frames to pixels), 1: B -> A */
    int32_t B, Fl, HWl, C;                /* samples, frames per rank, pixels per rank, channels */
    int32_t nowait;
} vmv_peer_exchange_params;
int vmv_peer_exchange(const vmv_peer_exchange_params* p, void* stream);

typedef struct vmv_peer_allreduce_params {
    double* data;                         /* local [n], in/out */
    double* slots[VMV_PEER_MAX_RANKS];    /* [world][n] in every rank's arena */
    void* flags[VMV_PEER_MAX_RANKS];
    void* epoch;
    int32_t world, rank, n, nowait;
} vmv_peer_allreduce_params;
int vmv_peer_allreduce_f64(const vmv_peer_allreduce_params* p, void* stream);

/* All-gather over peer memory: this rank's [nouter][inner_bytes] block is stored at dst[q] + dst_offset_bytes +
 * outer * dst_outer_stride_bytes in EVERY rank q's arena, then the ranks meet at the epoch flags (same protocol and
 * control-line layout as vmv_peer_exchange).  One call per UNet evaluation assembles the [cfg half][B,C,F,h,w] outputs of
 * all frame shards -- and of the cond / uncond halves when the classifier-free-guidance pair (diffusion_ddim.py:149-155)
 * is split over two rank groups -- on every rank. */
typedef struct vmv_peer_allgather_params {
    const void* src;                      /* local, contiguous [nouter][inner_bytes] */
    void* dst[VMV_PEER_MAX_RANKS];
    void* flags[VMV_PEER_MAX_RANKS];
    void* epoch; void* done;
    int32_t world, rank, nowait, pad_;
    int64_t nouter, inner_bytes, dst_offset_bytes, dst_outer_stride_bytes;
} vmv_peer_allgather_params;
int vmv_peer_allgather(const vmv_peer_allgather_params* p, void* stream);

/* CUDA IPC plumbing for the peer arenas: export = 64-byte handle of the allocation containing ptr + ptr's offset in it;
 * import = map a peer's allocation (peer access enabled lazily) and return base + offset. */
int vmv_ipc_export(const void* ptr, void* handle64, int64_t* offset);
int vmv_ipc_import(const void* handle64, int64_t offset, void** out);

/* sizeof() of the parameter structs as compiled, so a foreign-language binding can verify its mirrors. */
int vmv_sizeof_gemm_params(void);
int vmv_sizeof_gemm_scatter(void);
int vmv_sizeof_attn_params(void);
int vmv_sizeof_peer_exchange_params(void);
int vmv_sizeof_peer_allreduce_params(void);
int vmv_sizeof_gn_peer(void);
int vmv_sizeof_peer_allgather_params(void);

#ifdef __cplusplus
}
#endif
#endif /* VIDEOMV_B200_H_ */
