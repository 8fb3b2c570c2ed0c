/* mdk.h — C ABI of libmikudance_sm100.so: the B200 (sm_100a) kernels behind the MikuDance
 * denoising loop.
 *
 * The reference (Kebii/MikuDance) has no FFI on this path: every op below replaces a PyTorch /
 * diffusers call made from the Python modules cited per entry point (paths relative to the
 * reference checkout). The Python host side (the mikudance_b200 package, mirrored under src/ with the
 * reference's module paths) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions: plain C, caller owns every buffer (device pointers, normally storage of PyTorch
 * tensors), no allocation / host sync / default-stream use inside; every launch goes to the
 * `stream` argument (a cudaStream_t passed as void*), so calls are CUDA-graph capturable.
 * Return 0 on success, <0 on error; mdk_last_error() gives a thread-local message.
 * Activations are fp16, token-major ("NHWC"): an image batch is [(b f), h, w, C] == [M, C] rows.
 */
#ifndef MDK_H_
#define MDK_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDK_ABI_VERSION 1

typedef struct mdk_ctx mdk_ctx;

/* Context: caches device properties / kernel attributes for one device. */
int mdk_create(int device, mdk_ctx** out);
void mdk_destroy(mdk_ctx* ctx);
const char* mdk_last_error(void);
int mdk_abi_version(void);
/* number of kernel launches issued through this library since load (bench "gpu_launches") */
int64_t mdk_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * mdk_gemm_f16 — tcgen05 tensor-core GEMM / implicit-GEMM 3x3 convolution, fp16 in, fp32
 * accumulate in TMEM, fused epilogue.   D[m, n] = sum_k A[m, k] * B[n, k]  (+ epilogue)
 *
 * Replaces: nn.Linear / 1x1 conv / 3x3 conv dispatches of
 *   src/models/resnet.py:9-17,222,240,243 (InflatedConv3d conv1/conv2/conv_shortcut),
 *   src/models/transformer_3d.py:132,190 (proj_in/proj_out), src/models/motion_module.py:169,181,
 *   diffusers Attention.to_q/to_k/to_v/to_out and FeedForward/GEGLU (src/models/attention.py:323-364).
 *
 * A operand: one or two sources concatenated along K (the up-block skip concat of
 * src/models/unet_3d_blocks.py:735,870 is never materialised).
 *   conv_taps == 1: a_i is row-major [m, k_i] with leading dimension lda_i (elements).
 *   conv_taps == 9: a_i is an NHWC image batch [nimg, h, w, k_i] (dense), m = nimg*h*w, and
 *                   K = 9*(k0+k1) ordered (kh, kw, [c of a0 | c of a1]); stride 1, zero pad 1.
 * B operand: weights row-major [n, K] (K contiguous), ldb elements.
 * Epilogue, in this order, all in fp32 before one rounding to fp16:
 *   + bias[n] (fp32, optional) + row_bias[((row / row_div) % row_mod), n] (fp32, optional)
 *   geglu != 0: columns are tiled in blocks of mdk_gemm_geglu_block() = 256 packed as
 *               [128 value columns | 128 gate columns]; output has n/2 columns:
 *               out = value * gelu_erf(gate)   (diffusers GEGLU, src/models/attention.py:364)
 *   + residual[row, col] (fp16, optional, leading dimension ldr)
 * Output: the n (or n/2) result columns are split into up to 3 segments of seg_cols columns
 *   (seg_cols = 0: one segment). Segment s goes to out[s] with leading dimension ldo[s];
 *   out_trans[s] != 0 writes the segment transposed per image:
 *      out[s][((row / trans_rows) * seg_cols + col) * trans_ld + (row % trans_rows)]
 *   (used to emit V^T for the attention kernel's K-major B operand).
 * Requirements: k_i % 8 == 0, lda % 8 == 0, ldb % 8 == 0, 16-byte aligned pointers, n % 8 == 0.
 */
typedef struct {
  const void* a0;
  const void* a1;
  const void* b;
  int64_t lda0, lda1, ldb;
  int32_t k0, k1;
  int32_t m, n;
  int32_t conv_taps, nimg, h, w;
  const float* bias;
  const float* row_bias;
  int32_t row_div, row_mod;
  const void* residual;
  int64_t ldr;
  int32_t geglu;
  int32_t seg_cols;
  void* out[3];
  int64_t ldo[3];
  int32_t out_trans[3];
  int32_t trans_rows;
  int64_t trans_ld;
  /* transposed segments only: when trans_head_dp > trans_head_d > 0 the seg_cols channels are taken as
   * heads of trans_head_d channels and head h is written to rows [h*trans_head_dp, h*trans_head_dp +
   * trans_head_d) of an image (padded per-head V^T: the pad rows belong to the caller, e.g. a row of
   * ones that makes the attention kernel's P.V MMA also produce the softmax row sums). */
  int32_t trans_head_d, trans_head_dp;
} mdk_gemm_args;

int mdk_gemm_f16(mdk_ctx* ctx, const mdk_gemm_args* args, void* stream);
int mdk_gemm_geglu_block(void);

/* ------------------------------------------------------------------------------------------
 * mdk_attn_fwd_f16 — flash-style attention on tcgen05: S = Q K^T in TMEM, online softmax in
 * registers, P staged in shared memory, O += P V in TMEM.  Non-causal, no mask, no dropout.
 *
 * Replaces F.scaled_dot_product_attention under diffusers AttnProcessor2_0 as called from
 *   src/models/mutual_mix_attention.py:173-200 (spatial self-attention, K/V = LN(x)+bank) and
 *   src/models/mutual_mix_attention.py:213-220 (CLIP cross-attention, Lkv = 257).
 *
 * q   : [nimg, lq, heads*d] fp16 (row stride ldq elements)
 * k   : [nkv , lkv, heads*d] fp16 (row stride ldk)
 * vt  : [nkv , heads*vt_head_rows, ldvt] fp16 — V transposed per image (row = head*vt_head_rows +
 *       channel, col = kv index); vt_head_rows = 0 means d.  With vt_ones != 0 (needs d % 16 == 8 and
 *       vt_head_rows >= d + 8) row d of every head holds ones and the kernel takes the softmax row sums
 *       from that extra output column of P.V instead of adding them up in the softmax warps.
 * out : [nimg, lq, heads*d] fp16 (row stride ldo)
 * image i attends to kv batch (i / kv_div)   (kv_div = 1: self-attention; = frames per branch
 * for the shared CLIP context). scale = softmax scale (d^-0.5).
 * Requirements: d % 8 == 0, d <= 192.
 */
typedef struct {
  const void* q;
  const void* k;
  const void* vt;
  void* out;
  int64_t ldq, ldk, ldvt, ldo;
  int32_t nimg, nkv, kv_div;
  int32_t lq, lkv, heads, d;
  float scale;
  int32_t vt_head_rows, vt_ones;
} mdk_attn_args;

int mdk_attn_fwd_f16(mdk_ctx* ctx, const mdk_attn_args* args, void* stream);

/* Diagnostics (never on the product path): with MDK_ATTN_TRACE=1 in the environment, launches of the
 * ones-row 128-key attention kernel run an instrumented instantiation in which one CTA writes clock64()
 * at the hand-off points of every key tile into buf[tile*16 + slot] (device memory, int64, `tiles` tiles;
 * slots are listed in csrc/attn_tc.cu).  buf = NULL switches it off. */
int mdk_attn_debug_trace(void* buf, int32_t tiles);

/* ------------------------------------------------------------------------------------------
 * mdk_temporal_attn_f16 — attention across the frame axis at a fixed pixel (AnimateDiff motion
 * module). Replaces VersatileAttention.forward, src/models/motion_module.py:364-439 (the
 * "(b f) d c -> (b d) f c" transposes are folded into strided addressing).
 *
 * q   : fp16 rows [(b*f_q + i)*npix + px] with row stride q_ld, the query at column q_off
 *       (normally the fused to_q|to_k|to_v GEMM output, q_off = 0) for the f_q local query frames
 * kv  : fp16 K/V rows of all f_kv frames: frame j of batch b is row
 *       (((j / f_kv_rank)*nb + b)*f_kv_rank + j % f_kv_rank)*npix + px, row stride kv_ld, K at
 *       column k_off, V at v_off.  Single GPU: kv = q buffer, f_kv_rank = f_kv (0 means f_kv).
 *       Frame-sharded: kv = the NCCL all-gather of every rank's local rows, f_kv_rank = frames/rank.
 * pe_q: [>= f_q_offset+f_q, C] fp32 — (pe @ Wq^T): the positional term of the query only
 *       (motion_module.py:404-417: PE is added to the query input only), may be NULL
 * out : [nb*f_q, npix, C] fp16 (row stride out_ld)
 */
typedef struct {
  const void* q;
  const void* kv;
  const float* pe_q;
  void* out;
  int64_t q_ld, kv_ld, out_ld;
  int32_t q_off, k_off, v_off;
  int32_t nb, f_q, f_kv, f_kv_rank, f_q_offset, npix, heads, d;
  float scale;
} mdk_tattn_args;

int mdk_temporal_attn_f16(mdk_ctx* ctx, const mdk_tattn_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * GroupNorm (per image, NHWC) — replaces InflatedGroupNorm / nn.GroupNorm + SiLU:
 *   src/models/resnet.py:20-28,220-221,231-236; src/models/transformer_3d.py:60-62,130;
 *   src/models/motion_module.py:121-123,160; src/models/unet_3d_mix.py:591-592.
 * Input is one or two channel-concatenated NHWC sources [nimg, hw, c_i].
 * mdk_groupnorm_f16: stats (fp32 partials per CTA, fixed-order fp64 combine: bitwise deterministic, no
 *   atomics) + apply (+SiLU) -> out [nimg, hw, c0+c1].
 * ws: workspace of mdk_groupnorm_ws_bytes(nimg, groups) bytes.
 */
typedef struct {
  const void* x0;
  const void* x1;
  int32_t c0, c1;
  int32_t nimg, hw, groups;
  float eps;
  const void* gamma; /* fp16 [c0+c1] */
  const void* beta;  /* fp16 [c0+c1] */
  int32_t silu;
  void* out;
  void* ws;
  /* exchange layout (frame-sharded motion modules, mikudance_b200/sharding.py): out_chunk_pix > 0 writes pixel px
   * of image i at row ((px / out_chunk_pix) * nimg + i) * out_chunk_pix + px % out_chunk_pix of an
   * [out_chunks, nimg, out_chunk_pix, C] buffer — the send buffer of the frames -> pixels all-to-all, one
   * contiguous block per destination rank (rows of pixels >= hw are not written).  0: [nimg, hw, C]. */
  int32_t out_chunk_pix, out_chunks;
} mdk_gn_args;

int64_t mdk_groupnorm_ws_bytes(int32_t nimg, int32_t groups);
int mdk_groupnorm_f16(mdk_ctx* ctx, const mdk_gn_args* args, void* stream);

/* LayerNorm over the channel axis of [rows, c] tokens (eps 1e-5), optional second output
 * out2[r] = LN(x)[r] + add[r] for rows r >= add_row0 (the reference-feature "bank" add of
 * src/models/mutual_mix_attention.py:169-172; rows < add_row0 are the CFG uncond half and are
 * not written). Replaces nn.LayerNorm at src/models/attention.py:331-366,
 * src/models/motion_module.py:236,243. */
typedef struct {
  const void* x;
  int64_t rows;
  int32_t c;
  float eps;
  const void* gamma;
  const void* beta;
  void* out;
  const void* add; /* fp16 [rows - add_row0, c] or NULL */
  void* out2;      /* fp16 [rows - add_row0, c] or NULL */
  int64_t add_row0;
} mdk_ln_args;

int mdk_layernorm_f16(mdk_ctx* ctx, const mdk_ln_args* args, void* stream);

/* Nearest x2 upsample of NHWC [nimg, h, w, c] -> [nimg, 2h, 2w, c]
 * (F.interpolate(scale=[1,2,2], nearest), src/models/resnet.py:70-73). */
int mdk_upsample2x_f16(mdk_ctx* ctx, const void* x, void* out, int32_t nimg, int32_t h, int32_t w,
                       int32_t c, void* stream);

/* im2col for the 3x3 convolutions that are not run as implicit GEMM: stride-2 pad-1 Downsample3D
 * (src/models/resnet.py:106-120) and the Cin=4 conv_in (src/models/unet_3d_mix.py:94-96,503).
 * x: NHWC [nimg, h, w, c]; out: [nimg*ho*wo, kpad] with column (kh*3+kw)*c + ci, zero padded. */
int mdk_im2col3x3_f16(mdk_ctx* ctx, const void* x, void* out, int32_t nimg, int32_t h, int32_t w,
                      int32_t c, int32_t stride, int32_t kpad, void* stream);

/* Timestep embedding: sinusoid(320) -> Linear -> SiLU -> Linear -> SiLU (the activation every
 * ResnetBlock3D applies first), then all time_emb_proj rows at once:
 *   temb_out[r] = proj_w[r] . silu(emb) + proj_b[r] + conv1_bias[r]
 * Replaces Timesteps/TimestepEmbedding (src/models/unet_3d_mix.py:99-102,482-488) and
 * time_emb_proj (src/models/resnet.py:226-229). The timestep is read from device memory so a
 * captured CUDA graph can be replayed for every DDIM step. */
typedef struct {
  const int64_t* timestep; /* device, 1 element */
  int32_t dim;             /* 320 */
  int32_t flip_sin_to_cos;
  float freq_shift;
  const void* w1; const void* b1; /* fp16 [edim, dim], [edim] */
  const void* w2; const void* b2; /* fp16 [edim, edim], [edim] */
  int32_t edim;                   /* 1280 */
  const void* proj_w;             /* fp16 [nrows, edim]: all resnets' time_emb_proj stacked */
  const float* proj_b;            /* fp32 [nrows]: time_emb_proj.bias + conv1.bias */
  int32_t nrows;
  float* scratch;                 /* fp32 [2*edim + dim] */
  float* temb_out;                /* fp32 [nrows] */
} mdk_temb_args;

int mdk_time_embed_f16(mdk_ctx* ctx, const mdk_temb_args* args, void* stream);

/* Latent layout changes at the UNet boundary.
 * mdk_latents_to_nhwc: sample [b_src, c, F, h, w] fp16, frames frame_idx[0..fl) (device int32, NULL =
 *   identity) -> [(b fl), h, w, cpad]; output batch i reads source batch i % b_src (b_src = 1
 *   duplicates the latents for the two CFG branches, src/pipelines/pipeline_mikudance.py:626-630)
 * mdk_pred_accumulate: UNet output [(b fl), h, w, cpad] fp16 -> acc[b, c, frame_idx[j], h, w] += ,
 *   counter[frame_idx[j]] += 1  (window accumulate, src/pipelines/pipeline_mikudance.py:662-664) */
int mdk_latents_to_nhwc(mdk_ctx* ctx, const void* sample, void* out, int32_t b, int32_t b_src,
                        int32_t c, int32_t f_total, const int32_t* frame_idx, int32_t fl,
                        int32_t hw, int32_t cpad, void* stream);
int mdk_pred_accumulate(mdk_ctx* ctx, const void* pred, float* acc, float* counter, int32_t b,
                        int32_t c, int32_t f_total, const int32_t* frame_idx, int32_t fl,
                        int32_t hw, int32_t cpad, void* stream);

/* Back from the pixel-sharded layout of a frame-sharded motion module, fused with the module's residual add
 * (src/models/motion_module.py:188): back [chunks, nimg, chunk_pix, C] (what the pixels -> frames all-to-all
 * received: chunk d = pixels [d * chunk_pix, (d+1) * chunk_pix) of this rank's nimg images), x [nimg, hw, C]
 *   out[i, px, :] = back[px / chunk_pix, i, px % chunk_pix, :] + x[i, px, :]        (fp32 add, one rounding)
 * x == NULL: the layout change alone (bit-exact copy). */
int mdk_unshard_add_f16(mdk_ctx* ctx, const void* back, const void* x, void* out, int32_t nimg, int32_t hw,
                        int32_t chunk_pix, int32_t c, void* stream);

/* Window average + classifier-free guidance + DDIM (eta = 0) update in one pass, fp32 math:
 *   eps = acc / counter;  g = eps_u + s (eps_c - eps_u)
 *   v-prediction:  x0 = sqrt(a_t) x - sqrt(1-a_t) g ;  e = sqrt(a_t) g + sqrt(1-a_t) x
 *   epsilon     :  x0 = (x - sqrt(1-a_t) g)/sqrt(a_t); e = g
 *   x_prev = sqrt(a_prev) x0 + sqrt(1-a_prev) e
 * Replaces src/pipelines/pipeline_mikudance.py:670-678 + diffusers DDIMScheduler.step.
 * coef: device fp32 [4] = {sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)}.
 * acc: fp32 [nb, c, f, hw] with nb = 2 (uncond, cond) when guidance is on, else 1. */
int mdk_cfg_ddim_step(mdk_ctx* ctx, const float* acc, const float* counter, void* latents,
                      const float* coef, float guidance_scale, int32_t nb, int32_t c, int32_t f,
                      int32_t hw, int32_t v_prediction, void* stream);

/* ------------------------------------------------------------------------------------------
 * Reference UNet (the "writer" that produces the feature banks). It is the
 * SD-1.5 2-D UNet, so it runs on mdk_gemm_f16 / mdk_groupnorm_f16 / mdk_layernorm_f16 /
 * mdk_attn_fwd_f16 above; only the three ops below are specific to it.
 *
 * mdk_cond_to_nhwc_f16 — channel slice of the NCHW condition latents, zero-padded NHWC, optionally
 *   nearest-resized. Replaces `sample[:, :-2]`, `sample[:, -2:]` (src/models/unet_2d_mix.py:1208-1209)
 *   and F.interpolate(motion_map, size=x.size()[2:], mode="nearest") (src/models/man_module.py:28).
 *   x: [nimg, ctot, h, w] fp16; out: [nimg, ho, wo, cpad] fp16 holding channels
 *   [c_first, c_first + c) sampled at (floor(oy*h/ho), floor(ox*w/wo)); columns >= c are zero.
 * mdk_relu_f16 — in-place ReLU over n fp16 values (n % 8 == 0): MANModule.mlp_shared's activation
 *   (src/models/man_module.py:18-20).
 * mdk_man_modulate_f16 — MANModule.forward's normalisation and modulation
 *   (src/models/man_module.py:26,32): parameter-free InstanceNorm2d (per image and channel over the hw
 *   pixels, biased variance) followed by  out = normalized * (1 + gamma) + beta.
 *   x, out: [nimg, hw, c] fp16; gb: [nimg*hw, ldgb] fp16 with gamma in columns [0, c) and beta in
 *   [c, 2c) (the fused mlp_gamma | mlp_beta convolution output); ws: mdk_man_ws_bytes(nimg, c) bytes.
 *   Deterministic (per-CTA partial sums, fixed-order fp64 combine, no atomics).
 */
int mdk_cond_to_nhwc_f16(mdk_ctx* ctx, const void* x, void* out, int32_t nimg, int32_t ctot,
                         int32_t c_first, int32_t c, int32_t h, int32_t w, int32_t ho, int32_t wo,
                         int32_t cpad, void* stream);
int mdk_relu_f16(mdk_ctx* ctx, void* x, int64_t n, void* stream);

typedef struct {
  const void* x;
  const void* gb;
  int64_t ldgb;
  int32_t nimg, hw, c;
  float eps;
  void* out;
  void* ws;
} mdk_man_args;

int64_t mdk_man_ws_bytes(int32_t nimg, int32_t c);
int mdk_man_modulate_f16(mdk_ctx* ctx, const mdk_man_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * CLIP image encoder (transformers CLIPVisionModelWithProjection as used at
 * src/pipelines/pipeline_mikudance.py:405-417).  It runs on mdk_gemm_f16 (patch embedding as a GEMM whose
 * row bias is the position embedding; q|k|v, out, fc1, fc2, visual_projection), mdk_layernorm_f16 and
 * mdk_attn_fwd_f16; the only op of its own is the MLP activation:
 * mdk_quick_gelu_f16 — in place x * sigmoid(1.702 x) over n fp16 values (n % 8 == 0). */
int mdk_quick_gelu_f16(mdk_ctx* ctx, void* x, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * VAE (diffusers AutoencoderKL as used at src/pipelines/pipeline_mikudance.py:115-130,
 * 455-549).  Convolutions, GroupNorm+SiLU, the nearest upsample and the linears run on the entry points above;
 * two ops are its own:
 * mdk_softmax_rows_f16 — in-place softmax over the first `cols` columns of each of `rows` rows (row stride ld,
 *   fp16 storage, fp32 math).  The mid-block attention has ONE head of 512 channels (beyond the flash kernel's
 *   head size), so it runs as S = Q K^T (GEMM, the softmax scale folded into the q projection), this row softmax,
 *   O = P V (GEMM against V^T).
 * mdk_im2col3x3_ex_f16 — im2col of a 3x3 convolution with pad_lo (0 or 1) zero rows/columns in front and one
 *   behind: Downsample2D(padding=0) of the encoder pads right/bottom only (F.pad(x, (0,1,0,1)), stride 2).
 *   Output [nimg*ho*wo, kpad], ho = (h + pad_lo - 2)/stride + 1, column (kh*3+kw)*c + ci, zero padded. */
int mdk_softmax_rows_f16(mdk_ctx* ctx, void* x, int64_t rows, int32_t cols, int64_t ld, void* stream);
int mdk_im2col3x3_ex_f16(mdk_ctx* ctx, const void* x, void* out, int32_t nimg, int32_t h, int32_t w, int32_t c,
                         int32_t stride, int32_t pad_lo, int32_t kpad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MDK_H_ */
