/* libstad.so — C ABI of the B200-native (sm_100a) Video-ViT encoder forward used by simple-tad.
 *
 * The reference (tue-mps/simple-tad) has no FFI: the boundary this path sits behind is the Python nn.Module API of
 * modeling_finetune.py / modeling_pretrain.py / flash_attention_class.py.  Each entry point below replaces the
 * library kernels one reference call site reaches (cuDNN / cuBLAS / ATen / flash-attn 2); the call site is cited as
 * file:line relative to the reference root (mf = modeling_finetune.py, mp = modeling_pretrain.py,
 * fac = flash_attention_class.py, ri = run_inference.py, ris = run_inference_simple.py).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; the caller (PyTorch) owns all memory, the library
 *     never allocates or frees user-visible memory and never synchronises the device;
 *   - "bf16" buffers are row-major __nv_bfloat16; float buffers are fp32;
 *   - every call launches asynchronously on `stream` (a cudaStream_t passed as void*), is CUDA-graph capturable,
 *     and returns 0 or a negative STAD_E_* code; stad_last_error() returns the thread-local message;
 *   - there is no CPU path and no non-sm_100 path: stad_init() refuses other architectures.
 */
#ifndef STAD_H_
#define STAD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STAD_ABI_VERSION 6

#if defined(__GNUC__)
#define STAD_API __attribute__((visibility("default")))
#else
#define STAD_API
#endif

typedef void* stad_stream_t; /* cudaStream_t */

enum {
  STAD_OK = 0,
  STAD_E_SHAPE = -1, /* unsupported / inconsistent sizes            (reference: Python assert, mf:188)   */
  STAD_E_ALIGN = -2, /* pointer or leading dimension not 16-byte aligned                                 */
  STAD_E_ARCH = -3,  /* device is not compute capability 10.x                                            */
  STAD_E_CUDA = -4   /* a CUDA runtime / driver call failed                                              */
};

/* Epilogues of stad_ln_gemm. */
enum {
  STAD_EPI_BIAS = 0,     /* y = LN(x) W^T + b                     -> QKV projection, mf:92 / mf:119            */
  STAD_EPI_BIAS_GELU = 1 /* y = GELU_erf(LN(x) W^T + b)           -> Mlp.fc1 + nn.GELU, mf:48-49               */
};

/* Where the clips of a batch live. */
enum {
  STAD_IN_CLIPS = 0, /* x[B, C, T, H, W] bf16 — the tensor VisionTransformer.forward receives (mf:332)          */
  STAD_IN_FRAMES = 1 /* frames[F, C, H, W] bf16 of ONE video; frame t of clip b = frames[start + b*stride + t*frame_step]
                        — the sliding window of ri:69-109 / ris:428-465 / dota.py:204-223 without materialising it;
                        frame_step = orig_fps / target_fps is the in-window subsampling of RegularSequencer
                        (dataset/sequencing.py:45-58: DADA-2000 is read at 30 fps and scored at 10 fps, dada.py:31)   */
};

typedef struct stad_input {
  const void* data; /* bf16 */
  int32_t mode;     /* STAD_IN_CLIPS | STAD_IN_FRAMES */
  int32_t n_frames; /* FRAMES: F (number of frames resident); CLIPS: ignored */
  int32_t start;    /* FRAMES: first frame of clip 0 */
  int32_t stride;   /* FRAMES: frame step between consecutive clips (1 = every window, dota.py:209) */
  int32_t frame_step; /* FRAMES: frame distance inside a clip; 0 or 1 = consecutive frames */
  int32_t tubelet_reuse; /* FRAMES (ABI v6): non-zero lets stad_vit_forward embed every DISTINCT tubelet of the batch once
                            and assemble the windows from them (consecutive stride-1 windows share 7 of their 8 tubelets
                            with the window two frames on, ri:97-101).  Used when tubelet * frame_step is a multiple of
                            stride and at least a quarter of the embeddings is shared; the tubelet embedding is then
                            rounded to bf16 before the position table is added (one extra rounding, within the path's
                            tolerance; 0 keeps the frames path bit-identical to the clips path) */
  const int32_t* window_starts; /* FRAMES, optional (ABI v6): DEVICE array [B] with the first frame of every clip; when
                                   non-NULL it replaces start + b * stride.  One batch can then hold windows of several
                                   videos laid end to end in the frame buffer (final_test over a dataset of videos,
                                   eff:385-463: the DataLoader batches windows across video boundaries, rff:311-314).
                                   The caller guarantees window_starts[b] + (frames - 1) * frame_step < n_frames;
                                   frames beyond the buffer read as zeros (tensor-map bounds), never out of bounds. */
} stad_input;

/* Geometry of one model (PatchEmbed mf:172-183, VisionTransformer mf:211-234). */
typedef struct stad_dims {
  int32_t img_h, img_w; /* 224 */
  int32_t patch;        /* 16 */
  int32_t tubelet;      /* 2 */
  int32_t frames;       /* 16  (all_frames) */
  int32_t in_chans;     /* 3 */
  int32_t dim;          /* D: 384 / 768 / 1024 */
  int32_t depth;        /* L */
  int32_t heads;        /* H  (head dim is fixed at 64) */
  int32_t hidden;       /* mlp hidden = 4 D */
  int32_t num_classes;  /* 2 (0 = encoder only: return tokens after `norm`, mp:107) */
} stad_dims;

/* One transformer Block (mf:137-166) after weight preparation:
 *   LayerNorm gamma is folded into the following weight, beta into its bias (W' = W diag(gamma), b' = b + W beta),
 *   colsum[n] = sum_k bf16(W'[n,k]) — so  LN(x) W^T + b == rstd * (x W'^T - mean * colsum) + b'   (SURVEY §7 step 4). */
typedef struct stad_block {
  const void* w_qkv;   /* [3D, D] bf16, norm1-folded;  rows q|k|v (mf:69)                       */
  const float* b_qkv;  /* [3D] = cat(q_bias, 0, v_bias) + W beta1   (mf:88-90)                  */
  const float* cs_qkv; /* [3D] */
  const void* w_proj;  /* [D, D] bf16 (mf:78)  */
  const float* b_proj; /* [D] */
  const void* w_fc1;   /* [4D, D] bf16, norm2-folded (mf:42) */
  const float* b_fc1;  /* [4D] */
  const float* cs_fc1; /* [4D] */
  const void* w_fc2;   /* [D, 4D] bf16 (mf:44) */
  const float* b_fc2;  /* [D] */
} stad_block;

/* What the classifier returns (VisionTransformer.forward_features mf:323-330, `final_reduction`). */
enum {
  STAD_REDUCE_MEAN = 0, /* fc_norm(mean over tokens)  'fc_norm', mf:325-326: every simple-tad script            */
  STAD_REDUCE_CLS = 1,  /* norm(x)[:, 0]              'cls',     mf:327-328                                      */
  STAD_REDUCE_NONE = 2  /* norm(x), every token       'none',    mf:329-330: logits / probs / features per token */
};

typedef struct stad_model {
  stad_dims dims;
  const void* w_patch;      /* [D, C*tubelet*patch*patch] bf16 = Conv3d weight flattened (c,dt,dh,dw)  (mf:181-183) */
  const float* pos_bias;    /* [N, D] fp32 = sinusoid table (mf:195-205) + conv bias                               */
  const stad_block* blocks; /* HOST array of `depth` entries                                                       */
  const float* norm_g;      /* final LayerNorm: fc_norm (mf:270, classifier) or norm (mp:59, encoder)              */
  const float* norm_b;
  const float* w_head; /* [num_classes, D] fp32 (mf:272) or NULL */
  const float* b_head; /* [num_classes] */
  float eps;           /* 1e-6 (mf:342) */
  float attn_scale;    /* head_dim ** -0.5 (mf:67) */
  int32_t reduction;   /* STAD_REDUCE_* (classifier models; 0 = the fc_norm path)                                  */
  const float* cls_token; /* [D] fp32 or NULL: class token of the MVD sibling (other_models/MVD/modeling_finetune.py
                             :364-366), prepended to every clip after the position add (:431-435); each clip then has
                             n_tok + 1 rows and STAD_REDUCE_MEAN averages the patch tokens only (:447-449)          */
} stad_model;

/* PretrainVisionTransformer (mp:183-291) after weight preparation: encoder (visible tokens) -> encoder_to_decoder ->
 * decoder over all N tokens -> pixel head on the masked tokens.  The encoder's `norm` (mp:59,107) is folded into
 * encoder_to_decoder and the decoder's `norm` (mp:143,174) into the pixel head, exactly like norm1/norm2 in stad_block. */
typedef struct stad_mae_model {
  stad_model encoder;       /* dims.num_classes = 0; norm_g / norm_b / w_head / b_head are not read here             */
  stad_dims dec_dims;       /* dim = D_dec (192 / 384 / 512), depth, heads, hidden; num_classes = 1536 pixel values;
                               geometry fields as the encoder's                                                      */
  const void* w_e2d;        /* [D_dec, D_enc] bf16 = encoder_to_decoder.weight diag(norm.weight)   (mp:253, mp:281)  */
  const float* b_e2d;       /* [D_dec] = encoder_to_decoder.weight norm.bias   (the Linear itself has no bias)       */
  const float* cs_e2d;      /* [D_dec] column sums of w_e2d                                                          */
  const float* pos_dec;     /* [N, D_dec] fp32 sinusoid table of the decoder width                 (mp:257)          */
  const float* mask_token;  /* [D_dec] fp32                                                        (mp:255)          */
  const stad_block* dec_blocks; /* HOST array of dec_dims.depth entries                            (mp:133-139)      */
  const void* w_pix;        /* [1536, D_dec] bf16 = decoder.head.weight diag(decoder.norm.weight)  (mp:143-144)      */
  const float* b_pix;       /* [1536] = decoder.head.bias + decoder.head.weight decoder.norm.bias                    */
  const float* cs_pix;      /* [1536] */
} stad_mae_model;

/* Where stad_vit_forward writes its results; any pointer may be NULL (nothing is written for it). */
typedef struct stad_outputs {
  float* logits;   /* [B, num_classes]  head output, mf:334                                   (classifier models) */
  float* probs;    /* [B, num_classes]  softmax(logits), ris:381 / ri:107                     (classifier models) */
  float* features; /* [B, D]            forward_features, mf:326 / mf:328                     (classifier models)
                      STAD_REDUCE_NONE: the three are per token, [B, S, num_classes] / [B, S, D] (mf:330, mf:334)  */
  float* tokens;   /* [B, n_tok, D]     tokens after `norm`, mp:107-108                       (encoder models)    */
} stad_outputs;

/* ---- library ------------------------------------------------------------------------------------------------- */
STAD_API int stad_abi_version(void);
/* Binds nothing, checks that `device` is sm_100 and raises the kernels' shared-memory limits. Idempotent. */
STAD_API int stad_init(int device);
STAD_API const char* stad_last_error(void);

/* ---- per-launch timing (measurement only) ------------------------------------------------------------------------
 * When enabled, every kernel the library launches is bracketed by a pair of CUDA events recorded on the launching
 * stream.  stad_profile_read synchronises on the recorded events and returns one record per launch, in launch order.
 * Not CUDA-graph capturable; leave disabled (the default) outside benchmarks. */
enum {
  STAD_K_CAST = 0, STAD_K_GATHER = 1, STAD_K_GEMM = 2, STAD_K_ATTENTION = 3, STAD_K_ROW_STATS = 4,
  STAD_K_LAYERNORM = 5, STAD_K_POOL = 6, /* ROW_STATS with epi = 1: stad_stats_finalize */
  STAD_K_ASSEMBLE = 7, STAD_K_TAIL = 8, STAD_K_NORMALIZE = 9, STAD_K_EVAL = 10, STAD_K_RESIZE = 11
};
typedef struct stad_profile_record {
  int32_t kind;    /* STAD_K_*                                                        */
  int32_t epi;     /* GEMM: 1 LN-fold | 2 GELU | 4 residual | 8 pos table | 16 emits LN partial sums | 32 patch-embed A operand */
  int32_t m, n, k; /* GEMM: M, N, K; attention: B, H, S; row kernels: rows, cols, 0  */
  float ms;        /* device time between the two events                             */
} stad_profile_record;
STAD_API int stad_profile_enable(int capacity); /* capacity launches are kept; 0 disables and frees the events */
STAD_API int stad_profile_read(stad_profile_record* out, int max_records); /* returns #records, resets the log */

/* ---- element / row kernels (HBM-bound) ------------------------------------------------------------------------ */
/* fp32 -> bf16 cast of a clip batch; stands in for torch.cuda.amp.autocast's input cast (eff:428, te:177). */
STAD_API int stad_cast_f32_bf16(const float* x, void* y, size_t n, stad_stream_t stream);

/* Per-row LayerNorm statistics (mean, rstd) of x[M, D] bf16 -> stats[M] float2.   nn.LayerNorm stats, mf:143/149. */
STAD_API int stad_row_stats(const void* x, float* stats, int M, int D, float eps, stad_stream_t stream);

/* Full LayerNorm y = (x - mean) * rstd * g + b of x[M, D] bf16 -> y[M, D] fp32.   encoder `norm`, mp:107. */
STAD_API int stad_layernorm(const void* x, const float* g, const float* b, float* y, int M, int D, float eps,
                   stad_stream_t stream);

/* mean over tokens -> fc_norm -> head (-> softmax).   mf:325-326, mf:334, ris:381.
 * x[B, N, D] bf16; logits[B, C] fp32; probs[B, C] fp32 or NULL; features[B, D] fp32 (the fc_norm output that
 * forward_features returns, mf:326) or NULL; scratch: >= B * 16 * D floats. */
STAD_API int stad_pool_norm_head(const void* x, const float* g, const float* b, const float* w_head, const float* b_head,
                        float* logits, float* probs, float* features, float* scratch, int B, int N, int D, int C,
                        float eps, stad_stream_t stream);

/* LayerNorm of selected rows -> head -> softmax: `norm` followed by final_reduction 'cls' (x[:, 0], mf:327-328:
 * R = B, row_stride = tokens per clip, row_off = 0) or 'none' (every token, mf:329-330: R = B * S, row_stride = 1) and
 * the head (mf:334).  Row r is x[(r * row_stride + row_off), :D] bf16.  logits / probs [R, C], features [R, D] fp32;
 * probs and features may be NULL; w_head NULL (then b_head / logits are not read): features only. */
STAD_API int stad_rows_norm_head(const void* x, const float* g, const float* b, const float* w_head, const float* b_head,
                        float* logits, float* probs, float* features, int R, long long row_stride, long long row_off,
                        int D, int C, float eps, stad_stream_t stream);

/* x[B, N + 1, D] bf16 = cat(cls_token[D] fp32, emb[B, N, D] bf16) per clip, and the LayerNorm statistics (mean, rstd)
 * float[B * (N + 1), 2] of every row of x.   other_models/MVD/modeling_finetune.py:431-435 (use_cls_token). */
STAD_API int stad_prepend_cls(const void* emb, const float* cls_token, void* x, float* stats, int B, int N, int D, float eps,
                     stad_stream_t stream);

/* ---- tensor-core kernels (tcgen05 / TMEM / TMA) ---------------------------------------------------------------- */
/* Tubelet patch embedding: Conv3d(k = s = (tubelet, patch, patch)) as an im2col-free GEMM + pos/bias table add.
 *   PatchEmbed.forward mf:185-191 and the position add mf:312-313 (mp:93-95 for the encoder).
 * tok_idx == NULL : all N tokens of every clip, out[B*N, D] bf16, token order (t', h', w').
 * tok_idx != NULL : int32[B, n_tok] token ids (row-major order of surviving tokens, mp:98); only those tokens are
 *                   embedded; `gather` must hold B*n_tok*K bf16 of scratch. */
STAD_API int stad_patch_embed(const stad_input* in, const void* w, const float* pos_bias, const int32_t* tok_idx, void* out,
                     void* gather, const stad_dims* dims, int B, int n_tok, stad_stream_t stream);

/* y[M, N] = epilogue( rstd[m] * (x W'^T - mean[m] * colsum[n]) + bias[n] ), bf16 in / bf16 out, fp32 accumulate.
 *   norm1 -> qkv (mf:92,119);  norm2 -> fc1 -> GELU (mf:48-49).   stats from stad_row_stats. */
STAD_API int stad_ln_gemm(const void* x, const float* stats, const void* w, const float* bias, const float* colsum,
                 int epilogue, void* out, int M, int N, int K, stad_stream_t stream);

/* y[M, N] = a W^T + bias (+ residual), bf16 in / bf16 out.   attn.proj + residual (mf:104/128, mf:161);
 *   fc2 + residual (mf:52, mf:162).  residual may be NULL (plain Linear) and may alias out. */
STAD_API int stad_gemm_bias_residual(const void* a, const void* w, const float* bias, const void* residual, void* out, int M,
                            int N, int K, stad_stream_t stream);

/* ---- LayerNorm statistics without a pass over x ------------------------------------------------------------------
 * The GEMM that writes the residual stream x also emits, per row, `parts` partial (sum, sum of squares) pairs of the
 * bf16 values it stored (one pair per column tile and epilogue warpgroup, laid out [parts, M]); stad_stats_finalize
 * adds them in index order (deterministic, no atomics) into the (mean, rstd) that stad_ln_gemm takes.  Replaces
 * stad_row_stats on the whole-model path.   nn.LayerNorm statistics of norm1 / norm2, mf:143/149; x = x + ..., mf:161-162. */
/* Number of partial pairs per row that stad_gemm_bias_residual_stats writes for an [M, N] output. */
STAD_API int stad_stat_parts(int M, int N);
/* stad_gemm_bias_residual + statistics: stat_parts is float[parts, M, 2], parts = stad_stat_parts(M, N). */
STAD_API int stad_gemm_bias_residual_stats(const void* a, const void* w, const float* bias, const void* residual, void* out,
                                  float* stat_parts, int M, int N, int K, stad_stream_t stream);
/* stat_parts float[parts, M, 2] -> stats float[M, 2] = (mean, rstd) over D columns, rstd = 1 / sqrt(var + eps). */
STAD_API int stad_stats_finalize(const float* stat_parts, int parts, float* stats, int M, int D, float eps,
                        stad_stream_t stream);

/* Joint space-time softmax(q k^T * scale) v over packed qkv[B, S, 3, H, 64] bf16 -> out[B, S, H*64] bf16.
 *   Same tensor contract as FlashAttention.forward (fac:26-51, qkv "(B, S, 3, H, D)") and the math of
 *   Attention._naive_attn mf:93-103.  No N x N matrix is written to memory. */
STAD_API int stad_attention(const void* qkv, void* out, int B, int H, int S, float scale, stad_stream_t stream);

/* ---- whole forward ---------------------------------------------------------------------------------------------- */
/* Bytes of scratch stad_vit_forward needs for a batch of B clips with n_tok tokens each (a model with a class token:
 * pass n_tok + 1). */
STAD_API size_t stad_workspace_bytes(const stad_dims* dims, int B, int n_tok);

/* VisionTransformer.forward (mf:308-335) / PretrainVisionTransformerEncoder.forward_features (mp:91-108).
 *   tok_idx NULL  -> every token of every clip (n_tok must be the full token count).
 *   tok_idx given -> int32[B, n_tok] visible-token ids: the masked-encoder path.
 *   dims.num_classes > 0 -> classifier: out->logits required; out->probs / out->features optional.
 *   dims.num_classes == 0 -> encoder: out->tokens required.
 *   model->reduction selects what the classifier returns; model->cls_token != NULL prepends a class token (tok_idx
 *   must be NULL then; the workspace must hold n_tok + 1 tokens per clip).
 * Returns the number of kernels launched (>= 0) or a negative error. */
STAD_API int stad_vit_forward(const stad_model* model, const stad_input* in, const int32_t* tok_idx, int B, int n_tok,
                     const stad_outputs* out, void* workspace, size_t workspace_bytes, stad_stream_t stream);

/* ---- MAE pre-training forward (DAPT) ---------------------------------------------------------------------------- */
/* Decoder input of PretrainVisionTransformer.forward (mp:283-288):
 *   x_full[b, i] = vis[b, i]                                   for i <  n_vis  (x_vis + pos_emd_vis, the position rows
 *                                                              are added by the encoder_to_decoder GEMM epilogue)
 *   x_full[b, i] = mask_token + pos[mask_idx[b, i - n_vis]]    for i >= n_vis  (mask_token + pos_emd_mask)
 * plus the LayerNorm statistics (mean, rstd) of every row of x_full (norm1 of the first decoder block).
 * vis[B, n_vis, D] bf16, pos[N, D] fp32, mask_token[D] fp32, mask_idx int32[B, N - n_vis] (ids of the masked tokens in
 * row-major order = what expand_pos_embed[mask] keeps), x_full[B, N, D] bf16, stats float[B*N, 2]. */
STAD_API int stad_decoder_assemble(const void* vis, const float* pos, const float* mask_token, const int32_t* mask_idx,
                                   void* x_full, float* stats, int B, int N, int n_vis, int D, float eps,
                                   stad_stream_t stream);

/* y[B, n_keep, C] fp32 = last n_keep rows of every clip of x[B, N, C] bf16: `x[:, -return_token_num:]`, mp:174. */
STAD_API int stad_tail_rows_f32(const void* x, float* y, int B, int N, int n_keep, int C, stad_stream_t stream);

STAD_API size_t stad_mae_workspace_bytes(const stad_mae_model* model, int B, int n_vis);

/* PretrainVisionTransformer.forward(x, mask) (mp:276-291): pixels[B, N - n_vis, 1536] fp32, the decoder's predictions
 * for the masked tokens of every clip in row-major token order.  vis_idx int32[B, n_vis]: ids of the visible tokens
 * (x[~mask], mp:98); mask_idx int32[B, N - n_vis]: ids of the masked tokens (x[mask]); both row-major per clip.
 * Returns the number of kernels launched (>= 0) or a negative error. */
STAD_API int stad_mae_forward(const stad_mae_model* model, const stad_input* in, const int32_t* vis_idx,
                              const int32_t* mask_idx, int B, int n_vis, float* pixels, void* workspace, size_t workspace_bytes, stad_stream_t stream);

/* ---- frame preparation ------------------------------------------------------------------------------------------- */
/* uint8 HWC frames [F, H, W, 3] (bgr != 0: channel order of cv2.imread; 0: RGB) -> bf16 planes [F, 3, H, W] (RGB)
 *   = (v / 255 - mean[c]) / std[c].   prepare_image ri:15-34 (cvtColor + div 255 + normalise);  dota.py:357 +
 *   volume_transforms.ClipToTensor / video_transforms.normalize on the dataset path.
 * mean / std: HOST float[3] (RGB order).  The output is the frame buffer stad_input (STAD_IN_FRAMES) reads. */
STAD_API int stad_normalize_frames_u8(const void* frames_u8, void* out_bf16, int F, int H, int W, const float* mean,
                                      const float* std, int bgr, stad_stream_t stream);

/* uint8 HWC frames [F, H_s, W_s, 3] -> uint8 HWC [F, H_d, W_d, 3], bicubic: `cv2.resize(img, (W_d, H_d),
 * interpolation=cv2.INTER_CUBIC)` of ri:79-80 / dota.py:347-348 in OpenCV's 8-bit fixed-point arithmetic (imgproc
 * resize.cpp: taps with A = -0.75 in float32 -> 11-bit weights, clamped source taps, int32 passes, (v + 2^21) >> 22).
 * The tap tables are built by the caller the way resize.cpp builds them (simple_tad_b200.frames.cubic_taps):
 *   xofs int32[W_d], xw int16[W_d, 4]: source column of the second tap and the four weights; yofs / yw likewise.
 * Exact integer arithmetic.  (OpenCV's own result varies by 1 LSB with its build: its SIMD vertical pass works in
 * float32, and pip builds dispatch this resize to Intel IPP; see oracle/resize_oracle.py.) */
STAD_API int stad_resize_cubic_u8(const void* frames_u8, void* out_u8, int F, int Hs, int Ws, int Hd, int Wd,
                                  const int32_t* xofs, const int16_t* xw, const int32_t* yofs, const int16_t* yw,
                                  stad_stream_t stream);

/* ---- evaluation epilogue ------------------------------------------------------------------------------------------ */
/* Confusion counts of the per-frame risk probability probs[i][1] against T ascending fp32 thresholds
 * (engine_for_frame_finetuning.py:461-488 evaluates torchmetrics' binned AUROC / AP / PR / ROC on THRESHOLDS =
 * arange(0, 1.001, 0.01); anaysis/metrics.py:183-199 evaluates MCC / P / R / acc / F1 at every one of them).
 *   probs[n, 2] fp32 (softmax output of stad_vit_forward), labels int32[n] (non-zero = anomalous),
 *   hist  uint64[2][T + 1]: hist[y][b] = number of samples of label y with exactly b thresholds <= p, so
 *         TP_k = sum_{b > k} hist[1][b],  FP_k = sum_{b > k} hist[0][b]   (prediction `p >= t_k`, as the reference);
 *   conf  uint64[4] = tn, fp, fn, tp of the arg-max prediction (torch.max(softmax, 1), eff:464).
 * Counts are exact integers, additive over shards: ranks all-reduce (SUM) the two small tables instead of gathering
 * every prediction.  hist and conf are zeroed by the call. */
STAD_API int stad_eval_hist(const float* probs, const int32_t* labels, long long n, const float* thresholds, int T,
                            unsigned long long* hist, unsigned long long* conf, stad_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* STAD_H_ */
