// Internal launch interface between api.cu and the kernel translation units.
#pragma once
#include "common.h"

namespace stad {

// Geometry of the tubelet patch grid and of where the clips live (see stad_input in include/stad.h).
struct PatchGeom {
  int Tp, Hp, Wp;   // token grid: 8 x 14 x 14
  int hp_tile;      // h' rows per M-tile (Wp * hp_tile <= 128)
  int h_tiles;      // M-tiles per t' slot
  int C, T, tubelet;
  int img_h, img_w;
  int mode, start, stride;  // STAD_IN_CLIPS / STAD_IN_FRAMES
  int fstep;                // FRAMES: frame distance inside a clip (>= 1)
  int n_planes;             // extent of the plane dimension of the input tensor map
  const int32_t* win_start; // FRAMES, optional: device array [B], first frame of clip b (replaces start + b * stride)
};

enum : int { EPI_LN = 1, EPI_GELU = 2, EPI_RESID = 4, EPI_POS = 8, EPI_STATS = 16 };

// LayerNorm statistics without a pass over x: the epilogue of the GEMM that writes the residual stream x (patch embed,
// attn.proj, mlp.fc2) emits, for every output row and every (column tile, epilogue warpgroup), (sum, sum of squares)
// of the bf16 values it stored; launch_stats_finalize adds a row's partials in index order (deterministic, no atomics)
// and writes (mean, rstd) for the following LN-folded GEMM (nn.LayerNorm statistics, modeling_finetune.py:143/149).
constexpr int kMaxStatParts = 32;  // 2 * N / 64 at the narrowest column tile, N <= 1024
// up to this many partials per row an EPI_LN GEMM prefetches them a tile ahead (registers); see gemm.cu
constexpr int kGemmMaxFoldParts = 8;

struct GemmArgs {
  const bf16* a = nullptr;  // [M, K] (plain) or the bf16 plane tensor (patch mode)
  const bf16* w = nullptr;  // [N, K]
  int M = 0, N = 0, K = 0;
  int epi = 0;                       // EPI_* bits
  const float* bias = nullptr;       // [N] or null
  const float* colsum = nullptr;     // [N]   (EPI_LN)
  const float2* stats = nullptr;     // [M]   (EPI_LN) (mean, rstd); or, instead:
  const float2* stat_parts = nullptr;  // [n_stat_parts, M] partial (sum, sumsq) of the rows of `a` over its K columns,
  int n_stat_parts = 0;                //   as an EPI_STATS launch wrote them: the epilogue finishes the statistics itself
  float ln_eps = 1e-6f;                //   (LayerNorm eps for the stat_parts form)
  float2* stats_out = nullptr;       // [2 * n_tiles, M] (EPI_STATS) partial (sum, sumsq) of the stored rows
  const bf16* residual = nullptr;    // [M,N] (EPI_RESID)
  const float* pos = nullptr;        // [pos_rows, N] (EPI_POS)
  const int32_t* tok_idx = nullptr;  // [M] row -> pos row, or null: pos row = m % pos_rows
  int pos_rows = 0;
  bf16* out = nullptr;               // [M, N]
  const PatchGeom* patch = nullptr;  // non-null: A is gathered straight from the clip planes by 5-D TMA
};

int launch_gemm(const GemmArgs& g, cudaStream_t stream);
int gemm_stat_parts(int M, int N, bool patch, const PatchGeom* pg);  // partials per row an EPI_STATS launch writes
int gemm_init();  // raise dynamic smem limits

int launch_attention(const bf16* qkv, bf16* out, int B, int H, int S, float scale, cudaStream_t stream);
int attention_init();

int launch_cast_f32_bf16(const float* x, bf16* y, size_t n, cudaStream_t stream);
int launch_stats_finalize(const float2* parts, int n_parts, float2* stats, int M, int D, float eps, cudaStream_t stream);
int launch_row_stats(const bf16* x, float2* stats, int M, int D, float eps, cudaStream_t stream);
int launch_layernorm(const bf16* x, const float* g, const float* b, float* y, int M, int D, float eps,
                     cudaStream_t stream);
int launch_pool_norm_head(const bf16* x, const float* g, const float* b, const float* w_head, const float* b_head,
                          float* logits, float* probs, float* features, float* scratch, int B, int N, int D, int C,
                          float eps, cudaStream_t stream, size_t clip_stride = 0);  // 0: clips are dense, N * D apart
// final_reduction 'cls' / 'none': LayerNorm of rows r * row_stride + row_off (r < R) -> features / head / softmax
int launch_rows_norm_head(const bf16* x, const float* g, const float* b, const float* w_head, const float* b_head,
                          float* logits, float* probs, float* features, int R, long long row_stride, long long row_off,
                          int D, int C, float eps, cudaStream_t stream);
// MVD class token: x[B, N + 1, D] = cat(cls_token, emb[B, N, D]) + (mean, rstd) of every row
int launch_prepend_cls(const bf16* emb, const float* cls_token, bf16* x, float2* stats, int B, int N, int D, float eps,
                       cudaStream_t stream);
// visible-token im2col: out[B*n_tok, K] bf16 rows (c, dt, dh, dw) of the listed tokens
int launch_gather_patches(const bf16* planes, const PatchGeom& pg, const int32_t* tok_idx, bf16* out, int B,
                          int n_tok, cudaStream_t stream);

// tubelet-embedding reuse: residual stream of B overlapping windows from the embeddings of their distinct tubelets
// (window b, slot t' uses tubelet b + t' * step) + position table + LayerNorm statistics of every row
int launch_window_assemble(const bf16* emb, const float* pos_bias, bf16* x, float2* stats, int B, int Tp, int HW, int D,
                           int step, float eps, cudaStream_t stream);

// MAE decoder glue (modeling_pretrain.py:283-288, :174) and uint8 frame preparation (run_inference.py:15-34)
int launch_decoder_assemble(const bf16* vis, const float* pos, const float* mask_token, const int32_t* mask_idx, bf16* x,
                            float2* stats, int B, int N, int n_vis, int D, float eps, cudaStream_t stream);
int launch_tail_rows_f32(const bf16* x, float* y, int B, int N, int n_keep, int C, cudaStream_t stream);
int launch_normalize_u8(const uint8_t* in, bf16* out, int F, int H, int W, const float* mean, const float* std_, int bgr,
                        cudaStream_t stream);

// evaluation epilogue: threshold histogram of the risk probabilities (engine_for_frame_finetuning.py:461-488)
int launch_eval_hist(const float* probs, const int32_t* labels, long long n, const float* thresholds, int T,
                     unsigned long long* hist, unsigned long long* conf, cudaStream_t stream);

// bicubic uint8 frame resize in OpenCV's fixed-point arithmetic (run_inference.py:79-80, dota.py:347-348)
int launch_resize_cubic_u8(const uint8_t* in, uint8_t* out, int F, int Hs, int Ws, int Hd, int Wd, const int32_t* xofs,
                           const int16_t* xw, const int32_t* yofs, const int16_t* yw, cudaStream_t stream);

}  // namespace stad
