// extern "C" surface of libstad.so (declared in include/stad.h) and the host-side sequencing of one forward.
#include <cstdlib>
#include <mutex>

#include "kernels.h"

namespace stad {

namespace {

std::mutex g_init_mutex;
bool g_inited[64] = {};  // per device: cudaFuncSetAttribute (dynamic shared memory opt-in) is a per-device setting

inline cudaStream_t as_stream(stad_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int check_dims(const stad_dims* d) {
  STAD_CHECK_ARG(d != nullptr, "dims is NULL");
  STAD_CHECK_ARG(d->patch == 16, "patch size %d unsupported (every reference factory uses 16, mf:338-398)", d->patch);
  STAD_CHECK_ARG(d->tubelet >= 1 && d->frames % d->tubelet == 0, "frames=%d not divisible by tubelet=%d", d->frames,
                 d->tubelet);
  STAD_CHECK_ARG(d->img_h % 16 == 0 && d->img_w % 16 == 0 && d->img_h > 0 && d->img_w > 0,
                 "image %dx%d is not a multiple of the patch size", d->img_h, d->img_w);
  STAD_CHECK_ARG(d->img_w / 16 <= 128, "image width %d too large for one M-tile", d->img_w);
  STAD_CHECK_ARG(d->in_chans >= 1, "in_chans=%d", d->in_chans);
  STAD_CHECK_ARG(d->dim > 0 && d->dim % 64 == 0, "embed dim %d must be a multiple of 64", d->dim);
  return STAD_OK;
}

int make_geom(const stad_dims* d, const stad_input* in, int B, PatchGeom* pg) {
  int rc = check_dims(d);
  if (rc) return rc;
  STAD_CHECK_ARG(in != nullptr && in->data != nullptr, "input is NULL");
  pg->Tp = d->frames / d->tubelet;
  pg->Hp = d->img_h / 16;
  pg->Wp = d->img_w / 16;
  const int max_rows = 128 / pg->Wp;  // h' rows that fit one 128-row MMA tile
  int hp_tile = 1;                    // largest divisor of Hp that fits: every tile then holds the same token count
  for (int d = 1; d <= max_rows && d <= pg->Hp; ++d)
    if (pg->Hp % d == 0) hp_tile = d;
  pg->hp_tile = hp_tile;              // 224 px: 7 (98 tokens / tile); 384 px: 4; 512 px: 4
  pg->h_tiles = pg->Hp / hp_tile;
  pg->C = d->in_chans;
  pg->T = d->frames;
  pg->tubelet = d->tubelet;
  pg->img_h = d->img_h;
  pg->img_w = d->img_w;
  pg->mode = in->mode;
  pg->start = in->start;
  pg->stride = in->stride;
  pg->fstep = in->frame_step > 1 ? in->frame_step : 1;
  pg->win_start = nullptr;
  if (in->mode == STAD_IN_CLIPS) {
    pg->n_planes = B * d->in_chans * d->frames;
    pg->start = 0;
    pg->stride = 0;
  } else if (in->mode == STAD_IN_FRAMES) {
    STAD_CHECK_ARG(in->stride >= 1 && in->start >= 0, "frames input: start=%d stride=%d", in->start, in->stride);
    STAD_CHECK_ARG(in->frame_step >= 0, "frames input: frame_step=%d", in->frame_step);
    if (in->window_starts != nullptr) {
      // explicit first frames (device memory: their range is the caller's contract, see stad.h)
      STAD_CHECK_ARG((d->frames - 1) * pg->fstep < in->n_frames, "frames input: a clip spans %d frames but only %d are resident",
                     (d->frames - 1) * pg->fstep + 1, in->n_frames);
      if (reinterpret_cast<uintptr_t>(in->window_starts) & 3) return fail(STAD_E_ALIGN, "window_starts must be 4-byte aligned");
      pg->win_start = in->window_starts;
    } else {
      STAD_CHECK_ARG(in->start + (B - 1) * in->stride + (d->frames - 1) * pg->fstep < in->n_frames,
                     "frames input: clip %d needs frame %d but only %d frames are resident", B - 1,
                     in->start + (B - 1) * in->stride + (d->frames - 1) * pg->fstep, in->n_frames);
    }
    pg->n_planes = in->n_frames * d->in_chans;
  } else {
    return fail(STAD_E_SHAPE, "unknown input mode %d", in->mode);
  }
  return STAD_OK;
}

// When the LN-folded GEMMs finish the LayerNorm statistics themselves instead of a stats_finalize launch (see
// run_blocks).  With at most kGemmMaxFoldParts partials per row they are prefetched a tile ahead (registers), off the
// head of the epilogue: folding then pays up to a few ten thousand rows (DAPT, 16000 rows: 3.08 -> 2.89 ms per 100
// clips); at the 100352 rows of the headline batch the 24 extra-long epilogues still cost more than the 24 finalize
// launches they replace (same box, profiles/r2b_fold_ab.txt: 2742-2764 vs 2770-2774 clips/s).  With more partials
// (narrow column tiles, i.e. small batches) the loads stay at the head of each tile: up to 8192 rows
// (profiles/r1c_fold_threshold_ab.txt).
constexpr int kFoldStatsMaxRows = 8192;
constexpr int kFoldStatsMaxRowsPrefetched = 32768;
bool fold_stats_for(int M, int parts) {
  return M <= (parts <= kGemmMaxFoldParts ? kFoldStatsMaxRowsPrefetched : kFoldStatsMaxRows);
}

size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

struct Workspace {
  bf16* x;       // [M, D]   residual stream
  float2* stats; // [M]            LayerNorm (mean, rstd) of x
  float2* parts; // [2 D / 64, M]  partial (sum, sumsq) written by the GEMM that produced x
  bf16* qkv;     // [M, 3D]
  bf16* attn;    // [M, D]   (directly after qkv)
  bf16* hidden;  // [M, 4D]  aliases qkv+attn: qkv/attn are dead once proj has run
  float* pool;   // [B, 16, D]
  bf16* gather;  // [M, K]   visible-token im2col (masked path only)
  size_t bytes;
};

Workspace carve(const stad_dims* d, int B, int n_tok, void* base) {
  Workspace w;
  const size_t M = static_cast<size_t>(B) * n_tok;
  const size_t D = d->dim;
  const size_t K = static_cast<size_t>(d->in_chans) * d->tubelet * 256;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align256(bytes);
    return o;
  };
  uint8_t* p = static_cast<uint8_t*>(base);
  const size_t o_x = take(M * D * 2);
  const size_t o_stats = take(M * sizeof(float2));
  const size_t o_parts = take(M * (2 * D / 64) * sizeof(float2));
  const size_t hid = static_cast<size_t>(d->hidden) > 4 * D ? d->hidden : 4 * D;
  const size_t o_big = take(M * hid * 2);
  const size_t o_pool = take(static_cast<size_t>(B) * 16 * D * sizeof(float));
  const int n_full = (d->frames / d->tubelet) * (d->img_h / 16) * (d->img_w / 16);
  // im2col scratch only on the visible-token path (n_full + 1 rows: every token plus a class token)
  const size_t o_gather = take(n_tok == n_full || n_tok == n_full + 1 ? 0 : M * K * 2);
  w.x = reinterpret_cast<bf16*>(p + o_x);
  w.stats = reinterpret_cast<float2*>(p + o_stats);
  w.parts = reinterpret_cast<float2*>(p + o_parts);
  w.qkv = reinterpret_cast<bf16*>(p + o_big);
  w.attn = w.qkv + M * 3 * D;
  w.hidden = w.qkv;
  w.pool = reinterpret_cast<float*>(p + o_pool);
  w.gather = reinterpret_cast<bf16*>(p + o_gather);
  w.bytes = off;
  return w;
}

int patch_embed_impl(const stad_input* in, const void* w, const float* pos_bias, const int32_t* tok_idx, void* out,
                     void* gather, const stad_dims* d, int B, int n_tok, cudaStream_t stream, int* launches,
                     float2* stats_out = nullptr, int* stat_parts = nullptr) {
  PatchGeom pg;
  int rc = make_geom(d, in, B, &pg);
  if (rc) return rc;
  const int N_full = pg.Tp * pg.Hp * pg.Wp;
  const int K = d->in_chans * d->tubelet * 256;
  GemmArgs g;
  g.w = static_cast<const bf16*>(w);
  g.N = d->dim;
  g.K = K;
  g.epi = EPI_POS;
  g.pos = pos_bias;
  g.out = static_cast<bf16*>(out);
  if (tok_idx == nullptr) STAD_CHECK_ARG(n_tok == N_full, "patch_embed: n_tok=%d but the clip has %d tokens", n_tok, N_full);
  // n_tok == N_full with an index list: every token survives, so the list is the identity (mp:98 keeps row-major order)
  if (n_tok == N_full) {
    g.a = static_cast<const bf16*>(in->data);
    g.M = B * N_full;
    g.pos_rows = N_full;
    g.patch = &pg;
    if (stats_out) {
      g.epi |= EPI_STATS;
      g.stats_out = stats_out;
      *stat_parts = gemm_stat_parts(g.M, g.N, true, &pg);
    }
    if ((rc = launch_gemm(g, stream))) return rc;
    *launches += 1;
  } else {
    STAD_CHECK_ARG(n_tok >= 1 && n_tok <= N_full, "patch_embed: n_tok=%d out of range", n_tok);
    STAD_CHECK_ARG(gather != nullptr, "patch_embed: visible-token mode needs gather scratch");
    if ((rc = launch_gather_patches(static_cast<const bf16*>(in->data), pg, tok_idx, static_cast<bf16*>(gather), B,
                                    n_tok, stream)))
      return rc;
    g.a = static_cast<const bf16*>(gather);
    g.M = B * n_tok;
    g.tok_idx = tok_idx;
    g.pos_rows = N_full;
    if (stats_out) {
      g.epi |= EPI_STATS;
      g.stats_out = stats_out;
      *stat_parts = gemm_stat_parts(g.M, g.N, false, nullptr);
    }
    if ((rc = launch_gemm(g, stream))) return rc;
    *launches += 2;
  }
  return STAD_OK;
}


// The transformer stack: `depth` x Block.forward (mf:159-162) over the residual stream ws.x [B * n_tok, D].
//   parts_in > 0 : ws.parts holds that many LayerNorm partial sums per row of ws.x (written by the GEMM that produced x)
//   parts_in == 0: ws.stats already holds (mean, rstd) of ws.x (e.g. written by launch_decoder_assemble)
//   final_stats  : the last fc2 also emits partial sums (a LayerNorm-folded GEMM follows the stack); on return
//                  ws.parts holds gemm_stat_parts(M, D) partials per row.
int run_blocks(const stad_block* blocks, const stad_dims* d, float eps, float attn_scale, const Workspace& ws, int B,
               int n_tok, int parts_in, bool final_stats, cudaStream_t stream, int* launches) {
  const int M = B * n_tok;
  const int D = d->dim;
  const int parts_resid = gemm_stat_parts(M, D, false, nullptr);
  int parts = parts_in;
  int rc;
  for (int l = 0; l < d->depth; ++l) {
    const stad_block& blk = blocks[l];
    const bool last = l + 1 == d->depth;
    const bool emit = !last || final_stats;
    // x = x + proj(attn(norm1(x)))                                     (mf:161)
    // LayerNorm statistics from the partial sums of the GEMM that wrote x.  Few rows (small batch, launch-latency
    // bound): the LN-folded GEMM finishes them in its own epilogue, one launch less per LayerNorm (batch 1: 0.95 ->
    // 0.89 ms).  Many rows: a 5 us finalize kernel, because the extra loads sit on the critical path of every tile's
    // epilogue of kernels that are epilogue-bound (batch 64: folding costs 2.7 %).  parts == 0: ws.stats is ready.
    GemmArgs q;
    q.a = ws.x; q.w = static_cast<const bf16*>(blk.w_qkv); q.M = M; q.N = 3 * D; q.K = D;
    q.epi = EPI_LN; q.bias = blk.b_qkv; q.colsum = blk.cs_qkv; q.out = ws.qkv; q.ln_eps = eps;
    if (parts > 0 && fold_stats_for(M, parts)) {
      q.stat_parts = ws.parts; q.n_stat_parts = parts;
    } else {
      if (parts > 0) {
        if ((rc = launch_stats_finalize(ws.parts, parts, ws.stats, M, D, eps, stream))) return rc;
        ++*launches;
      }
      q.stats = ws.stats;
    }
    if ((rc = launch_gemm(q, stream))) return rc;
    if ((rc = launch_attention(ws.qkv, ws.attn, B, d->heads, n_tok, attn_scale, stream))) return rc;
    GemmArgs pr;
    pr.a = ws.attn; pr.w = static_cast<const bf16*>(blk.w_proj); pr.M = M; pr.N = D; pr.K = D;
    pr.epi = EPI_RESID | EPI_STATS; pr.bias = blk.b_proj; pr.residual = ws.x; pr.out = ws.x; pr.stats_out = ws.parts;
    if ((rc = launch_gemm(pr, stream))) return rc;
    parts = parts_resid;
    // x = x + fc2(gelu(fc1(norm2(x))))                                 (mf:162)
    GemmArgs f1;
    f1.a = ws.x; f1.w = static_cast<const bf16*>(blk.w_fc1); f1.M = M; f1.N = d->hidden; f1.K = D;
    f1.epi = EPI_LN | EPI_GELU; f1.bias = blk.b_fc1; f1.colsum = blk.cs_fc1; f1.out = ws.hidden; f1.ln_eps = eps;
    if (fold_stats_for(M, parts)) {
      f1.stat_parts = ws.parts; f1.n_stat_parts = parts;
    } else {
      if ((rc = launch_stats_finalize(ws.parts, parts, ws.stats, M, D, eps, stream))) return rc;
      ++*launches;
      f1.stats = ws.stats;
    }
    if ((rc = launch_gemm(f1, stream))) return rc;
    GemmArgs f2;
    f2.a = ws.hidden; f2.w = static_cast<const bf16*>(blk.w_fc2); f2.M = M; f2.N = D; f2.K = d->hidden;
    f2.epi = emit ? (EPI_RESID | EPI_STATS) : EPI_RESID; f2.bias = blk.b_fc2; f2.residual = ws.x; f2.out = ws.x;
    f2.stats_out = emit ? ws.parts : nullptr;
    if ((rc = launch_gemm(f2, stream))) return rc;
    *launches += 5;
  }
  return STAD_OK;
}

// Workspace of the whole MAE forward: encoder (B x n_vis tokens) | encoder_to_decoder output | decoder (B x N tokens).
struct MaeWorkspace {
  Workspace enc, dec;
  bf16* vis;  // [B * n_vis, D_dec]
  bf16* pix;  // [B * N, 1536]  pixel head over every row, before the masked rows are widened to fp32
  size_t bytes;
};

MaeWorkspace carve_mae(const stad_mae_model* m, int B, int n_vis, int n_full, void* base) {
  MaeWorkspace w;
  uint8_t* p = static_cast<uint8_t*>(base);
  w.enc = carve(&m->encoder.dims, B, n_vis, p);
  size_t off = w.enc.bytes;
  w.vis = reinterpret_cast<bf16*>(p + off);
  off += align256(static_cast<size_t>(B) * n_vis * m->dec_dims.dim * 2);
  w.pix = reinterpret_cast<bf16*>(p + off);
  off += align256(static_cast<size_t>(B) * n_full * m->dec_dims.num_classes * 2);
  w.dec = carve(&m->dec_dims, B, n_full, p ? p + off : nullptr);
  w.bytes = off + w.dec.bytes;
  return w;
}

int full_tokens(const stad_dims* d) { return (d->frames / d->tubelet) * (d->img_h / 16) * (d->img_w / 16); }

}  // namespace
}  // namespace stad

using namespace stad;

extern "C" {

int stad_abi_version(void) { return STAD_ABI_VERSION; }

const char* stad_last_error(void) { return last_error(); }

int stad_init(int device) {
  std::lock_guard<std::mutex> lock(g_init_mutex);
  int count = 0;
  STAD_CUDA_OK(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return fail(STAD_E_SHAPE, "stad_init: device %d out of range (%d visible)", device, count);
  int major = 0, minor = 0;
  STAD_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  STAD_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10)
    return fail(STAD_E_ARCH, "stad_init: device %d is sm_%d%d; libstad is built for sm_100a only and has no other path",
                device, major, minor);
  STAD_CUDA_OK(cudaSetDevice(device));
  if (device < 64 && g_inited[device]) return STAD_OK;
  int rc;
  if ((rc = gemm_init())) return rc;
  if ((rc = attention_init())) return rc;
  if (device < 64) g_inited[device] = true;
  return STAD_OK;
}

int stad_profile_enable(int capacity) { return prof_enable(capacity); }
int stad_profile_read(stad_profile_record* out, int max_records) {
  if (out == nullptr || max_records <= 0) return fail(STAD_E_SHAPE, "stad_profile_read: empty output");
  return prof_read(out, max_records);
}

int stad_cast_f32_bf16(const float* x, void* y, size_t n, stad_stream_t stream) {
  return launch_cast_f32_bf16(x, static_cast<bf16*>(y), n, as_stream(stream));
}

int stad_row_stats(const void* x, float* stats, int M, int D, float eps, stad_stream_t stream) {
  return launch_row_stats(static_cast<const bf16*>(x), reinterpret_cast<float2*>(stats), M, D, eps, as_stream(stream));
}

int stad_layernorm(const void* x, const float* g, const float* b, float* y, int M, int D, float eps,
                   stad_stream_t stream) {
  return launch_layernorm(static_cast<const bf16*>(x), g, b, y, M, D, eps, as_stream(stream));
}

int stad_pool_norm_head(const void* x, const float* g, const float* b, const float* w_head, const float* b_head,
                        float* logits, float* probs, float* features, float* scratch, int B, int N, int D, int C,
                        float eps, stad_stream_t stream) {
  return launch_pool_norm_head(static_cast<const bf16*>(x), g, b, w_head, b_head, logits, probs, features, scratch, B,
                               N, D, C, eps, as_stream(stream));
}

int stad_rows_norm_head(const void* x, const float* g, const float* b, const float* w_head, const float* b_head,
                        float* logits, float* probs, float* features, int R, long long row_stride, long long row_off,
                        int D, int C, float eps, stad_stream_t stream) {
  return launch_rows_norm_head(static_cast<const bf16*>(x), g, b, w_head, b_head, logits, probs, features, R, row_stride,
                               row_off, D, C, eps, as_stream(stream));
}

int stad_prepend_cls(const void* emb, const float* cls_token, void* x, float* stats, int B, int N, int D, float eps,
                     stad_stream_t stream) {
  return launch_prepend_cls(static_cast<const bf16*>(emb), cls_token, static_cast<bf16*>(x),
                            reinterpret_cast<float2*>(stats), B, N, D, eps, as_stream(stream));
}

int stad_patch_embed(const stad_input* in, const void* w, const float* pos_bias, const int32_t* tok_idx, void* out,
                     void* gather, const stad_dims* dims, int B, int n_tok, stad_stream_t stream) {
  int launches = 0;
  return patch_embed_impl(in, w, pos_bias, tok_idx, out, gather, dims, B, n_tok, as_stream(stream), &launches);
}

int stad_ln_gemm(const void* x, const float* stats, const void* w, const float* bias, const float* colsum,
                 int epilogue, void* out, int M, int N, int K, stad_stream_t stream) {
  STAD_CHECK_ARG(epilogue == STAD_EPI_BIAS || epilogue == STAD_EPI_BIAS_GELU, "ln_gemm: unknown epilogue %d", epilogue);
  GemmArgs g;
  g.a = static_cast<const bf16*>(x);
  g.w = static_cast<const bf16*>(w);
  g.M = M;
  g.N = N;
  g.K = K;
  g.epi = EPI_LN | (epilogue == STAD_EPI_BIAS_GELU ? EPI_GELU : 0);
  g.bias = bias;
  g.colsum = colsum;
  g.stats = reinterpret_cast<const float2*>(stats);
  g.out = static_cast<bf16*>(out);
  return launch_gemm(g, as_stream(stream));
}

int stad_gemm_bias_residual(const void* a, const void* w, const float* bias, const void* residual, void* out, int M,
                            int N, int K, stad_stream_t stream) {
  GemmArgs g;
  g.a = static_cast<const bf16*>(a);
  g.w = static_cast<const bf16*>(w);
  g.M = M;
  g.N = N;
  g.K = K;
  g.epi = residual ? EPI_RESID : 0;
  g.bias = bias;
  g.residual = static_cast<const bf16*>(residual);
  g.out = static_cast<bf16*>(out);
  return launch_gemm(g, as_stream(stream));
}

int stad_stat_parts(int M, int N) {
  if (M <= 0 || N <= 0 || N % 64 != 0) return fail(STAD_E_SHAPE, "stat_parts: M=%d N=%d", M, N);
  return gemm_stat_parts(M, N, false, nullptr);
}

int stad_gemm_bias_residual_stats(const void* a, const void* w, const float* bias, const void* residual, void* out,
                                  float* stat_parts, int M, int N, int K, stad_stream_t stream) {
  STAD_CHECK_ARG(stat_parts != nullptr, "gemm_bias_residual_stats: stat_parts is NULL");
  GemmArgs g;
  g.a = static_cast<const bf16*>(a);
  g.w = static_cast<const bf16*>(w);
  g.M = M;
  g.N = N;
  g.K = K;
  g.epi = (residual ? EPI_RESID : EPI_POS) | EPI_STATS;
  STAD_CHECK_ARG(residual != nullptr, "gemm_bias_residual_stats: the statistics epilogue exists for the residual GEMMs");
  g.bias = bias;
  g.residual = static_cast<const bf16*>(residual);
  g.out = static_cast<bf16*>(out);
  g.stats_out = reinterpret_cast<float2*>(stat_parts);
  return launch_gemm(g, as_stream(stream));
}

int stad_stats_finalize(const float* stat_parts, int parts, float* stats, int M, int D, float eps, stad_stream_t stream) {
  return launch_stats_finalize(reinterpret_cast<const float2*>(stat_parts), parts, reinterpret_cast<float2*>(stats), M, D,
                               eps, as_stream(stream));
}

int stad_attention(const void* qkv, void* out, int B, int H, int S, float scale, stad_stream_t stream) {
  return launch_attention(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), B, H, S, scale, as_stream(stream));
}

size_t stad_workspace_bytes(const stad_dims* dims, int B, int n_tok) {
  if (dims == nullptr || B <= 0 || n_tok <= 0) return 0;
  return carve(dims, B, n_tok, nullptr).bytes;
}

int stad_vit_forward(const stad_model* m, const stad_input* in, const int32_t* tok_idx, int B, int n_tok,
                     const stad_outputs* out, void* workspace, size_t workspace_bytes, stad_stream_t stream_) {
  STAD_CHECK_ARG(out != nullptr, "vit_forward: outputs is NULL");
  float* logits = out->logits;
  float* probs = out->probs;
  float* features = out->features;
  float* tokens_out = out->tokens;
  STAD_CHECK_ARG(m != nullptr && m->blocks != nullptr, "vit_forward: model is NULL");
  STAD_CHECK_ARG(B > 0 && n_tok > 0, "vit_forward: B=%d n_tok=%d", B, n_tok);
  const stad_dims* d = &m->dims;
  int rc = check_dims(d);
  if (rc) return rc;
  STAD_CHECK_ARG(d->heads * 64 == d->dim, "vit_forward: head dim must be 64 (dim=%d heads=%d)", d->dim, d->heads);
  STAD_CHECK_ARG(workspace != nullptr, "vit_forward: workspace is NULL");
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(STAD_E_ALIGN, "vit_forward: workspace must be 256-byte aligned");
  const bool with_cls = m->cls_token != nullptr;
  const int S = n_tok + (with_cls ? 1 : 0);  // rows per clip in the residual stream
  if (with_cls)
    STAD_CHECK_ARG(tok_idx == nullptr && n_tok == full_tokens(d),
                   "vit_forward: a class token needs every patch token (n_tok=%d of %d, no index list)", n_tok,
                   full_tokens(d));
  STAD_CHECK_ARG(m->reduction >= STAD_REDUCE_MEAN && m->reduction <= STAD_REDUCE_NONE, "vit_forward: reduction=%d",
                 m->reduction);
  Workspace ws = carve(d, B, S, workspace);
  STAD_CHECK_ARG(ws.bytes <= workspace_bytes, "vit_forward: workspace too small (%zu < %zu bytes)", workspace_bytes,
                 ws.bytes);
  const bool classifier = d->num_classes > 0;
  if (classifier) STAD_CHECK_ARG(logits && m->w_head && m->b_head, "vit_forward: classifier path needs logits/head");
  else STAD_CHECK_ARG(tokens_out, "vit_forward: encoder path needs tokens_out");
  STAD_CHECK_ARG(m->norm_g && m->norm_b, "vit_forward: final norm weights missing");

  cudaStream_t stream = as_stream(stream_);
  const int M = B * S;
  const int D = d->dim;
  int launches = 0;

  // PatchEmbed + position table (mf:309-313 / mp:93-98).  Its epilogue also emits the LayerNorm partial sums of the
  // rows it stores, as does every GEMM below that writes the residual stream: no separate statistics pass.
  int parts = 0;
  // Tubelet-embedding reuse across overlapping windows (ri:97-101): window b, slot t' holds the tubelet
  // that starts at frame start + b * stride + t' * tubelet * frame_step.  When tubelet * frame_step is a multiple of the
  // window stride these first frames form ONE arithmetic progression start + u * stride, u = b + t' * step, so the
  // batch has n_u = (B - 1) + (Tp - 1) * step + 1 distinct tubelets instead of B * Tp (stride-1 windows of a 16-frame,
  // tubelet-2 model: 64 + 14 instead of 512).  Each is embedded once (the patch GEMM over n_u one-tubelet "clips", no
  // position table, into the not-yet-used hidden buffer) and one row kernel lays out the windows, adds the position
  // table and writes the LayerNorm statistics of norm1 of the first block.
  int reuse_step = -1, n_u = 0;
  const int Tp = d->frames / d->tubelet;
  if (!with_cls && tok_idx == nullptr && in != nullptr && in->mode == STAD_IN_FRAMES && in->tubelet_reuse &&
      in->window_starts == nullptr && in->stride >= 1 && n_tok == full_tokens(d)) {
    const int tf = d->tubelet * (in->frame_step > 1 ? in->frame_step : 1);
    if (tf % in->stride == 0) {
      const int step = tf / in->stride;
      n_u = (B - 1) + (Tp - 1) * step + 1;
      if (4ll * n_u <= 3ll * B * Tp) reuse_step = step;  // at least a quarter of the embeddings is shared
    }
  }
  if (reuse_step >= 0) {
    stad_dims d1 = *d;
    d1.frames = d->tubelet;  // a "clip" of one tubelet: Tp = 1
    if ((rc = patch_embed_impl(in, m->w_patch, /*pos_bias=*/nullptr, nullptr, ws.hidden, nullptr, &d1, n_u, full_tokens(&d1),
                               stream, &launches)))
      return rc;
    const int HW = (d->img_h / 16) * (d->img_w / 16);
    if ((rc = launch_window_assemble(ws.hidden, m->pos_bias, ws.x, ws.stats, B, Tp, HW, D, reuse_step, m->eps, stream)))
      return rc;
    launches += 1;
  } else if (!with_cls) {
    if ((rc = patch_embed_impl(in, m->w_patch, m->pos_bias, tok_idx, ws.x, ws.gather, d, B, n_tok, stream, &launches,
                               ws.parts, &parts)))
      return rc;
  } else {
    // MVD class token (MVD mf:431-435): the patch rows go to the (still unused) hidden buffer, one row kernel lays
    // out [cls | patches] per clip in the residual stream and writes the statistics of norm1 of the first block.
    if ((rc = patch_embed_impl(in, m->w_patch, m->pos_bias, nullptr, ws.hidden, ws.gather, d, B, n_tok, stream,
                               &launches)))
      return rc;
    if ((rc = launch_prepend_cls(ws.hidden, m->cls_token, ws.x, ws.stats, B, n_tok, D, m->eps, stream))) return rc;
    launches += 1;
  }
  if ((rc = run_blocks(m->blocks, d, m->eps, m->attn_scale, ws, B, S, parts, /*final_stats=*/false, stream,
                       &launches)))
    return rc;

  if (classifier && m->reduction == STAD_REDUCE_MEAN) {
    // norm = Identity; mean over (patch) tokens; fc_norm; head        (mf:323-326, mf:334; MVD mf:447-449)
    if ((rc = launch_pool_norm_head(ws.x + (with_cls ? D : 0), m->norm_g, m->norm_b, m->w_head, m->b_head, logits, probs,
                                    features, ws.pool, B, n_tok, D, d->num_classes, m->eps, stream,
                                    static_cast<size_t>(S) * D)))
      return rc;
    launches += 2;
  } else if (classifier) {
    // norm over the rows that are returned; head                      (mf:323, mf:327-330, mf:334)
    const bool cls_only = m->reduction == STAD_REDUCE_CLS;
    if ((rc = launch_rows_norm_head(ws.x, m->norm_g, m->norm_b, m->w_head, m->b_head, logits, probs, features,
                                    cls_only ? B : M, cls_only ? S : 1, 0, D, d->num_classes, m->eps, stream)))
      return rc;
    launches += 1;
  } else {
    // encoder: norm over every visible token, head = Identity         (mp:107, mp:112)
    if ((rc = launch_layernorm(ws.x, m->norm_g, m->norm_b, tokens_out, M, D, m->eps, stream))) return rc;
    launches += 1;
  }
  return launches;
}

int stad_decoder_assemble(const void* vis, const float* pos, const float* mask_token, const int32_t* mask_idx,
                          void* x_full, float* stats, int B, int N, int n_vis, int D, float eps, stad_stream_t stream) {
  STAD_CHECK_ARG(vis && pos && mask_token && mask_idx && x_full && stats, "decoder_assemble: NULL argument");
  return launch_decoder_assemble(static_cast<const bf16*>(vis), pos, mask_token, mask_idx, static_cast<bf16*>(x_full),
                                 reinterpret_cast<float2*>(stats), B, N, n_vis, D, eps, as_stream(stream));
}

int stad_tail_rows_f32(const void* x, float* y, int B, int N, int n_keep, int C, stad_stream_t stream) {
  STAD_CHECK_ARG(x && y, "tail_rows: NULL argument");
  return launch_tail_rows_f32(static_cast<const bf16*>(x), y, B, N, n_keep, C, as_stream(stream));
}

int stad_normalize_frames_u8(const void* frames_u8, void* out_bf16, int F, int H, int W, const float* mean,
                             const float* std_, int bgr, stad_stream_t stream) {
  STAD_CHECK_ARG(frames_u8 && out_bf16 && mean && std_, "normalize_frames_u8: NULL argument");
  return launch_normalize_u8(static_cast<const uint8_t*>(frames_u8), static_cast<bf16*>(out_bf16), F, H, W, mean, std_,
                             bgr, as_stream(stream));
}

size_t stad_mae_workspace_bytes(const stad_mae_model* m, int B, int n_vis) {
  if (m == nullptr || B <= 0 || n_vis <= 0) return 0;
  return carve_mae(m, B, n_vis, full_tokens(&m->encoder.dims), nullptr).bytes;
}

int stad_mae_forward(const stad_mae_model* m, const stad_input* in, const int32_t* vis_idx, const int32_t* mask_idx, int B,
                     int n_vis, float* pixels, void* workspace, size_t workspace_bytes, stad_stream_t stream_) {
  STAD_CHECK_ARG(m != nullptr && m->encoder.blocks != nullptr && m->dec_blocks != nullptr, "mae_forward: model is NULL");
  STAD_CHECK_ARG(vis_idx != nullptr && mask_idx != nullptr && pixels != nullptr, "mae_forward: vis_idx / mask_idx / pixels is NULL");
  const stad_dims* de = &m->encoder.dims;
  const stad_dims* dd = &m->dec_dims;
  int rc = check_dims(de);
  if (rc) return rc;
  const int N = full_tokens(de);
  STAD_CHECK_ARG(B > 0 && n_vis > 0 && n_vis < N, "mae_forward: B=%d n_vis=%d (N=%d)", B, n_vis, N);
  STAD_CHECK_ARG(de->heads * 64 == de->dim && dd->heads * 64 == dd->dim,
                 "mae_forward: head dim must be 64 (encoder %d/%d, decoder %d/%d)", de->dim, de->heads, dd->dim, dd->heads);
  STAD_CHECK_ARG(dd->dim % 64 == 0 && dd->num_classes > 0 && dd->num_classes % 64 == 0,
                 "mae_forward: decoder dim=%d classes=%d", dd->dim, dd->num_classes);
  STAD_CHECK_ARG(m->w_e2d && m->b_e2d && m->cs_e2d && m->pos_dec && m->mask_token && m->w_pix && m->b_pix && m->cs_pix,
                 "mae_forward: decoder weights missing");
  STAD_CHECK_ARG(workspace != nullptr, "mae_forward: workspace is NULL");
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(STAD_E_ALIGN, "mae_forward: workspace must be 256-byte aligned");
  MaeWorkspace ws = carve_mae(m, B, n_vis, N, workspace);
  STAD_CHECK_ARG(ws.bytes <= workspace_bytes, "mae_forward: workspace too small (%zu < %zu bytes)", workspace_bytes,
                 ws.bytes);
  cudaStream_t stream = as_stream(stream_);
  const float eps = m->encoder.eps, scale = m->encoder.attn_scale;
  int launches = 0;

  // ---- encoder over the visible tokens (mp:91-106)
  int parts = 0;
  if ((rc = patch_embed_impl(in, m->encoder.w_patch, m->encoder.pos_bias, vis_idx, ws.enc.x, ws.enc.gather, de, B, n_vis,
                             stream, &launches, ws.enc.parts, &parts)))
    return rc;
  if ((rc = run_blocks(m->encoder.blocks, de, eps, scale, ws.enc, B, n_vis, parts, /*final_stats=*/true, stream,
                       &launches)))
    return rc;
  // ---- x_vis = encoder_to_decoder(norm(x)) + pos_emd_vis              (mp:107, mp:281, mp:287)
  const int Mv = B * n_vis;
  GemmArgs e;
  e.a = ws.enc.x; e.w = static_cast<const bf16*>(m->w_e2d); e.M = Mv; e.N = dd->dim; e.K = de->dim;
  e.epi = EPI_LN | EPI_POS; e.bias = m->b_e2d; e.colsum = m->cs_e2d; e.out = ws.vis; e.ln_eps = eps;
  e.stat_parts = ws.enc.parts; e.n_stat_parts = gemm_stat_parts(Mv, de->dim, false, nullptr);
  e.pos = m->pos_dec; e.tok_idx = vis_idx; e.pos_rows = N;
  if ((rc = launch_gemm(e, stream))) return rc;
  // ---- x_full = cat(x_vis, mask_token + pos_emd_mask) + statistics     (mp:283-288)
  if ((rc = launch_decoder_assemble(ws.vis, m->pos_dec, m->mask_token, mask_idx, ws.dec.x, ws.dec.stats, B, N, n_vis,
                                    dd->dim, eps, stream)))
    return rc;
  launches += 2;
  // ---- decoder blocks over all N tokens                                 (mp:165-171)
  if ((rc = run_blocks(m->dec_blocks, dd, eps, scale, ws.dec, B, N, 0, /*final_stats=*/true, stream, &launches)))
    return rc;
  // ---- head(norm(x[:, -n_mask:]))                                       (mp:173-174)
  // The LayerNorm-folded head runs over every row (the visible rows, 10 % at mask ratio 0.9, are computed and dropped:
  // rows of one clip are contiguous, the masked rows of the batch are not); the masked rows of its bf16 result are
  // then widened to fp32 into `pixels`.
  const int Mf = B * N;
  GemmArgs h;
  h.a = ws.dec.x; h.w = static_cast<const bf16*>(m->w_pix); h.M = Mf; h.N = dd->num_classes; h.K = dd->dim;
  h.epi = EPI_LN; h.bias = m->b_pix; h.colsum = m->cs_pix; h.out = ws.pix; h.ln_eps = eps;
  h.stat_parts = ws.dec.parts; h.n_stat_parts = gemm_stat_parts(Mf, dd->dim, false, nullptr);
  if ((rc = launch_gemm(h, stream))) return rc;
  if ((rc = launch_tail_rows_f32(ws.pix, pixels, B, N, N - n_vis, dd->num_classes, stream))) return rc;
  launches += 2;
  return launches;
}

int stad_eval_hist(const float* probs, const int32_t* labels, long long n, const float* thresholds, int T,
                   unsigned long long* hist, unsigned long long* conf, stad_stream_t stream) {
  STAD_CHECK_ARG(probs && labels && thresholds && hist && conf, "eval_hist: NULL argument");
  return launch_eval_hist(probs, labels, n, thresholds, T, hist, conf, as_stream(stream));
}

int stad_resize_cubic_u8(const void* frames_u8, void* out_u8, int F, int Hs, int Ws, int Hd, int Wd, const int32_t* xofs,
                         const int16_t* xw, const int32_t* yofs, const int16_t* yw, stad_stream_t stream) {
  STAD_CHECK_ARG(frames_u8 && out_u8 && xofs && xw && yofs && yw, "resize_cubic_u8: NULL argument");
  return launch_resize_cubic_u8(static_cast<const uint8_t*>(frames_u8), static_cast<uint8_t*>(out_u8), F, Hs, Ws, Hd, Wd,
                                xofs, xw, yofs, yw, as_stream(stream));
}

}  // extern "C"
