// HBM-bound row kernels: input cast, LayerNorm statistics, final LayerNorm, mean-pool + fc_norm + head, and the
// visible-token patch gather.  Each is a single pass over its input with 16-byte vector accesses; the roofline
// that bounds them is HBM bandwidth (algorithmic bytes are stated per kernel).
#include "kernels.h"
#include "ptx.cuh"

namespace stad {

namespace {

constexpr int kMaxChunks = 8;  // 8 x (32 lanes x 8 bf16) = D <= 2048
// one-warp-per-row kernels keep the row in registers: instantiate them for the number of 16-byte chunks per lane
#define STAD_DISPATCH_CHUNKS(D, launch)                  \
  do {                                                   \
    if ((D) <= 512) { constexpr int kC = 2; launch; }    \
    else if ((D) <= 768) { constexpr int kC = 3; launch; } \
    else if ((D) <= 1024) { constexpr int kC = 4; launch; } \
    else { constexpr int kC = kMaxChunks; launch; }      \
  } while (0)

STAD_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

STAD_DEVICE void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x);
  f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z);
  f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

// ---------------------------------------------------------------------------------------------------------------
// fp32 -> bf16.  Bytes: 4n read + 2n written.
__global__ void cast_kernel(const float* __restrict__ x, bf16* __restrict__ y, size_t n8, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  uint4* y4 = reinterpret_cast<uint4*>(y);
  // two groups of 8 elements per thread and iteration: four independent 16-byte loads in flight per thread
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + stride < n8; i += 2 * stride) {
    const float4 a = __ldg(x4 + 2 * i), b = __ldg(x4 + 2 * i + 1);
    const float4 c = __ldg(x4 + 2 * (i + stride)), d = __ldg(x4 + 2 * (i + stride) + 1);
    uint4 o, q;
    o.x = pack_bf16(a.x, a.y);
    o.y = pack_bf16(a.z, a.w);
    o.z = pack_bf16(b.x, b.y);
    o.w = pack_bf16(b.z, b.w);
    q.x = pack_bf16(c.x, c.y);
    q.y = pack_bf16(c.z, c.w);
    q.z = pack_bf16(d.x, d.y);
    q.w = pack_bf16(d.z, d.w);
    y4[i] = o;
    y4[i + stride] = q;
  }
  if (i < n8) {
    const float4 a = __ldg(x4 + 2 * i), b = __ldg(x4 + 2 * i + 1);
    uint4 o;
    o.x = pack_bf16(a.x, a.y);
    o.y = pack_bf16(a.z, a.w);
    o.z = pack_bf16(b.x, b.y);
    o.w = pack_bf16(b.z, b.w);
    y4[i] = o;
  }
  // tail (n not a multiple of 8)
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const size_t i = (n8 << 3) + threadIdx.x;
    y[i] = __float2bfloat16_rn(x[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// One warp per row; the row stays in registers between the mean pass and the variance pass (two-pass variance,
// like ATen's layer_norm, so no E[x^2]-mean^2 cancellation).  Bytes: 2 M D read (+ 8 M or 4 M D written).
// kChunks = 16-byte chunks per lane (D <= 256 kChunks): instantiated for the widths in use so that a D = 768 row costs
// 24 registers, not the 64 of the 2048-wide maximum (measured with the single instantiation: 117-128 registers, 20 %
// occupancy, 0.28-0.33 of the HBM copy rate).
template <bool kWriteNorm, int kChunks>
__global__ void __launch_bounds__(256)
row_norm_kernel(const bf16* __restrict__ x, float2* __restrict__ stats, const float* __restrict__ g,
                const float* __restrict__ b, float* __restrict__ y, int M, int D, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const int chunks = D >> 3;  // 16-byte chunks per row
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(warp) * D);
  float v[kChunks][8];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int idx = c * 32 + lane;
    if (idx < chunks) {
      unpack8(__ldg(xr + idx), v[c]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[c][j];
    }
  }
  const float mean = warp_sum(s) / static_cast<float>(D);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int idx = c * 32 + lane;
    if (idx < chunks) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[c][j] - mean;
        q = fmaf(d, d, q);
      }
    }
  }
  const float var = warp_sum(q) / static_cast<float>(D);
  const float rstd = rsqrtf(var + eps);
  if constexpr (!kWriteNorm) {
    if (lane == 0) stats[warp] = make_float2(mean, rstd);
  } else {
    float* yr = y + static_cast<size_t>(warp) * D;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int idx = c * 32 + lane;
      if (idx < chunks) {
        const int col = idx * 8;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + col));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(g + col + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + col));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(b + col + 4));
        float4 o0, o1;
        o0.x = fmaf((v[c][0] - mean) * rstd, g0.x, b0.x);
        o0.y = fmaf((v[c][1] - mean) * rstd, g0.y, b0.y);
        o0.z = fmaf((v[c][2] - mean) * rstd, g0.z, b0.z);
        o0.w = fmaf((v[c][3] - mean) * rstd, g0.w, b0.w);
        o1.x = fmaf((v[c][4] - mean) * rstd, g1.x, b1.x);
        o1.y = fmaf((v[c][5] - mean) * rstd, g1.y, b1.y);
        o1.z = fmaf((v[c][6] - mean) * rstd, g1.z, b1.z);
        o1.w = fmaf((v[c][7] - mean) * rstd, g1.w, b1.w);
        reinterpret_cast<float4*>(yr + col)[0] = o0;
        reinterpret_cast<float4*>(yr + col)[1] = o1;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// (sum, sumsq) partials [n_parts, M] written by the statistics epilogue of the residual GEMMs -> (mean, rstd) [M].
// One thread per row; partials are added in index order (deterministic).  Bytes: 8 M (n_parts + 1).
__global__ void __launch_bounds__(256)
stats_finalize_kernel(const float2* __restrict__ parts, int n_parts, float2* __restrict__ stats, int M, float inv_d,
                      float eps) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_wait();  // launched as a programmatic dependent of the GEMM that wrote `parts`
  if (row >= M) return;
  float s1 = 0.f, s2 = 0.f;
  // all loads of a row in flight together, eight at a time (2 * N / BN partials: 6 for the 256-wide tiles of a ViT-B
  // step, up to 32 with 64-wide tiles); the adds stay in index order.  The kernel moves a few MB: it is bound by its
  // launch and by this one round trip to L2, not by bandwidth.
  const float2* src = parts + row;
  for (int i = 0; i < n_parts; i += 8) {
    float2 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      v[k] = (i + k < n_parts) ? __ldg(&src[static_cast<size_t>(i + k) * M]) : make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (i + k < n_parts) {  // (adding the zeros of absent parts would turn a -0 sum into +0)
        s1 += v[k].x;
        s2 += v[k].y;
      }
    }
  }
  const float mean = s1 * inv_d;
  const float var = fmaxf(fmaf(-mean, mean, s2 * inv_d), 0.f);
  stats[row] = make_float2(mean, rsqrtf(var + eps));
}

// ---------------------------------------------------------------------------------------------------------------
// mean-pool stage 1: grid (kPoolChunks, B); block sums its slice of the tokens for every column.  Clip b starts at
// x + b * clip_stride elements (N * D, or (N + 1) * D with x advanced by one row when a class token leads every clip).
// Bytes: 2 B N D read + 4 B kPoolChunks D written.  Deterministic (no atomics).
constexpr int kPoolChunks = 16;

__global__ void __launch_bounds__(384)
pool_partial_kernel(const bf16* __restrict__ x, float* __restrict__ partial, int N, int D, size_t clip_stride) {
  extern __shared__ float red[];  // [groups][D]
  const int cpr = D >> 3;                 // 16-byte chunks per row
  const int groups = blockDim.x / cpr;    // row groups working in parallel
  const int grp = threadIdx.x / cpr;
  const int ch = threadIdx.x % cpr;
  const int b = blockIdx.y;
  const int per = (N + kPoolChunks - 1) / kPoolChunks;
  const int n0 = blockIdx.x * per;
  const int n1 = min(N, n0 + per);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (grp < groups) {
    const uint4* base = reinterpret_cast<const uint4*>(x + static_cast<size_t>(b) * clip_stride) + ch;
    // eight rows in flight per thread (independent 16-byte loads): the loop is latency-bound otherwise (measured with one
    // load per iteration: 0.31 of the HBM copy rate, with four: 0.61)
    int n = n0 + grp;
    for (; n + 7 * groups < n1; n += 8 * groups) {
      uint4 u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) u[k] = __ldg(base + static_cast<size_t>(n + k * groups) * cpr);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float f[8];
        unpack8(u[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    }
    for (; n < n1; n += groups) {
      float f[8];
      unpack8(__ldg(base + static_cast<size_t>(n) * cpr), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[grp * D + ch * 8 + j] = acc[j];
  }
  __syncthreads();
  float* out = partial + (static_cast<size_t>(b) * kPoolChunks + blockIdx.x) * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float s = 0.f;
    for (int gidx = 0; gidx < groups; ++gidx) s += red[gidx * D + c];
    out[c] = s;
  }
}

STAD_DEVICE float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = 0.f;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) t += sh[i];
  return t;
}

// stage 2: one block per clip: finish the mean, fc_norm (LayerNorm), head, optional softmax.
__global__ void __launch_bounds__(256)
pool_head_kernel(const float* __restrict__ partial, const float* __restrict__ g, const float* __restrict__ b,
                 const float* __restrict__ w_head, const float* __restrict__ b_head, float* __restrict__ logits,
                 float* __restrict__ probs, float* __restrict__ features, int N, int D, int C, float eps) {
  extern __shared__ float sm[];  // [D] pooled / normalised vector, then [32] scratch, then [C] logits
  float* vec = sm;
  float* sh = sm + D;
  float* lg = sh + 32;
  const int bidx = blockIdx.x;
  const float inv_n = 1.0f / static_cast<float>(N);
  float s = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float t = 0.f;
    for (int k = 0; k < kPoolChunks; ++k) t += partial[(static_cast<size_t>(bidx) * kPoolChunks + k) * D + c];
    t *= inv_n;
    vec[c] = t;
    s += t;
  }
  const float mean = block_sum(s, sh) / static_cast<float>(D);
  float q = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float d = vec[c] - mean;
    q = fmaf(d, d, q);
  }
  const float var = block_sum(q, sh) / static_cast<float>(D);
  const float rstd = rsqrtf(var + eps);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float y = fmaf((vec[c] - mean) * rstd, g[c], b[c]);
    vec[c] = y;
    if (features != nullptr) features[static_cast<size_t>(bidx) * D + c] = y;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int cls = warp; cls < C; cls += nw) {
    float d = 0.f;
    for (int c = lane; c < D; c += 32) d = fmaf(vec[c], w_head[static_cast<size_t>(cls) * D + c], d);
    d = warp_sum(d);
    if (lane == 0) {
      d += b_head[cls];
      lg[cls] = d;
      logits[static_cast<size_t>(bidx) * C + cls] = d;
    }
  }
  __syncthreads();
  if (probs != nullptr && threadIdx.x == 0) {
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, lg[c]);
    float den = 0.f;
    for (int c = 0; c < C; ++c) den += expf(lg[c] - mx);
    for (int c = 0; c < C; ++c) probs[static_cast<size_t>(bidx) * C + c] = expf(lg[c] - mx) / den;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// final_reduction 'cls' / 'none' (modeling_finetune.py:323-334): LayerNorm of selected rows of the residual stream,
// then the head and the softmax.  One warp per selected row r, which is row r * row_stride + row_off of x:
// 'cls' reads row 0 of every clip (row_stride = tokens per clip), 'none' every row (row_stride = 1).
// Bytes: 2 R D read + 4 R (D + 2 C) written (features optional).
__global__ void __launch_bounds__(256)
rows_norm_head_kernel(const bf16* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                      const float* __restrict__ w_head, const float* __restrict__ b_head, float* __restrict__ logits,
                      float* __restrict__ probs, float* __restrict__ features, int R, long long row_stride,
                      long long row_off, int D, int C, float eps) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const int chunks = D >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(r * row_stride + row_off) * D);
  float v[kMaxChunks][8];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int idx = c * 32 + lane;
    if (idx < chunks) {
      unpack8(__ldg(xr + idx), v[c]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[c][j];
    }
  }
  const float mean = warp_sum(s) / static_cast<float>(D);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int idx = c * 32 + lane;
    if (idx < chunks) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[c][j] - mean;
        q = fmaf(d, d, q);
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(D) + eps);
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int idx = c * 32 + lane;
    if (idx < chunks) {
      const int col = idx * 8;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + col));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(g + col + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + col));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(b + col + 4));
      v[c][0] = fmaf((v[c][0] - mean) * rstd, g0.x, b0.x);
      v[c][1] = fmaf((v[c][1] - mean) * rstd, g0.y, b0.y);
      v[c][2] = fmaf((v[c][2] - mean) * rstd, g0.z, b0.z);
      v[c][3] = fmaf((v[c][3] - mean) * rstd, g0.w, b0.w);
      v[c][4] = fmaf((v[c][4] - mean) * rstd, g1.x, b1.x);
      v[c][5] = fmaf((v[c][5] - mean) * rstd, g1.y, b1.y);
      v[c][6] = fmaf((v[c][6] - mean) * rstd, g1.z, b1.z);
      v[c][7] = fmaf((v[c][7] - mean) * rstd, g1.w, b1.w);
      if (features != nullptr) {
        float4* fr = reinterpret_cast<float4*>(features + static_cast<size_t>(r) * D + col);
        fr[0] = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
        fr[1] = make_float4(v[c][4], v[c][5], v[c][6], v[c][7]);
      }
    }
  }
  if (w_head == nullptr) return;
  float* lr = logits + static_cast<size_t>(r) * C;
  float mx = -INFINITY;
  for (int cls = 0; cls < C; ++cls) {
    const float* wr = w_head + static_cast<size_t>(cls) * D;
    float d = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int idx = c * 32 + lane;
      if (idx < chunks) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wr + idx * 8));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wr + idx * 8 + 4));
        d = fmaf(v[c][0], w0.x, d); d = fmaf(v[c][1], w0.y, d); d = fmaf(v[c][2], w0.z, d); d = fmaf(v[c][3], w0.w, d);
        d = fmaf(v[c][4], w1.x, d); d = fmaf(v[c][5], w1.y, d); d = fmaf(v[c][6], w1.z, d); d = fmaf(v[c][7], w1.w, d);
      }
    }
    d = warp_sum(d) + __ldg(&b_head[cls]);
    mx = fmaxf(mx, d);
    if (lane == 0) lr[cls] = d;
  }
  if (probs == nullptr) return;
  __syncwarp();  // lane 0's logits are visible to the warp
  float den = 0.f;
  for (int cls = lane; cls < C; cls += 32) den += expf(lr[cls] - mx);
  den = warp_sum(den);
  for (int cls = lane; cls < C; cls += 32) probs[static_cast<size_t>(r) * C + cls] = expf(lr[cls] - mx) / den;
}

// ---------------------------------------------------------------------------------------------------------------
// Class token of the MVD sibling (other_models/MVD/modeling_finetune.py:431-435): x[b] = cat(cls_token, emb[b]) — the
// token carries no position row — plus the LayerNorm statistics of every row of x for norm1 of the first block.
// One warp per output row.  Bytes: 2 B N D read + 2 B (N + 1) D written (+ 8 B (N + 1)).
__global__ void __launch_bounds__(256)
prepend_cls_kernel(const bf16* __restrict__ emb, const float* __restrict__ cls_token, bf16* __restrict__ x,
                   float2* __restrict__ stats, int B, int N, int D, float eps) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int S = N + 1;
  if (row >= B * S) return;
  const int b = row / S;
  const int i = row - b * S;
  const int chunks = D >> 3;
  uint4* xr = reinterpret_cast<uint4*>(x + static_cast<size_t>(row) * D);
  const uint4* er = reinterpret_cast<const uint4*>(emb + (static_cast<size_t>(b) * N + (i > 0 ? i - 1 : 0)) * D);
  float v[kMaxChunks][8];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int idx = c * 32 + lane;
    if (idx < chunks) {
      uint4 u;
      if (i > 0) {
        u = __ldg(er + idx);
      } else {
        const float4 c0 = __ldg(reinterpret_cast<const float4*>(cls_token + idx * 8));
        const float4 c1 = __ldg(reinterpret_cast<const float4*>(cls_token + idx * 8 + 4));
        u.x = pack_bf16(c0.x, c0.y);
        u.y = pack_bf16(c0.z, c0.w);
        u.z = pack_bf16(c1.x, c1.y);
        u.w = pack_bf16(c1.z, c1.w);
      }
      xr[idx] = u;
      unpack8(u, v[c]);  // statistics of the values as stored (bf16)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[c][j];
    }
  }
  const float mean = warp_sum(s) / static_cast<float>(D);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int idx = c * 32 + lane;
    if (idx < chunks) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[c][j] - mean;
        q = fmaf(d, d, q);
      }
    }
  }
  const float var = warp_sum(q) / static_cast<float>(D);
  if (lane == 0) stats[row] = make_float2(mean, rsqrtf(var + eps));
}

// ---------------------------------------------------------------------------------------------------------------
// visible-token im2col (masked encoder, modeling_pretrain.py:93-98 restricted to the tokens that survive):
// one thread moves 8 contiguous dw pixels (16 bytes).
__global__ void __launch_bounds__(256)
gather_patches_kernel(const bf16* __restrict__ planes, PatchGeom pg, const int32_t* __restrict__ tok_idx,
                      bf16* __restrict__ out, int B, int n_tok, int K) {
  const int k8 = K >> 3;
  const size_t total = static_cast<size_t>(B) * n_tok * k8;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int kc = static_cast<int>(i % k8);
  const size_t row = i / k8;
  const int b = static_cast<int>(row / n_tok);
  const int tok = __ldg(&tok_idx[row]);
  const int wp = tok % pg.Wp;
  const int hp = (tok / pg.Wp) % pg.Hp;
  const int tp = tok / (pg.Wp * pg.Hp);
  const int k = kc * 8;  // k = ((c*tubelet + dt)*16 + dh)*16 + dw
  const int dw = k & 15;
  const int dh = (k >> 4) & 15;
  const int dt = (k >> 8) % pg.tubelet;
  const int c = (k >> 8) / pg.tubelet;
  const int t = tp * pg.tubelet + dt;
  const int f0 = pg.win_start != nullptr ? __ldg(pg.win_start + b) : pg.start + b * pg.stride;
  const int plane = pg.mode == STAD_IN_CLIPS ? (b * pg.C + c) * pg.T + t : (f0 + t * pg.fstep) * pg.C + c;
  const size_t src = (static_cast<size_t>(plane) * pg.img_h + hp * 16 + dh) * pg.img_w + wp * 16 + dw;
  // a caller-supplied first frame beyond the buffer reads as zeros, like the tensor-map path (never out of bounds)
  reinterpret_cast<uint4*>(out)[i] = (plane >= 0 && plane < pg.n_planes) ? __ldg(reinterpret_cast<const uint4*>(planes + src))
                                                                         : make_uint4(0u, 0u, 0u, 0u);
}

// ---------------------------------------------------------------------------------------------------------------
// MAE decoder input (modeling_pretrain.py:283-288): x_full[b] = cat(x_vis[b] (+ pos, already added by the
// encoder_to_decoder GEMM epilogue), mask_token + pos[masked ids of clip b]) and, in the same pass, the LayerNorm
// statistics of every row of x_full for the first decoder block's norm1.  One warp per output row.
// Bytes: 2 B n_vis D read + 2 B N D written (+ 8 B N); the position table [N, D] fp32 stays in L2.
template <int kChunks>
__global__ void __launch_bounds__(256)
decoder_assemble_kernel(const bf16* __restrict__ vis, const float* __restrict__ pos,
                        const float* __restrict__ mask_token, const int32_t* __restrict__ mask_idx,
                        bf16* __restrict__ x, float2* __restrict__ stats, int B, int N, int n_vis, int D, float eps) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B * N) return;
  const int b = row / N;
  const int i = row - b * N;
  const int chunks = D >> 3;
  uint4* xr = reinterpret_cast<uint4*>(x + static_cast<size_t>(row) * D);
  float v[kChunks][8];
  float s = 0.f;
  const bool visible = i < n_vis;
  const uint4* vr = reinterpret_cast<const uint4*>(vis + (static_cast<size_t>(b) * n_vis + (visible ? i : 0)) * D);
  const float* pr =
      pos + static_cast<size_t>(visible ? 0 : __ldg(&mask_idx[static_cast<size_t>(b) * (N - n_vis) + (i - n_vis)])) * D;
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int idx = c * 32 + lane;
    if (idx < chunks) {
      uint4 u;
      if (visible) {
        u = __ldg(vr + idx);
      } else {
        const int col = idx * 8;
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(pr + col));
        const float4 p1 = __ldg(reinterpret_cast<const float4*>(pr + col + 4));
        const float4 m0 = __ldg(reinterpret_cast<const float4*>(mask_token + col));
        const float4 m1 = __ldg(reinterpret_cast<const float4*>(mask_token + col + 4));
        u.x = pack_bf16(m0.x + p0.x, m0.y + p0.y);
        u.y = pack_bf16(m0.z + p0.z, m0.w + p0.w);
        u.z = pack_bf16(m1.x + p1.x, m1.y + p1.y);
        u.w = pack_bf16(m1.z + p1.z, m1.w + p1.w);
      }
      xr[idx] = u;
      unpack8(u, v[c]);  // statistics of the values as stored (bf16)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[c][j];
    }
  }
  const float mean = warp_sum(s) / static_cast<float>(D);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int idx = c * 32 + lane;
    if (idx < chunks) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[c][j] - mean;
        q = fmaf(d, d, q);
      }
    }
  }
  const float var = warp_sum(q) / static_cast<float>(D);
  if (lane == 0) stats[row] = make_float2(mean, rsqrtf(var + eps));
}

// Tubelet-embedding reuse across overlapping sliding windows (run_inference.py:97-101 shifts the window by one frame:
// 15 of its 16 frames, i.e. 7 of the 8 tubelets of every second window, are shared).  The tubelet embeddings E[u] =
// Conv3d(frames of tubelet u) (no bias, no position) are computed ONCE per distinct tubelet by the patch-embed GEMM;
// this kernel lays out the residual stream of every window from them:
//   x[b, t' * HW + hw, :] = bf16(E[(b + t' * step) * HW + hw, :] + pos_bias[t' * HW + hw, :])      (mf:309-313)
// and, in the same pass, the LayerNorm statistics (mean, rstd) of every row for norm1 of the first block.
// One warp per output row.  Bytes: 2 B N D written (+ 8 B N); E (2 n_u HW D) and pos_bias (4 N D) are re-read from L2.
template <int kChunks>
__global__ void __launch_bounds__(256)
window_assemble_kernel(const bf16* __restrict__ emb, const float* __restrict__ pos_bias, bf16* __restrict__ x,
                       float2* __restrict__ stats, int B, int Tp, int HW, int D, int step, float eps) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int N = Tp * HW;
  if (row >= B * N) return;
  const int b = row / N;
  const int tok = row - b * N;
  const int tp = tok / HW;
  const int hw = tok - tp * HW;
  const int chunks = D >> 3;
  const uint4* er = reinterpret_cast<const uint4*>(emb + (static_cast<size_t>(b + tp * step) * HW + hw) * D);
  const float* pr = pos_bias + static_cast<size_t>(tok) * D;
  uint4* xr = reinterpret_cast<uint4*>(x + static_cast<size_t>(row) * D);
  float v[kChunks][8];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int idx = c * 32 + lane;
    if (idx < chunks) {
      float e[8];
      unpack8(__ldg(er + idx), e);
      const float4 p0 = __ldg(reinterpret_cast<const float4*>(pr + idx * 8));
      const float4 p1 = __ldg(reinterpret_cast<const float4*>(pr + idx * 8 + 4));
      uint4 u;
      u.x = pack_bf16(e[0] + p0.x, e[1] + p0.y);
      u.y = pack_bf16(e[2] + p0.z, e[3] + p0.w);
      u.z = pack_bf16(e[4] + p1.x, e[5] + p1.y);
      u.w = pack_bf16(e[6] + p1.z, e[7] + p1.w);
      xr[idx] = u;
      unpack8(u, v[c]);  // statistics of the values as stored (bf16)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[c][j];
    }
  }
  const float mean = warp_sum(s) / static_cast<float>(D);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int idx = c * 32 + lane;
    if (idx < chunks) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dlt = v[c][j] - mean;
        q = fmaf(dlt, dlt, q);
      }
    }
  }
  const float var = warp_sum(q) / static_cast<float>(D);
  if (lane == 0) stats[row] = make_float2(mean, rsqrtf(var + eps));
}

// Last n_keep rows of every clip of x[B, N, C] bf16 -> y[B, n_keep, C] fp32 (the decoder returns only the predictions
// of the masked tokens, modeling_pretrain.py:174).  Bytes: 2 B n_keep C read + 4 B n_keep C written.
__global__ void __launch_bounds__(256)
tail_rows_f32_kernel(const bf16* __restrict__ x, float* __restrict__ y, int N, int n_keep, int c8, size_t total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ch = static_cast<int>(i % c8);
  const size_t r = i / c8;
  const size_t b = r / n_keep;
  const int j = static_cast<int>(r - b * n_keep);
  const size_t src = (b * N + (N - n_keep) + j) * static_cast<size_t>(c8) + ch;
  float f[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(x) + src), f);
  float4* o = reinterpret_cast<float4*>(y) + 2 * i;
  o[0] = make_float4(f[0], f[1], f[2], f[3]);
  o[1] = make_float4(f[4], f[5], f[6], f[7]);
}

// ---------------------------------------------------------------------------------------------------------------
// Frame preparation (run_inference.py:15-34 prepare_image; dota.py:347-348 + volume_transforms ClipToTensor/normalize):
// uint8 HWC frames (BGR as cv2.imread returns them, or RGB) -> bf16 planes [F, 3, H, W] = (v / 255 - mean[c]) / std[c]
// with c the RGB channel.  One thread converts 8 consecutive pixels of one row (24 input bytes, three 16-byte stores).
// Bytes: 3 F H W read + 6 F H W written.
struct NormArgs {
  float scale[3];  // 1 / (255 std[c])
  float shift[3];  // -mean[c] / std[c]
};
__global__ void __launch_bounds__(256)
normalize_u8_kernel(const uint8_t* __restrict__ in, bf16* __restrict__ out, int HW, int hw8, size_t total, NormArgs na,
                    int bgr) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t f = i / hw8;
  const int p8 = static_cast<int>(i - f * hw8);
  // 24 bytes = 8 pixels x 3 interleaved channels; 8-byte aligned because HW % 8 == 0
  const uint2* src = reinterpret_cast<const uint2*>(in + (f * HW + static_cast<size_t>(p8) * 8) * 3);
  const uint2 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
  const uint32_t w[6] = {a.x, a.y, b.x, b.y, c.x, c.y};
  float px[3][8];
#pragma unroll
  for (int k = 0; k < 24; ++k) {
    const float val = static_cast<float>((w[k >> 2] >> ((k & 3) * 8)) & 0xffu);
    px[k % 3][k / 3] = val;
  }
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const int srcc = bgr ? 2 - ch : ch;  // output plane ch (R, G, B) reads interleaved channel srcc
    uint4 o;
    const float sc = na.scale[ch], sh = na.shift[ch];
    o.x = pack_bf16(fmaf(px[srcc][0], sc, sh), fmaf(px[srcc][1], sc, sh));
    o.y = pack_bf16(fmaf(px[srcc][2], sc, sh), fmaf(px[srcc][3], sc, sh));
    o.z = pack_bf16(fmaf(px[srcc][4], sc, sh), fmaf(px[srcc][5], sc, sh));
    o.w = pack_bf16(fmaf(px[srcc][6], sc, sh), fmaf(px[srcc][7], sc, sh));
    reinterpret_cast<uint4*>(out + (f * 3 + ch) * HW)[p8] = o;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Evaluation epilogue (engine_for_frame_finetuning.py:461-488, anaysis/metrics.py:183-199): confusion counts of the
// per-frame risk probability against T ascending thresholds, one pass over the n scores.
// For a sample with probability p the number b of thresholds with t_k <= p (found by binary search with exactly the
// fp32 comparisons `p >= t_k` the reference makes) puts it into bin b of its label's histogram; the prediction at
// threshold k is positive iff b > k, so TP_k / FP_k are suffix sums of the two histograms.  Integer arithmetic
// throughout: the counts are exact and independent of the summation order (integer atomics).
// hist: [2][T + 1] (label, bin); conf: [4] = tn, fp, fn, tp of the arg-max prediction (eff:464, ties -> class 0).
// Bytes: 12 n read.
constexpr int kMaxThresholds = 1024;
__global__ void __launch_bounds__(256)
eval_hist_kernel(const float* __restrict__ probs, const int32_t* __restrict__ labels, long long n,
                 const float* __restrict__ thresholds, int T, unsigned long long* __restrict__ hist,
                 unsigned long long* __restrict__ conf) {
  extern __shared__ unsigned int sh[];  // [T] thresholds as float bits, then [2][T + 1] counts, then [4] arg-max conf
  float* th = reinterpret_cast<float*>(sh);
  unsigned int* cnt = sh + T;
  unsigned int* cf = cnt + 2 * (T + 1);
  for (int i = threadIdx.x; i < T; i += blockDim.x) th[i] = thresholds[i];
  for (int i = threadIdx.x; i < 2 * (T + 1) + 4; i += blockDim.x) cnt[i] = 0u;
  __syncthreads();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float2 p = __ldg(reinterpret_cast<const float2*>(probs) + i);
    const int y = __ldg(labels + i) != 0;
    int lo = 0, hi = T;  // b = #{k : th[k] <= p.y}; thresholds ascending
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (p.y >= th[mid]) lo = mid + 1;
      else hi = mid;
    }
    atomicAdd(&cnt[y * (T + 1) + lo], 1u);
    const int pred = p.y > p.x;  // torch.max(softmax, 1): first maximum wins, so a tie predicts class 0
    atomicAdd(&cf[y * 2 + pred], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * (T + 1); i += blockDim.x)
    if (cnt[i]) atomicAdd(&hist[i], static_cast<unsigned long long>(cnt[i]));
  if (threadIdx.x < 4 && cf[threadIdx.x]) atomicAdd(&conf[threadIdx.x], static_cast<unsigned long long>(cf[threadIdx.x]));
}

// ---------------------------------------------------------------------------------------------------------------
// Bicubic frame resize, uint8 HWC -> uint8 HWC: cv2.resize(img, (W_d, H_d), interpolation=cv2.INTER_CUBIC) as the
// reference's callers invoke it (run_inference.py:79-80, dota.py:347-348), in OpenCV's 8-bit fixed-point arithmetic:
// 11-bit tap weights (tables built on the host exactly as resize.cpp builds them), source taps clamped at the borders,
// horizontal pass in int32, vertical pass, (v + 2^21) >> 22, saturation.  Integer arithmetic: bit-exact against the
// oracle.  One thread per destination pixel (3 channels); block = one destination row segment.
// Bytes: ~3 F H_s W_s read (every source pixel is touched; re-reads hit L1/L2) + 3 F H_d W_d written.
__global__ void __launch_bounds__(256)
resize_cubic_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int Hs, int Ws, int Hd, int Wd,
                       const int32_t* __restrict__ xofs, const int16_t* __restrict__ xw,
                       const int32_t* __restrict__ yofs, const int16_t* __restrict__ yw) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x;
  const int dy = blockIdx.y;
  const int f = blockIdx.z;
  if (dx >= Wd) return;
  const int sx = __ldg(&xofs[dx]);
  const int sy = __ldg(&yofs[dy]);
  int wx[4], wy[4], cx[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    wx[k] = __ldg(&xw[dx * 4 + k]);
    wy[k] = __ldg(&yw[dy * 4 + k]);
    cx[k] = min(max(sx - 1 + k, 0), Ws - 1) * 3;
  }
  const uint8_t* src = in + static_cast<size_t>(f) * Hs * Ws * 3;
  int acc[3] = {0, 0, 0};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint8_t* row = src + static_cast<size_t>(min(max(sy - 1 + j, 0), Hs - 1)) * Ws * 3;
    int h[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int c = 0; c < 3; ++c) h[c] += static_cast<int>(__ldg(row + cx[k] + c)) * wx[k];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += h[c] * wy[j];
  }
  uint8_t* dst = out + ((static_cast<size_t>(f) * Hd + dy) * Wd + dx) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) dst[c] = static_cast<uint8_t>(min(max((acc[c] + (1 << 21)) >> 22, 0), 255));
}

}  // namespace

int launch_cast_f32_bf16(const float* x, bf16* y, size_t n, cudaStream_t stream) {
  if (n == 0) return STAD_OK;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15)
    return fail(STAD_E_ALIGN, "cast: pointers must be 16-byte aligned");
  const size_t n8 = n >> 3;
  const int threads = 256;
  size_t blocks = (n8 + threads - 1) / threads;
  const size_t cap = static_cast<size_t>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  ProfScope prof(STAD_K_CAST, 0, static_cast<int>(n >> 10), 1024, 0, stream);
  cast_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(x, y, n8, n);
  STAD_LAUNCH_OK("cast_kernel");
  return STAD_OK;
}

static int check_row_args(const void* x, int M, int D) {
  STAD_CHECK_ARG(M > 0 && D > 0, "row kernel: empty input M=%d D=%d", M, D);
  STAD_CHECK_ARG(D % 8 == 0 && D <= kMaxChunks * 256, "row kernel: D=%d must be a multiple of 8 and <= %d", D,
                 kMaxChunks * 256);
  if (reinterpret_cast<uintptr_t>(x) & 15) return fail(STAD_E_ALIGN, "row kernel: x must be 16-byte aligned");
  return STAD_OK;
}

int launch_stats_finalize(const float2* parts, int n_parts, float2* stats, int M, int D, float eps, cudaStream_t stream) {
  STAD_CHECK_ARG(M > 0 && D > 0 && n_parts >= 1 && n_parts <= kMaxStatParts, "stats_finalize: M=%d D=%d parts=%d", M, D,
                 n_parts);
  if ((reinterpret_cast<uintptr_t>(parts) | reinterpret_cast<uintptr_t>(stats)) & 7)
    return fail(STAD_E_ALIGN, "stats_finalize: buffers must be 8-byte aligned");
  ProfScope prof(STAD_K_ROW_STATS, 1, M, D, n_parts, stream);
  STAD_CUDA_OK(launch_pdl(stats_finalize_kernel, dim3(ceil_div(M, 256)), dim3(256), 0, stream, 1, parts, n_parts, stats, M,
                          1.0f / static_cast<float>(D), eps));
  STAD_LAUNCH_OK("stats_finalize");
  return STAD_OK;
}

int launch_row_stats(const bf16* x, float2* stats, int M, int D, float eps, cudaStream_t stream) {
  int rc = check_row_args(x, M, D);
  if (rc) return rc;
  const int rows_per_block = 8;
  ProfScope prof(STAD_K_ROW_STATS, 0, M, D, 0, stream);
  STAD_DISPATCH_CHUNKS(D, (row_norm_kernel<false, kC><<<ceil_div(M, rows_per_block), rows_per_block * 32, 0, stream>>>(
                              x, stats, nullptr, nullptr, nullptr, M, D, eps)));
  STAD_LAUNCH_OK("row_stats");
  return STAD_OK;
}

int launch_layernorm(const bf16* x, const float* g, const float* b, float* y, int M, int D, float eps,
                     cudaStream_t stream) {
  int rc = check_row_args(x, M, D);
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) & 15)
    return fail(STAD_E_ALIGN, "layernorm: g, b, y must be 16-byte aligned");
  const int rows_per_block = 8;
  ProfScope prof(STAD_K_LAYERNORM, 0, M, D, 0, stream);
  STAD_DISPATCH_CHUNKS(D, (row_norm_kernel<true, kC><<<ceil_div(M, rows_per_block), rows_per_block * 32, 0, stream>>>(
                              x, nullptr, g, b, y, M, D, eps)));
  STAD_LAUNCH_OK("layernorm");
  return STAD_OK;
}

int launch_pool_norm_head(const bf16* x, const float* g, const float* b, const float* w_head, const float* b_head,
                          float* logits, float* probs, float* features, float* scratch, int B, int N, int D, int C,
                          float eps, cudaStream_t stream, size_t clip_stride) {
  if (clip_stride == 0) clip_stride = static_cast<size_t>(N) * D;
  STAD_CHECK_ARG(B > 0 && N > 0 && C > 0 && C <= 1024, "pool_norm_head: bad sizes B=%d N=%d C=%d", B, N, C);
  STAD_CHECK_ARG(D % 8 == 0 && D >= 8 && D <= 2048, "pool_norm_head: D=%d must be a multiple of 8 and <= 2048", D);
  if (reinterpret_cast<uintptr_t>(x) & 15) return fail(STAD_E_ALIGN, "pool_norm_head: x must be 16-byte aligned");
  const int threads = 256;
  // pooling block: as many whole row groups (D / 8 threads each) as fit 384 threads (D = 768: 4 groups, no idle threads)
  const int groups = 384 / (D >> 3);
  STAD_CHECK_ARG(groups >= 1, "pool_norm_head: D too large for the pooling block");
  const int threads1 = groups * (D >> 3);
  const size_t smem1 = static_cast<size_t>(groups) * D * sizeof(float);
  ProfScope prof(STAD_K_POOL, 0, B * N, D, 0, stream);
  pool_partial_kernel<<<dim3(kPoolChunks, B), threads1, smem1, stream>>>(x, scratch, N, D, clip_stride);
  STAD_LAUNCH_OK("pool_partial");
  const size_t smem2 = (static_cast<size_t>(D) + 32 + C) * sizeof(float);
  pool_head_kernel<<<B, threads, smem2, stream>>>(scratch, g, b, w_head, b_head, logits, probs, features, N, D, C, eps);
  STAD_LAUNCH_OK("pool_head");
  return STAD_OK;
}

int launch_rows_norm_head(const bf16* x, const float* g, const float* b, const float* w_head, const float* b_head,
                          float* logits, float* probs, float* features, int R, long long row_stride, long long row_off,
                          int D, int C, float eps, cudaStream_t stream) {
  int rc = check_row_args(x, R, D);
  if (rc) return rc;
  STAD_CHECK_ARG(row_stride >= 1 && row_off >= 0, "rows_norm_head: row_stride=%lld row_off=%lld", row_stride, row_off);
  STAD_CHECK_ARG(w_head == nullptr || (b_head && logits && C > 0), "rows_norm_head: a head needs b_head, logits, C > 0");
  STAD_CHECK_ARG(w_head != nullptr || features != nullptr, "rows_norm_head: nothing to write");
  if ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(features) |
       reinterpret_cast<uintptr_t>(w_head)) & 15)
    return fail(STAD_E_ALIGN, "rows_norm_head: g, b, w_head, features must be 16-byte aligned");
  const int rows_per_block = 8;
  ProfScope prof(STAD_K_LAYERNORM, 1, R, D, C, stream);
  rows_norm_head_kernel<<<ceil_div(R, rows_per_block), rows_per_block * 32, 0, stream>>>(
      x, g, b, w_head, b_head, logits, probs, features, R, row_stride, row_off, D, C, eps);
  STAD_LAUNCH_OK("rows_norm_head");
  return STAD_OK;
}

int launch_prepend_cls(const bf16* emb, const float* cls_token, bf16* x, float2* stats, int B, int N, int D, float eps,
                       cudaStream_t stream) {
  STAD_CHECK_ARG(B > 0 && N > 0, "prepend_cls: B=%d N=%d", B, N);
  int rc = check_row_args(x, B * (N + 1), D);
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(emb) | reinterpret_cast<uintptr_t>(cls_token)) & 15)
    return fail(STAD_E_ALIGN, "prepend_cls: emb, cls_token must be 16-byte aligned");
  const int rows_per_block = 8;
  ProfScope prof(STAD_K_ASSEMBLE, 1, B * (N + 1), D, 0, stream);
  prepend_cls_kernel<<<ceil_div(B * (N + 1), rows_per_block), rows_per_block * 32, 0, stream>>>(emb, cls_token, x, stats, B,
                                                                                             N, D, eps);
  STAD_LAUNCH_OK("prepend_cls");
  return STAD_OK;
}

int launch_gather_patches(const bf16* planes, const PatchGeom& pg, const int32_t* tok_idx, bf16* out, int B,
                          int n_tok, cudaStream_t stream) {
  const int K = pg.C * pg.tubelet * 256;
  if ((reinterpret_cast<uintptr_t>(planes) | reinterpret_cast<uintptr_t>(out)) & 15)
    return fail(STAD_E_ALIGN, "gather_patches: pointers must be 16-byte aligned");
  const size_t total = static_cast<size_t>(B) * n_tok * (K >> 3);
  const int threads = 256;
  ProfScope prof(STAD_K_GATHER, 0, B * n_tok, K, 0, stream);
  gather_patches_kernel<<<static_cast<unsigned>((total + threads - 1) / threads), threads, 0, stream>>>(
      planes, pg, tok_idx, out, B, n_tok, K);
  STAD_LAUNCH_OK("gather_patches");
  return STAD_OK;
}

int launch_window_assemble(const bf16* emb, const float* pos_bias, bf16* x, float2* stats, int B, int Tp, int HW, int D,
                           int step, float eps, cudaStream_t stream) {
  int rc = check_row_args(x, B * Tp * HW, D);
  if (rc) return rc;
  STAD_CHECK_ARG(step >= 0 && Tp >= 1 && HW >= 1, "window_assemble: Tp=%d HW=%d step=%d", Tp, HW, step);
  if ((reinterpret_cast<uintptr_t>(emb) | reinterpret_cast<uintptr_t>(pos_bias)) & 15)
    return fail(STAD_E_ALIGN, "window_assemble: emb, pos_bias must be 16-byte aligned");
  const int rows_per_block = 8;
  ProfScope prof(STAD_K_ASSEMBLE, 2, B * Tp * HW, D, step, stream);
  STAD_DISPATCH_CHUNKS(D, (window_assemble_kernel<kC><<<ceil_div(B * Tp * HW, rows_per_block), rows_per_block * 32, 0, stream>>>(
                              emb, pos_bias, x, stats, B, Tp, HW, D, step, eps)));
  STAD_LAUNCH_OK("window_assemble");
  return STAD_OK;
}

int launch_decoder_assemble(const bf16* vis, const float* pos, const float* mask_token, const int32_t* mask_idx, bf16* x,
                            float2* stats, int B, int N, int n_vis, int D, float eps, cudaStream_t stream) {
  int rc = check_row_args(x, B * N, D);
  if (rc) return rc;
  STAD_CHECK_ARG(n_vis > 0 && n_vis < N, "decoder_assemble: n_vis=%d must be in (0, %d)", n_vis, N);
  if ((reinterpret_cast<uintptr_t>(vis) | reinterpret_cast<uintptr_t>(pos) | reinterpret_cast<uintptr_t>(mask_token)) & 15)
    return fail(STAD_E_ALIGN, "decoder_assemble: vis, pos, mask_token must be 16-byte aligned");
  const int rows_per_block = 8;
  ProfScope prof(STAD_K_ASSEMBLE, 0, B * N, D, n_vis, stream);
  STAD_DISPATCH_CHUNKS(D, (decoder_assemble_kernel<kC><<<ceil_div(B * N, rows_per_block), rows_per_block * 32, 0, stream>>>(
                              vis, pos, mask_token, mask_idx, x, stats, B, N, n_vis, D, eps)));
  STAD_LAUNCH_OK("decoder_assemble");
  return STAD_OK;
}

int launch_tail_rows_f32(const bf16* x, float* y, int B, int N, int n_keep, int C, cudaStream_t stream) {
  STAD_CHECK_ARG(B > 0 && n_keep > 0 && n_keep <= N && C > 0 && C % 8 == 0, "tail_rows: B=%d N=%d n_keep=%d C=%d", B, N,
                 n_keep, C);
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15)
    return fail(STAD_E_ALIGN, "tail_rows: pointers must be 16-byte aligned");
  const size_t total = static_cast<size_t>(B) * n_keep * (C >> 3);
  const int threads = 256;
  ProfScope prof(STAD_K_TAIL, 0, B * n_keep, C, 0, stream);
  tail_rows_f32_kernel<<<static_cast<unsigned>((total + threads - 1) / threads), threads, 0, stream>>>(x, y, N, n_keep,
                                                                                                      C >> 3, total);
  STAD_LAUNCH_OK("tail_rows_f32");
  return STAD_OK;
}

int launch_normalize_u8(const uint8_t* in, bf16* out, int F, int H, int W, const float* mean, const float* std_, int bgr,
                        cudaStream_t stream) {
  STAD_CHECK_ARG(F > 0 && H > 0 && W > 0 && (static_cast<long long>(H) * W) % 8 == 0,
                 "normalize_u8: F=%d H=%d W=%d (H*W must be a multiple of 8)", F, H, W);
  if ((reinterpret_cast<uintptr_t>(in) & 7) | (reinterpret_cast<uintptr_t>(out) & 15))
    return fail(STAD_E_ALIGN, "normalize_u8: in must be 8-byte and out 16-byte aligned");
  NormArgs na;
  for (int c = 0; c < 3; ++c) {
    STAD_CHECK_ARG(std_[c] > 0.f, "normalize_u8: std[%d] must be positive", c);
    na.scale[c] = 1.0f / (255.0f * std_[c]);
    na.shift[c] = -mean[c] / std_[c];
  }
  const int HW = H * W;
  const size_t total = static_cast<size_t>(F) * (HW >> 3);
  const int threads = 256;
  ProfScope prof(STAD_K_NORMALIZE, 0, F, HW, 0, stream);
  normalize_u8_kernel<<<static_cast<unsigned>((total + threads - 1) / threads), threads, 0, stream>>>(
      in, out, HW, HW >> 3, total, na, bgr);
  STAD_LAUNCH_OK("normalize_u8");
  return STAD_OK;
}

int launch_eval_hist(const float* probs, const int32_t* labels, long long n, const float* thresholds, int T,
                     unsigned long long* hist, unsigned long long* conf, cudaStream_t stream) {
  STAD_CHECK_ARG(n > 0 && T >= 1 && T <= kMaxThresholds, "eval_hist: n=%lld T=%d (T <= %d)", n, T, kMaxThresholds);
  if ((reinterpret_cast<uintptr_t>(probs) | reinterpret_cast<uintptr_t>(hist) | reinterpret_cast<uintptr_t>(conf)) & 7)
    return fail(STAD_E_ALIGN, "eval_hist: probs, hist, conf must be 8-byte aligned");
  const int threads = 256;
  long long blocks = (n + threads * 8 - 1) / (threads * 8);
  const long long cap = static_cast<long long>(sm_count()) * 4;
  if (blocks > cap) blocks = cap;
  const size_t smem = (static_cast<size_t>(T) + 2 * (T + 1) + 4) * sizeof(unsigned int);
  STAD_CUDA_OK(cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * 2 * (T + 1), stream));
  STAD_CUDA_OK(cudaMemsetAsync(conf, 0, sizeof(unsigned long long) * 4, stream));
  ProfScope prof(STAD_K_EVAL, 0, static_cast<int>(n >> 10), T, 0, stream);
  eval_hist_kernel<<<static_cast<unsigned>(blocks), threads, smem, stream>>>(probs, labels, n, thresholds, T, hist, conf);
  STAD_LAUNCH_OK("eval_hist");
  return STAD_OK;
}

int launch_resize_cubic_u8(const uint8_t* in, uint8_t* out, int F, int Hs, int Ws, int Hd, int Wd, const int32_t* xofs,
                           const int16_t* xw, const int32_t* yofs, const int16_t* yw, cudaStream_t stream) {
  STAD_CHECK_ARG(F > 0 && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0 && Hd <= 65535 && F <= 65535,
                 "resize_cubic: F=%d src %dx%d dst %dx%d", F, Hs, Ws, Hd, Wd);
  // |acc| <= 255 * (sum |w_x|) * (sum |w_y|) < 255 * 2^12.4 * 2^12.4: fits int32
  const int threads = Wd < 256 ? ((Wd + 31) / 32) * 32 : 256;
  ProfScope prof(STAD_K_RESIZE, 0, F, Hd * Wd, Hs * Ws, stream);
  resize_cubic_u8_kernel<<<dim3(ceil_div(Wd, threads), Hd, F), threads, 0, stream>>>(in, out, Hs, Ws, Hd, Wd, xofs, xw, yofs,
                                                                                    yw);
  STAD_LAUNCH_OK("resize_cubic_u8");
  return STAD_OK;
}

}  // namespace stad
