// Fused joint space-time attention for sm_100a:  out = softmax(q k^T * scale) v  per (clip, head), head dim 64.
//
// Replaces Attention._naive_attn (modeling_finetune.py:93-103) and Attention._flash_attn -> FlashAttention.forward ->
// flash_attn_varlen_qkvpacked_func (modeling_finetune.py:121-128, flash_attention_class.py:39-51).  Input is the
// packed projection qkv[B, S, 3, H, 64] (the layout FlashAttention.forward takes); output is [B, S, H*64], the
// layout Attention.proj consumes.  The S x S score matrix never leaves the SM.
//
//   CTA = one 128-row query tile of one (b, h).            grid = (ceil(S/128), H, B)
//   warp 0     : TMA producer: Q once, then K/V tiles (128 keys) through a smem ring
//   warp 1     : MMA issuer (one thread): S_j = Q K_j^T  -> TMEM (2 buffers);  O += P_j V_j  (P_j read from TMEM)
//   warps 2..5 : softmax, one query row per thread: tcgen05.ld S_j, online max / sum in registers, exp2,
//                P_j (bf16) written back over S_j with tcgen05.st; O is rescaled lazily (only when the running
//                max grew by more than 2^8), and normalised + stored at the end.
//
// Roofline: dense BF16 tensor; algorithmic FLOPs = 4 S^2 64 per (b, h).  With head dim 64 the MUFU ex2 rate
// (one exp per 256 MMA flops) is the practical ceiling (SURVEY §7 "hard parts").
#include "kernels.h"
#include "ptx.cuh"

namespace stad {

namespace {

constexpr int HD = 64;        // head dim
constexpr int BQ = 128;       // query rows per CTA
constexpr int BKV = 128;      // keys per tile
constexpr int KV_STAGES = 4;
constexpr int Q_BYTES = BQ * HD * 2;
constexpr int K_BYTES = BKV * HD * 2;
constexpr int STAGE_BYTES = 2 * K_BYTES;  // K tile + V tile
constexpr int ATT_THREADS = 192;
constexpr int SMEM_BYTES = Q_BYTES + KV_STAGES * STAGE_BYTES + 256 + 1024;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t S_COL0 = 0;    // S buffer b at columns [b*128, b*128+128); P_j aliases its first 64 columns
constexpr uint32_t O_COL = 256;   // O accumulator: 64 fp32 columns
constexpr float kRescaleThreshold = 8.0f;  // log2 units

struct AttArgs {
  bf16* out;
  int B, H, S;
  float scale_log2;  // softmax scale * log2(e)
};

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_q = smem;
  uint8_t* smem_kv = smem + Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_kv + KV_STAGES * STAGE_BYTES);
  uint64_t* q_full = bars;                       // 1
  uint64_t* kv_full = bars + 1;                  // KV_STAGES
  uint64_t* kv_empty = kv_full + KV_STAGES;      // KV_STAGES
  uint64_t* s_full = kv_empty + KV_STAGES;       // 2
  uint64_t* p_full = s_full + 2;                 // 2
  uint64_t* o_full = p_full + 2;                 // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_kv = (p.S + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 4);  // one arrive per softmax warp
    }
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<TMEM_COLS>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int col_q = h * HD;
      const int col_k = (p.H + h) * HD;
      const int col_v = (2 * p.H + h) * HD;
      mbar_arrive_expect_tx(q_full, Q_BYTES);
      tma_load_3d(smem_q, &tmap_qkv, q_full, col_q, q0, b);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        uint8_t* sk = smem_kv + stage * STAGE_BYTES;
        mbar_arrive_expect_tx(&kv_full[stage], STAGE_BYTES);
        tma_load_3d(sk, &tmap_qkv, &kv_full[stage], col_k, j * BKV, b);
        tma_load_3d(sk + K_BYTES, &tmap_qkv, &kv_full[stage], col_v, j * BKV, b);
        if (++stage == KV_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0, 0);  // Q, K both K-major (head dim contiguous)
      constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, HD, 0, 1);   // P from TMEM (K-major), V MN-major (d contiguous)
      const uint64_t desc_q = make_smem_desc_sw128(smem_u32(smem_q), 16, 1024);
      const uint32_t tmem_o = tmem_base + O_COL;

      auto issue_qk = [&](int j, int stage) {
        const uint64_t desc_k = make_smem_desc_sw128(smem_u32(smem_kv + stage * STAGE_BYTES), 16, 1024);
        const uint32_t tmem_s = tmem_base + S_COL0 + (j & 1) * BKV;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_ss(tmem_s, desc_q + 2 * k, desc_k + 2 * k, idesc_qk, k != 0);
        umma_commit(&s_full[j & 1]);
      };

      mbar_wait(q_full, 0);
      int stage_qk = 0;  // ring position of the next K tile to multiply
      uint32_t phase_qk = 0;
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      stage_qk = 1 % KV_STAGES;
      if (stage_qk == 0) phase_qk ^= 1;

      int stage_pv = 0;
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) {
          mbar_wait(&kv_full[stage_qk], phase_qk);
          tc_fence_after();
          issue_qk(j + 1, stage_qk);
          if (++stage_qk == KV_STAGES) {
            stage_qk = 0;
            phase_qk ^= 1;
          }
        }
        mbar_wait(&p_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        // V tile: 128 keys x 64 d, one 128-byte row per key -> MN-major B operand; 16 keys = 2 x 1024 B per MMA
        const uint64_t desc_v = make_smem_desc_sw128(smem_u32(smem_kv + stage_pv * STAGE_BYTES + K_BYTES), 0, 1024);
        const uint32_t tmem_p = tmem_base + S_COL0 + (j & 1) * BKV;
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)
          umma_ts(tmem_o, tmem_p + k * 8, desc_v + (k * 2048 >> 4), idesc_pv, (j | k) != 0);
        umma_commit(&kv_empty[stage_pv]);
        umma_commit(o_full);
        if (++stage_pv == KV_STAGES) stage_pv = 0;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps (one query row per thread)
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // row inside the tile
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const float c = p.scale_log2;
    float m_ref = 0.f;  // reference max (log2 domain, already scaled)
    float l_sum = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const uint32_t s_addr = lane_addr + S_COL0 + (j & 1) * BKV;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      uint32_t sv[4][32];
      tmem_ld32(s_addr + 0, sv[0]);
      tmem_ld32(s_addr + 32, sv[1]);
      tmem_ld32(s_addr + 64, sv[2]);
      tmem_ld32(s_addr + 96, sv[3]);
      tmem_ld_wait();

      const int kv_valid = p.S - j * BKV;  // >= 1
      if (kv_valid < BKV) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (q * 32 + i >= kv_valid) sv[q][i] = 0xFF800000u;  // -inf
      }

      float mx = -INFINITY;
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(sv[q][i]));
      mx *= c;

      if (j == 0) {
        m_ref = mx;
      } else {
        // O and l_sum are relative to m_ref; only move the reference when the max grew by > 2^8 (keeps P <= 256)
        const bool grow = mx > m_ref + kRescaleThreshold;
        mbar_wait(o_full, (j - 1) & 1);  // PV(j-1) finished: O is stable, P(j-1) consumed
        tc_fence_after();
        if (__any_sync(0xffffffffu, grow)) {
          const float alpha = grow ? fast_exp2(m_ref - mx) : 1.0f;
          if (grow) {
            m_ref = mx;
            l_sum *= alpha;
          }
          uint32_t ov[2][32];
          tmem_ld32(lane_addr + O_COL, ov[0]);
          tmem_ld32(lane_addr + O_COL + 32, ov[1]);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[q][i] = __float_as_uint(__uint_as_float(ov[q][i]) * alpha);
          tmem_st32(lane_addr + O_COL, ov[0]);
          tmem_st32(lane_addr + O_COL + 32, ov[1]);
        }
      }

      // P = exp2(s*c - m_ref), packed to bf16 pairs: TMEM column k of P holds keys (2k, 2k+1)
      uint32_t pk[2][32];
      float sum = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float e0 = fast_exp2(fmaf(__uint_as_float(sv[q][i]), c, -m_ref));
          const float e1 = fast_exp2(fmaf(__uint_as_float(sv[q][i + 1]), c, -m_ref));
          sum += e0 + e1;
          pk[q >> 1][(q & 1) * 16 + (i >> 1)] = pack_bf16(e0, e1);
        }
      }
      l_sum += sum;
      tmem_st32(s_addr, pk[0]);
      tmem_st32(s_addr + 32, pk[1]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j & 1]);
    }

    // ---- finalise: O / l  -> bf16 -> out[b, q0 + r, h*64 .. h*64+63]
    mbar_wait(o_full, (n_kv - 1) & 1);
    tc_fence_after();
    uint32_t ov[2][32];
    tmem_ld32(lane_addr + O_COL, ov[0]);
    tmem_ld32(lane_addr + O_COL + 32, ov[1]);
    tmem_ld_wait();
    const int row = q0 + r;
    if (row < p.S) {
      const float inv = 1.0f / l_sum;
      uint4* op = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(b) * p.S + row) * (p.H * HD) + h * HD);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 o;
          o.x = pack_bf16(__uint_as_float(ov[q][i + 0]) * inv, __uint_as_float(ov[q][i + 1]) * inv);
          o.y = pack_bf16(__uint_as_float(ov[q][i + 2]) * inv, __uint_as_float(ov[q][i + 3]) * inv);
          o.z = pack_bf16(__uint_as_float(ov[q][i + 4]) * inv, __uint_as_float(ov[q][i + 5]) * inv);
          o.w = pack_bf16(__uint_as_float(ov[q][i + 6]) * inv, __uint_as_float(ov[q][i + 7]) * inv);
          op[q * 4 + (i >> 3)] = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace

int attention_init() {
  STAD_CUDA_OK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  return STAD_OK;
}

int launch_attention(const bf16* qkv, bf16* out, int B, int H, int S, float scale, cudaStream_t stream) {
  STAD_CHECK_ARG(B > 0 && H > 0 && S > 0, "attention: empty problem B=%d H=%d S=%d", B, H, S);
  STAD_CHECK_ARG(B <= 65535 && H <= 65535, "attention: B=%d / H=%d exceed the grid limits", B, H);
  if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15)
    return fail(STAD_E_ALIGN, "attention: qkv and out must be 16-byte aligned");
  const uint64_t row = static_cast<uint64_t>(3) * H * HD;  // elements per token in the packed projection
  const uint64_t dims[3] = {row, (uint64_t)S, (uint64_t)B};
  const uint64_t strides[2] = {row * 2, row * 2 * (uint64_t)S};
  const uint32_t box[3] = {HD, BQ, 1};
  CUtensorMap tm;
  int rc = make_tmap_bf16(&tm, qkv, 3, dims, strides, box);
  if (rc) return rc;
  AttArgs a;
  a.out = out;
  a.B = B;
  a.H = H;
  a.S = S;
  a.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid(ceil_div(S, BQ), H, B);
  ProfScope prof(STAD_K_ATTENTION, 0, B, H, S, stream);
  attention_kernel<<<grid, ATT_THREADS, SMEM_BYTES, stream>>>(tm, a);
  STAD_LAUNCH_OK("attention_kernel");
  return STAD_OK;
}

}  // namespace stad
