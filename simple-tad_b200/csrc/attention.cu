// Fused joint space-time attention for sm_100a:  out = softmax(q k^T * scale) v  per (clip, head), head dim 64.
//
// Replaces Attention._naive_attn (modeling_finetune.py:93-103) and Attention._flash_attn -> FlashAttention.forward ->
// flash_attn_varlen_qkvpacked_func (modeling_finetune.py:121-128, flash_attention_class.py:39-51).  Input is the
// packed projection qkv[B, S, 3, H, 64] (the layout FlashAttention.forward takes); output is [B, S, H*64], the
// layout Attention.proj consumes.  The S x S score matrix never leaves the SM.
//
// Persistent kernel, one CTA per SM.  A work unit = TWO 128-row query tiles ("slots") of one (b, h) that share every
// K/V tile streamed through shared memory (head dim 64 makes the kernel exp-bound, so the two slots also give every
// SM sub-partition two independent softmax warps to interleave on the MUFU).
//
//   warp 12      : TMA producer: Q tiles of the unit, then K and V tiles (128 keys each) through two smem rings
//   warps 13, 14 : MMA issuers, one thread per slot (warp 1 also owns TMEM).  S = Q K_j^T  (SS, fp32, 128 TMEM columns),
//                  O += P_j V_j  (A = P from TMEM, B = V as MN-major smem operand).  S, P and O live in separate
//                  TMEM columns, so Q K_{j+1}^T is issued as soon as the softmax warps have pulled S_j into registers.
//   warps 0..3   : softmax of slot 0, one query row per thread: tcgen05.ld S_j, row max (FMNMX3), exp2 of
//   warps 4..7   : softmax of slot 1   s*c - m (FFMA2 + MUFU.EX2; a fixed fraction of the pairs on the FMA pipe with a
//                  Cody-Waite / degree-3 polynomial), row sum (FADD2), P_j (bf16) -> TMEM by tcgen05.st.
//                  O is rescaled lazily (only when the running max grew by more than 2^8), by the same thread.
//                  The loop is software-pipelined: S_{j+1} is pulled into registers while the P_j stores drain.
//   warps 8..11  : epilogue: at the end of a unit they take the row sums from the softmax warps, read O, normalise
//                  and store it, so the softmax warps start the next unit immediately.
//
// TMEM (512 columns): S0 [0,128) S1 [128,256) P0 [256,320) P1 [320,384) O0 [384,448) O1 [448,512).
//
// Roofline: dense BF16 tensor; algorithmic FLOPs = 4 S^2 64 per (b, h).  With head dim 64 the MUFU ex2 rate
// (one exp per 256 MMA flops, 16 exp/clk/SM) is the practical ceiling (SURVEY §7 "hard parts").
#include "kernels.h"
#include "ptx.cuh"
#include "softmax_math.cuh"

namespace stad {

namespace {

constexpr int HD = 64;    // head dim
constexpr int BQ = 128;   // query rows per slot
constexpr int BKV = 128;  // keys per tile
constexpr int K_STAGES = 4;
constexpr int V_STAGES = 4;
constexpr int TILE_BYTES = BQ * HD * 2;  // 16 KB: every Q / K / V tile
constexpr int ATT_THREADS = 512;  // warpgroups 0, 1: softmax slots; 2: epilogue; 3: TMA + 2 MMA issuers (+1 idle warp)
constexpr int LSUM_BYTES = 2 * BQ * 4;  // row sums handed from the softmax warps to the epilogue warps
// Key-split tail units (see Unit::split): the four partial O rows of a query are combined through this scratch,
// [key quarter][column][query] fp32 (query fastest: conflict-free for writers and readers)
constexpr int SPLIT_ROWS = 32;
constexpr int SPLIT_SCRATCH_BYTES = 4 * HD * SPLIT_ROWS * 4;
constexpr int SMEM_BYTES = (2 + K_STAGES + V_STAGES) * TILE_BYTES + LSUM_BYTES + 512 + SPLIT_SCRATCH_BYTES + 1024;
#ifndef STAD_ATT_SPLIT_TAIL
#define STAD_ATT_SPLIT_TAIL 1
#endif
// Warp roles.  The single-thread TMA / MMA issuers sit in the HIGHEST warps: the sub-partition arbiter favours high warp
// ids, and an issuer that has to queue behind two always-ready softmax warps paces the whole kernel (measured: ~135
// clk per tcgen05.mma issue and ~350 clk per already-complete mbarrier wait when the issuers were warps 0-2).
constexpr int kTmaWarp = 12;   // warps 0-3: softmax slot 0, 4-7: softmax slot 1, 8-11: epilogue
constexpr int kMmaWarp0 = 13;  // 13: MMA issuer of slot 0, 14: of slot 1, 15: idle
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t S_COL = 0;    // + slot * 128
constexpr uint32_t P_COL = 256;  // + slot * 64   (bf16 pairs: column k holds keys 2k, 2k+1)
constexpr uint32_t O_COL = 384;  // + slot * 64
constexpr float kRescaleThreshold = 8.0f;  // log2 units

// -DSTAD_ATT_TRACE: CTA 0 records (clock, tag) events of one softmax warp per slot and of the two MMA issuers into
// att_trace (development builds only; read through stad_debug_read_att_trace, see tools/att_trace.py).
#ifdef STAD_ATT_TRACE
constexpr int kTraceCap = 2048;
__device__ unsigned long long att_trace[4][kTraceCap];
__device__ int att_trace_n[4];
#define ATT_EV(role, tag)                                                                          \
  do {                                                                                             \
    if ((STAD_ATT_TRACE >= 2 || (role) < 2) && blockIdx.x == 0 && tr_n < kTraceCap && (threadIdx.x & 31) == 0) {                          \
      att_trace[role][tr_n] = (static_cast<unsigned long long>(clock64()) << 8) | (tag);           \
      att_trace_n[role] = ++tr_n;                                                                  \
    }                                                                                              \
  } while (0)
#define ATT_T(i) do { if ((STAD_ATT_TRACE >= 2 || (i) == 7 || (i) == 2) && lane == 0 && quarter == 0) ATT_EV(slot, i); } while (0)
#else
#define ATT_EV(role, tag) do {} while (0)
#define ATT_T(i) do {} while (0)
#endif

struct AttArgs {
  bf16* out;
  int B, H, S;
  float scale_log2;  // softmax scale * log2(e)
};

// Development switches (A/B builds): software-pipelined S load; when to wait for the P buffer.
#ifndef STAD_ATT_PIPE
#define STAD_ATT_PIPE 0
#endif
#ifndef STAD_ATT_STAGGER
#define STAD_ATT_STAGGER 2
#endif
#ifndef STAD_ATT_STAGGER_AT
#define STAD_ATT_STAGGER_AT 0  // slot 1 is released once slot 0 has stored this many + 1 chunks of its first P tile (measured,
                                // S = 1568: no stagger 722 us; after max 728; chunk 0: 670; chunk 1: 689; chunk 2: 698; chunk 3: 709)
#endif
#ifndef STAD_ATT_PROBE
#define STAD_ATT_PROBE 1
#endif
#ifndef STAD_ATT_SPLIT_LD
#define STAD_ATT_SPLIT_LD 1
#endif
#ifndef STAD_ATT_STORE_MODE
#define STAD_ATT_STORE_MODE 0
#endif

// A one-slot unit whose query tile holds at most 32 rows (the ragged end of the sequence: rows 1536..1567 of 1568)
// would keep ONE warp busy per tile while costing a whole unit's worth of MMAs and latencies (measured: 11 % of the
// kernel for 2 % of the rows).  It runs "key-split" instead: the 32 queries are replicated into all four 32-row groups
// of the Q tile, so every TMEM lane quarter holds their scores against all 128 keys of a tile, and the softmax warp of
// quarter k handles only keys [32k, 32k+32) of each tile with its own running max / sum (its P entries for the other
// keys stay zero).  The four partial outputs of a query are merged by the epilogue warps.
struct Unit {
  int b, h, q0, slots;
  bool split;
};

STAD_DEVICE Unit decode_unit(int u, int units_per_head, int H, int S) {
  Unit w;
  const int bh = u / units_per_head;
  const int t = u - bh * units_per_head;
  w.b = bh / H;
  w.h = bh - w.b * H;
  w.q0 = t * 2 * BQ;
  w.slots = (w.q0 + BQ < S) ? 2 : 1;
  w.split = STAD_ATT_SPLIT_TAIL && w.slots == 1 && S - w.q0 <= SPLIT_ROWS;
  return w;
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_q32,
                 const AttArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_q = smem;                                   // [2][16 KB]
  uint8_t* smem_k = smem_q + 2 * TILE_BYTES;                // [K_STAGES][16 KB]
  uint8_t* smem_v = smem_k + K_STAGES * TILE_BYTES;         // [V_STAGES][16 KB]
  float* lsum_smem = reinterpret_cast<float*>(smem_v + V_STAGES * TILE_BYTES);  // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + V_STAGES * TILE_BYTES + LSUM_BYTES);
  uint64_t* q_full = bars;                   // [2]  TMA -> MMA
  uint64_t* q_free = q_full + 2;             // [2]  MMA (last Q K^T of the unit) -> TMA
  uint64_t* k_full = q_free + 2;             // [K_STAGES]
  uint64_t* k_free = k_full + K_STAGES;      // [K_STAGES]
  uint64_t* v_full = k_free + K_STAGES;      // [V_STAGES]
  uint64_t* v_free = v_full + V_STAGES;      // [V_STAGES]
  uint64_t* s_full = v_free + V_STAGES;      // [2]  MMA -> softmax: S_j complete
  uint64_t* s_free = s_full + 2;             // [2]  softmax (4 warps) -> MMA: S_j is in registers
  uint64_t* p_full = s_free + 2;             // [2]  softmax (4 warps) -> MMA: P_j stored
  uint64_t* o_full = p_full + 2;             // [2]  MMA -> softmax: P_j V_j complete
  uint64_t* l_ready = o_full + 2;            // [2]  softmax (128 threads) -> epilogue: unit done, row sums in smem
  uint64_t* o_free = l_ready + 2;            // [2]  epilogue (4 warps) -> MMA: O has been read, next unit may overwrite
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);
  float* split_scratch = reinterpret_cast<float*>(smem_v + V_STAGES * TILE_BYTES + LSUM_BYTES + 512);

  // warp index through a shuffle: the compiler then knows every value derived from it is warp-uniform (uniform
  // registers feed tcgen05.mma / TMA directly instead of a per-lane ELECT + R2UR.BROADCAST loop per instruction)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n_q = (p.S + BQ - 1) / BQ;
  const int units_per_head = (n_q + 1) / 2;
  const int total_units = p.B * p.H * units_per_head;
  const int n_kv = (p.S + BKV - 1) / BKV;
  const int last_valid = p.S - (n_kv - 1) * BKV;   // valid keys of the last K/V tile, 1..128
  const int last_chunks = (last_valid + 31) >> 5;  // 32-column chunks of the last tile that hold any valid key

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_q32);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_free[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_free[s], 4);  // one arrive per softmax warp
      mbar_init(&p_full[s], 4);
      mbar_init(&o_full[s], 1);
      mbar_init(&l_ready[s], BQ);
      mbar_init(&o_free[s], 4);
    }
    for (int s = 0; s < K_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_free[s], 2);  // one arrive per slot
    }
    for (int s = 0; s < V_STAGES; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_free[s], 2);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp0) {
    tmem_alloc<TMEM_COLS>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // prologue done (own shared memory / TMEM only); see launch_pdl in common.h
  pdl_launch_dependents();
  pdl_wait();

  if (warp >= kTmaWarp) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == kTmaWarp) {
      // ---------------------------------------------------------------- TMA producer (whole warp; one elected lane issues)
      uint32_t kst = 0, kph = 0, vst = 0, vph = 0;
      uint32_t ucnt0 = 0, ucnt1 = 0;  // units started per slot (phase of q_full / q_free)
      auto load_tile = [&](uint64_t* full, uint8_t* dst, int col, int row, int b) {
        if (elect_one()) {
          mbar_arrive_expect_tx(full, TILE_BYTES);
          tma_load_3d(dst, &tmap_qkv, full, col, row, b);
        }
        __syncwarp();
      };
      auto load_k = [&](int j, int b, int col_k) {
        mbar_wait(&k_free[kst], kph ^ 1);
        load_tile(&k_full[kst], smem_k + kst * TILE_BYTES, col_k, j * BKV, b);
        if (++kst == K_STAGES) {
          kst = 0;
          kph ^= 1;
        }
      };
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const Unit w = decode_unit(u, units_per_head, p.H, p.S);
        const int col_q = w.h * HD;
        const int col_k = (p.H + w.h) * HD;
        const int col_v = (2 * p.H + w.h) * HD;
        mbar_wait(&q_free[0], (ucnt0 & 1) ^ 1);
        if (w.split) {
          // the same 32 query rows into each of the four 32-row groups of the tile (4 KB apart in the swizzled layout)
          if (elect_one()) {
            mbar_arrive_expect_tx(&q_full[0], TILE_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tma_load_3d(smem_q + k * (SPLIT_ROWS * HD * 2), &tmap_q32, &q_full[0], col_q, w.q0, w.b);
          }
          __syncwarp();
        } else {
          load_tile(&q_full[0], smem_q, col_q, w.q0, w.b);
        }
        ++ucnt0;
        if (w.slots > 1) {
          mbar_wait(&q_free[1], (ucnt1 & 1) ^ 1);
          load_tile(&q_full[1], smem_q + TILE_BYTES, col_q, w.q0 + BQ, w.b);
          ++ucnt1;
        }
        load_k(0, w.b, col_k);
        for (int j = 0; j < n_kv; ++j) {
          if (j + 1 < n_kv) load_k(j + 1, w.b, col_k);
          mbar_wait(&v_free[vst], vph ^ 1);
          load_tile(&v_full[vst], smem_v + vst * TILE_BYTES, col_v, j * BKV, w.b);
          if (++vst == V_STAGES) {
            vst = 0;
            vph ^= 1;
          }
        }
      }
    } else if (warp == kMmaWarp0 || warp == kMmaWarp0 + 1) {
      // ---------------------------------------------------------------- MMA issuers: warp 13 -> slot 0, 14 -> slot 1
      // One issuing warp per slot keeps the two slots independent: neither ever waits behind a barrier of the other.
      // The whole warp runs the (warp-uniform) control flow and waits; one elected lane issues tcgen05.mma / commit.
      // Both slots read the same K / V ring stages; a stage is released by two arrivals (one per slot; in a one-slot
      // unit the slot-0 warp commits twice).
      const int slot = warp - kMmaWarp0;
#ifdef STAD_ATT_TRACE
      int tr_n = 0;
#endif
      constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, HD, 0, 1);  // P from TMEM (K-major), V MN-major (d contiguous)
      constexpr uint32_t idesc_qk_full = make_idesc_bf16(BQ, BKV, 0, 0);  // Q, K both K-major
      const uint64_t desc_q = make_smem_desc_sw128(smem_u32(smem_q + slot * TILE_BYTES), 16, 1024);
      const uint32_t tmem_s = tmem_base + S_COL + slot * BKV;
      const uint32_t tmem_p = tmem_base + P_COL + slot * (BKV / 2);
      const uint32_t tmem_o = tmem_base + O_COL + slot * HD;
      uint32_t kc = 0, vc = 0;  // K / V tiles consumed so far (ring position = count % stages)
      uint32_t g = 0;           // softmax iterations completed by this slot (phase of s_full/s_free/p_full/o_full)
      uint32_t ucnt = 0;        // units started by this slot (phase of q_full / q_free)

      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const Unit w = decode_unit(u, units_per_head, p.H, p.S);
        if (slot >= w.slots) {  // one-slot unit: slot 1 sits it out but stays in step with the rings
          kc += n_kv;
          vc += n_kv;
          continue;
        }
        const bool solo = w.slots == 1;
        // S = Q K^T over the first `cols` keys of K tile number kc (its k_full wait has been done); releases the stage
        // (`solo_unit`: the unit the tile belongs to has one slot, so this warp releases the stage for both)
        auto issue_qk = [&](int cols, bool last_of_unit, bool solo_unit) {
          const uint32_t kst = kc % K_STAGES;
          tc_fence_after();
          ATT_EV(2 + slot, 20);
          if (elect_one()) {
            const uint64_t desc_k = make_smem_desc_sw128(smem_u32(smem_k + kst * TILE_BYTES), 16, 1024);
#ifdef STAD_ATT_NO_MMA
            // development build (-DSTAD_ATT_NO_MMA, tools/build_variant.sh): the barrier protocol runs, the tensor core
            // does not (results are garbage).  Splits the cost of the hand-overs from the cost of sharing TMEM / the SM
            // with the MMAs.
            (void)desc_k;
            (void)cols;
#else
            if (cols == BKV) {
#pragma unroll
              for (int k = 0; k < HD / 16; ++k) umma_ss(tmem_s, desc_q + 2 * k, desc_k + 2 * k, idesc_qk_full, k != 0);
            } else {
              const uint32_t idesc = make_idesc_bf16(BQ, static_cast<uint32_t>(cols), 0, 0);
#pragma unroll
              for (int k = 0; k < HD / 16; ++k) umma_ss(tmem_s, desc_q + 2 * k, desc_k + 2 * k, idesc, k != 0);
            }
#endif
            ATT_EV(2 + slot, 21);
            umma_commit(&s_full[slot]);
            if (last_of_unit) umma_commit(&q_free[slot]);
            umma_commit(&k_free[kst]);
            if (solo_unit) umma_commit(&k_free[kst]);
          }
          __syncwarp();
          ++kc;
        };
        // O (+)= P V over the first 16 * ksteps keys of V tile number vc.  V tile: one 128-byte row per key -> MN-major
        // B operand; 16 keys = 2 x 1024 B per MMA.  Releases the V stage.
        auto issue_pv = [&](bool accumulate, int ksteps) {
          const uint32_t vst = vc % V_STAGES;
          tc_fence_after();
          ATT_EV(2 + slot, 22);
          if (elect_one()) {
            const uint64_t desc_v = make_smem_desc_sw128(smem_u32(smem_v + vst * TILE_BYTES), 0, 1024);
#ifdef STAD_ATT_NO_MMA
            (void)desc_v;
            (void)accumulate;
            (void)ksteps;
#else
            if (ksteps == BKV / 16) {
#pragma unroll
              for (int k = 0; k < BKV / 16; ++k)
                umma_ts(tmem_o, tmem_p + k * 8, desc_v + (k * 2048 >> 4), idesc_pv, (accumulate || k != 0) ? 1u : 0u);
            } else {
              for (int k = 0; k < ksteps; ++k)
                umma_ts(tmem_o, tmem_p + k * 8, desc_v + (k * 2048 >> 4), idesc_pv, (accumulate || k != 0) ? 1u : 0u);
            }
#endif
            ATT_EV(2 + slot, 23);
            umma_commit(&o_full[slot]);
            umma_commit(&v_free[vst]);
            if (solo) umma_commit(&v_free[vst]);
          }
          __syncwarp();
          ++vc;
        };
        auto wait_k = [&]() { mbar_wait(&k_full[kc % K_STAGES], (kc / K_STAGES) & 1); };
        auto wait_v = [&]() { mbar_wait(&v_full[vc % V_STAGES], (vc / V_STAGES) & 1); };

        mbar_wait(&q_full[slot], ucnt & 1);
        wait_k();
        if (g > 0) mbar_wait(&s_free[slot], (g - 1) & 1);  // previous S of the slot has been pulled out of TMEM
        issue_qk(n_kv == 1 ? last_chunks * 32 : BKV, n_kv == 1, solo);
        for (int j = 0; j < n_kv; ++j) {
          const bool has_next = j + 1 < n_kv;
          const bool next_last = j + 2 == n_kv;
          // the operand tiles arrive long before the softmax signals: wait for them first, so that only the MMA issue
          // itself follows the s_free / p_full arrival
          if (has_next) {
            wait_k();
            ATT_EV(2 + slot, 10);
            mbar_wait(&s_free[slot], (g + j) & 1);
            ATT_EV(2 + slot, 11);
            issue_qk(next_last ? last_chunks * 32 : BKV, next_last, solo);
            ATT_EV(2 + slot, 12);
          }
          wait_v();
          if (j == 0 && ucnt > 0) mbar_wait(&o_free[slot], (ucnt - 1) & 1);  // epilogue has read the previous unit's O
          mbar_wait(&p_full[slot], (g + j) & 1);
          ATT_EV(2 + slot, 13);
          issue_pv(j != 0, has_next ? BKV / 16 : last_chunks * 2);
          ATT_EV(2 + slot, 14);
        }
        g += n_kv;
        ++ucnt;
      }
    }
  } else if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    // ------------------------------------------------------------------ epilogue warps: O / l -> bf16 -> global
    // Takes the end-of-unit work off the softmax warps: they hand over the row sums and move on to the next unit.
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint32_t gs[2] = {0, 0};    // iterations completed per slot
    uint32_t ucs[2] = {0, 0};   // units completed per slot
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const Unit w = decode_unit(u, units_per_head, p.H, p.S);
#pragma unroll
      for (int slot = 0; slot < 2; ++slot) {
        if (slot >= w.slots) continue;
        gs[slot] += n_kv;
        mbar_wait(&l_ready[slot], ucs[slot] & 1);          // every softmax thread of the slot finished the unit
        mbar_wait(&o_full[slot], (gs[slot] - 1) & 1);      // last P V of the unit
        tc_fence_after();
        ++ucs[slot];
        if (w.split) {
          // ---- merge the four key-quarter partials of each of the 32 queries:
          //   out = sum_k 2^(m_k - M) O_k / sum_k 2^(m_k - M) l_k,  M = max_k m_k
          float mk[4], M = -INFINITY;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            mk[k] = lsum_smem[BQ + k * 32 + lane];
            M = fmaxf(M, mk[k]);
          }
          float L = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) L += (mk[k] == -INFINITY ? 0.f : ex2(mk[k] - M)) * lsum_smem[k * 32 + lane];
          const float wgt = (mk[quarter] == -INFINITY ? 0.f : ex2(mk[quarter] - M)) / L;
          named_bar_sync(2, 128);  // the previous split unit's combine has finished reading the scratch
          const uint32_t o_addr = lane_addr + O_COL;
          float* mine = split_scratch + quarter * (HD * SPLIT_ROWS) + lane;
#pragma unroll 1
          for (int q = 0; q < 4; ++q) {  // 16 columns at a time: these warps run with 56 registers
            uint32_t ov[16];
            tmem_ld16(o_addr + q * 16, ov);
            tmem_ld_wait16(ov);
            if (q == 3) {  // O is in registers: the next unit's first P V may overwrite it
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&o_free[slot]);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) mine[(q * 16 + i) * SPLIT_ROWS] = __uint_as_float(ov[i]) * wgt;
          }
          named_bar_sync(2, 128);  // all four partials are in the scratch
          // thread (cg = quarter, query = lane): 16 output columns [16 cg, +16) of query q0 + lane
          const int qrow = w.q0 + lane;
          if (qrow < p.S) {
            float acc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float* src = split_scratch + (quarter * 16 + i) * SPLIT_ROWS + lane;
              acc[i] = (src[0] + src[HD * SPLIT_ROWS]) + (src[2 * HD * SPLIT_ROWS] + src[3 * HD * SPLIT_ROWS]);
            }
            uint4* op = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(w.b) * p.S + qrow) * (p.H * HD) + w.h * HD +
                                                 quarter * 16);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              uint4 o;
              o.x = pack_bf16(acc[8 * i + 0], acc[8 * i + 1]);
              o.y = pack_bf16(acc[8 * i + 2], acc[8 * i + 3]);
              o.z = pack_bf16(acc[8 * i + 4], acc[8 * i + 5]);
              o.w = pack_bf16(acc[8 * i + 6], acc[8 * i + 7]);
              op[i] = o;
            }
          }
          continue;
        }
        const int row = w.q0 + slot * BQ + r;
        const bool warp_valid = w.q0 + slot * BQ + quarter * 32 < p.S;
        const float inv = 1.0f / lsum_smem[slot * BQ + r];
        const uint32_t o_addr = lane_addr + O_COL + slot * HD;
        uint4* op = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(w.b) * p.S + row) * (p.H * HD) + w.h * HD);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint32_t ov[32];
          if (warp_valid) {
            tmem_ld32(o_addr + q * 32, ov);
            tmem_ld_wait32(ov);
          }
          if (q == 1) {  // O is in registers: the next unit's first P V may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_free[slot]);
          }
          if (warp_valid && row < p.S) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              uint4 o;
              o.x = pack_bf16(__uint_as_float(ov[i + 0]) * inv, __uint_as_float(ov[i + 1]) * inv);
              o.y = pack_bf16(__uint_as_float(ov[i + 2]) * inv, __uint_as_float(ov[i + 3]) * inv);
              o.z = pack_bf16(__uint_as_float(ov[i + 4]) * inv, __uint_as_float(ov[i + 5]) * inv);
              o.w = pack_bf16(__uint_as_float(ov[i + 6]) * inv, __uint_as_float(ov[i + 7]) * inv);
              op[q * 4 + (i >> 3)] = o;
            }
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // ------------------------------------------------------------------ softmax warps (one query row per thread)
    const int slot = warp >> 2;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t s_addr = lane_addr + S_COL + slot * BKV;
    const uint32_t p_addr = lane_addr + P_COL + slot * (BKV / 2);
    const uint32_t o_addr = lane_addr + O_COL + slot * HD;
    const float c = p.scale_log2;
    const bool last_full = last_valid == BKV;
    uint32_t g = 0;  // iterations completed by this slot
#ifdef STAD_ATT_TRACE
    int tr_n = 0;
#endif

    // Lazy rescale of this thread's O row (rare: only when the running max grew by more than 2^8).
    auto rescale_o = [&](float alpha) {
#pragma unroll 1
      for (int q = 0; q < 2; ++q) {
        uint32_t ov[32];
        tmem_ld32(o_addr + q * 32, ov);
        tmem_ld_wait32(ov);
#pragma unroll
        for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
        tmem_st32(o_addr + q * 32, ov);
      }
    };

    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const Unit w = decode_unit(u, units_per_head, p.H, p.S);
      if (slot >= w.slots) continue;
      if (w.split) {
        // ---- key-split tail unit (slot 0 only): lane i of EVERY quarter holds query q0 + i; this warp handles the
        // keys [32 * quarter, +32) of each K/V tile: one 32-column chunk of S, one 16-column chunk of P.
        float m_ref = -INFINITY, l_sum = 0.f;
        const uint32_t my_s = s_addr + quarter * 32;
        const uint32_t my_p = p_addr + quarter * 16;
        for (int j = 0; j < n_kv; ++j, ++g) {
          const int valid = ((j + 1 < n_kv) ? BKV : last_valid) - quarter * 32;  // keys of this tile in my chunk
          mbar_wait(&s_full[slot], g & 1);
          tc_fence_after();
          uint32_t t[32];
          if (valid > 0) {
            tmem_ld32(my_s, t);
            tmem_ld_wait32(t);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[slot]);
          if (j == 0) {
            // P entries of my lanes for the OTHER key chunks stay zero for the whole unit (the P buffer is free here:
            // s_full of the first tile completes after the last P V of the previous unit)
            uint32_t z[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (q != quarter) tmem_st16(p_addr + q * 16, z);
          }
          if (valid > 0) {
            if (valid < 32) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i >= valid) t[i] = 0xFF800000u;  // -inf
            }
            const float mx = c * chunk_max(t);
            if (j > 0) {
              mbar_wait(&o_full[slot], (g - 1) & 1);  // P V of the previous tile: P buffer reusable, O stable
              tc_fence_after();
            }
            const bool grow = mx > m_ref + kRescaleThreshold;  // always true on the first tile with keys (m_ref = -inf)
            if (__any_sync(0xffffffffu, grow)) {
              const float alpha = (grow && m_ref != -INFINITY) ? ex2(m_ref - mx) : 1.0f;
              if (grow) {
                m_ref = mx;
                l_sum *= alpha;
              }
              if (j > 0) rescale_o(alpha);
            }
            uint32_t pk[16];
            float a0 = 0.f, a1 = 0.f;
            exp_chunk<false>(t, c, -m_ref, a0, a1, pk);
            l_sum += a0 + a1;
            tmem_st16(my_p, pk);
          } else if (j > 0) {
            // no key of this tile falls into my chunk (ragged last tile): its P V reads none of my P columns beyond
            // the zeroed ones; just keep the barrier protocol in step
            mbar_wait(&o_full[slot], (g - 1) & 1);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[slot]);
        }
        lsum_smem[r] = l_sum;        // partial sum of (query lane, key quarter)
        lsum_smem[BQ + r] = m_ref;   // its reference max (-inf: this quarter never saw a key); slot 1's half is idle here
        mbar_arrive(&l_ready[slot]);
        continue;
      }
      const int row0 = w.q0 + slot * BQ;
      const bool warp_valid = row0 + quarter * 32 < p.S;  // warp-uniform: any valid query row in this warp
      float m_ref = 0.f;  // reference max (log2 domain, already scaled)
      float l_sum = 0.f;
      uint32_t sv[4][32];
      bool have_s = false;  // S_j was pulled into sv by the previous iteration (software pipelining)
      bool s_probe = false; // s_full of the coming iteration was already seen complete

      for (int j = 0; j < n_kv; ++j, ++g) {
        ATT_T(7);
#if STAD_ATT_STAGGER
        const bool kStaggerNow = STAD_ATT_STAGGER == 2 ? (j == 0 && w.slots == 2) : (g == 0);
        // Phase offset between the two slots: without it both softmax warpgroups run in lockstep (same phase of the
        // iteration at the same time), i.e. they fight for the MUFU together and idle together.  Mode 1 offsets them once
        // per launch; the offset then drifts back to in-phase within ~20 iterations (measured, tools/att_offsets.py), so
        // mode 2 re-establishes it at the first tile of every two-slot unit: slot 1 starts its softmax when slot 0 has
        // stored the first half of its P tile.
        if (kStaggerNow) {
          if (slot == 1) named_bar_sync(1, 2 * BQ);
          else if (!warp_valid) asm volatile("bar.arrive 1, %0;" ::"n"(2 * BQ) : "memory");
        }
#endif
        if (!warp_valid) {
          // rows beyond S: nothing to compute (their P / O rows are never stored); keep the pipeline moving.
          // The p_full arrival must not overtake phase j-1 of that barrier (the computing warps may still be
          // working on P_{j-1}); P V_{j-1} done implies p_full phase j-1 has completed.
          mbar_wait(&s_full[slot], g & 1);
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&s_free[slot]);
            if (j > 0) mbar_wait(&o_full[slot], (g - 1) & 1);
            mbar_arrive(&p_full[slot]);
          }
          continue;
        }
        const bool full_tile = (j + 1 < n_kv) || last_full;
        if (full_tile) {
          float mx;
          if (!have_s) {
            if (!s_probe) mbar_wait(&s_full[slot], g & 1);  // usually already seen complete by last iteration's probe
            tc_fence_after();
            ATT_T(0);
#if STAD_ATT_SPLIT_LD
            // two loads, wait, two more loads in flight while the first half's row max is computed
            tmem_ld32(s_addr + 0, sv[0]);
            tmem_ld32(s_addr + 32, sv[1]);
            tmem_ld_wait32(sv[0]);
            tmem_ld_wait32(sv[1]);
            tmem_ld32(s_addr + 64, sv[2]);
            tmem_ld32(s_addr + 96, sv[3]);
            const float m01 = fmaxf(chunk_max(sv[0]), chunk_max(sv[1]));
            tmem_ld_wait32(sv[2]);
            tmem_ld_wait32(sv[3]);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[slot]);  // S_j is in registers: Q K_{j+1}^T may overwrite it
            ATT_T(1);
#if STAD_ATT_STAGGER && STAD_ATT_STAGGER_AT == -2
            if (kStaggerNow && slot == 0) asm volatile("bar.arrive 1, %0;" ::"n"(2 * BQ) : "memory");
#endif
            mx = c * max3(m01, chunk_max(sv[2]), chunk_max(sv[3]));
#else
            tmem_ld32(s_addr + 0, sv[0]);
            tmem_ld32(s_addr + 32, sv[1]);
            tmem_ld32(s_addr + 64, sv[2]);
            tmem_ld32(s_addr + 96, sv[3]);
            tmem_ld_wait32(sv[0]);
            tmem_ld_wait32(sv[1]);
            tmem_ld_wait32(sv[2]);
            tmem_ld_wait32(sv[3]);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[slot]);  // S_j is in registers: Q K_{j+1}^T may overwrite it
            ATT_T(1);
            mx = c * max3(fmaxf(chunk_max(sv[0]), chunk_max(sv[1])), chunk_max(sv[2]), chunk_max(sv[3]));
#endif
          } else {
            ATT_T(1);
            mx = c * max3(fmaxf(chunk_max(sv[0]), chunk_max(sv[1])), chunk_max(sv[2]), chunk_max(sv[3]));
          }
          s_probe = false;
          bool pv_done = (j == 0);  // P V of the previous iteration finished (P buffer reusable, O stable)
          if (j == 0) {
            m_ref = mx;
          } else {
            // O and l_sum are relative to m_ref; only move the reference when the max grew by > 2^8 (keeps P <= 256)
            const bool grow = mx > m_ref + kRescaleThreshold;
            if (__any_sync(0xffffffffu, grow)) {
              mbar_wait(&o_full[slot], (g - 1) & 1);
              tc_fence_after();
              pv_done = true;
              const float alpha = grow ? ex2(m_ref - mx) : 1.0f;
              if (grow) {
                m_ref = mx;
                l_sum *= alpha;
              }
              rescale_o(alpha);
            }
          }
          ATT_T(2);
#if STAD_ATT_STAGGER && STAD_ATT_STAGGER_AT == -1
          if (kStaggerNow && slot == 0) asm volatile("bar.arrive 1, %0;" ::"n"(2 * BQ) : "memory");
#endif
          // probe the barrier the first P store needs while chunk 0 is computed (the probe's latency is hidden)
          const bool o_probe = !pv_done && STAD_ATT_PROBE && mbar_try_wait(&o_full[slot], (g - 1) & 1);
          const float neg_m = -m_ref;
          float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
          const bool next_full = STAD_ATT_PIPE && ((j + 2 < n_kv) || (j + 1 < n_kv && last_full));
          // Each chunk's P is stored as soon as it is packed (keeps the chunks' instruction streams apart: FFMA2, MUFU
          // and F2FP of neighbouring chunks overlap, but the scheduler cannot lump all MUFUs of the tile together).
          {
            uint32_t pk0[16], pk1[16];
            exp_chunk<true>(sv[0], c, neg_m, a0, a1, pk0);
#if STAD_ATT_STORE_MODE == 0
            ATT_T(8);
            if (!pv_done && !o_probe) mbar_wait(&o_full[slot], (g - 1) & 1);  // P V_{j-1} has consumed the P buffer
            tc_fence_after();
            ATT_T(9);
            tmem_st16(p_addr, pk0);
#if STAD_ATT_STAGGER && STAD_ATT_STAGGER_AT == 0
            if (kStaggerNow && slot == 0) asm volatile("bar.arrive 1, %0;" ::"n"(2 * BQ) : "memory");
#endif
            exp_chunk<true>(sv[1], c, neg_m, b0, b1, pk1);
            tmem_st16(p_addr + 16, pk1);
#if STAD_ATT_STAGGER && STAD_ATT_STAGGER_AT == 1
            if (kStaggerNow && slot == 0) asm volatile("bar.arrive 1, %0;" ::"n"(2 * BQ) : "memory");
#endif
#else
            const bool o_ok = pv_done || mbar_try_wait(&o_full[slot], (g - 1) & 1);
            exp_chunk<true>(sv[1], c, neg_m, b0, b1, pk1);
            if (!o_ok) mbar_wait(&o_full[slot], (g - 1) & 1);
            tc_fence_after();
            tmem_st16(p_addr, pk0);
            tmem_st16(p_addr + 16, pk1);
#endif
          }
          {
            uint32_t pk2[16];
            exp_chunk<true>(sv[2], c, neg_m, a0, a1, pk2);
            tmem_st16(p_addr + 32, pk2);
#if STAD_ATT_STAGGER && STAD_ATT_STAGGER_AT == 2
            if (kStaggerNow && slot == 0) asm volatile("bar.arrive 1, %0;" ::"n"(2 * BQ) : "memory");
#endif
          }
          bool s_ok = false;
          if (next_full) s_ok = mbar_try_wait(&s_full[slot], (g + 1) & 1);
          {
            uint32_t pk3[16];
            exp_chunk<true>(sv[3], c, neg_m, b0, b1, pk3);
            tmem_st16(p_addr + 48, pk3);
#if STAD_ATT_STAGGER && STAD_ATT_STAGGER_AT == 3
            if (kStaggerNow && slot == 0) asm volatile("bar.arrive 1, %0;" ::"n"(2 * BQ) : "memory");
#endif
          }
          l_sum += (a0 + a1) + (b0 + b1);
          ATT_T(3);
          if (next_full) {
            // pull S_{j+1} while the P stores drain: the score registers are free again
            if (!s_ok) mbar_wait(&s_full[slot], (g + 1) & 1);
            tc_fence_after();
            tmem_ld32(s_addr + 0, sv[0]);
            tmem_ld32(s_addr + 32, sv[1]);
            tmem_ld32(s_addr + 64, sv[2]);
            tmem_ld32(s_addr + 96, sv[3]);
          }
          ATT_T(4);
          if (STAD_ATT_PROBE && !next_full && j + 1 < n_kv) s_probe = mbar_try_wait(&s_full[slot], (g + 1) & 1);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[slot]);
          if (next_full) {
            tmem_ld_wait32(sv[0]);
            tmem_ld_wait32(sv[1]);
            tmem_ld_wait32(sv[2]);
            tmem_ld_wait32(sv[3]);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[slot]);
          }
          have_s = next_full;
          ATT_T(5);
        } else {
          // ---- last K/V tile with fewer than 128 valid keys
          mbar_wait(&s_full[slot], g & 1);
          tc_fence_after();
          if (last_chunks == 1) {
            // at most 32 valid keys (S = 1568: exactly 32): one chunk, kept in registers between the max and the exps;
            // S is released before the exps so the issuer is not held up
            uint32_t t[32];
            tmem_ld32(s_addr, t);
            tmem_ld_wait32(t);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[slot]);
            if (last_valid < 32) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i >= last_valid) t[i] = 0xFF800000u;  // -inf
            }
            const float mx = c * chunk_max(t);
            if (j == 0) {
              m_ref = mx;
            } else {
              mbar_wait(&o_full[slot], (g - 1) & 1);
              tc_fence_after();
              const bool grow = mx > m_ref + kRescaleThreshold;
              if (__any_sync(0xffffffffu, grow)) {
                const float alpha = grow ? ex2(m_ref - mx) : 1.0f;
                if (grow) {
                  m_ref = mx;
                  l_sum *= alpha;
                }
                rescale_o(alpha);
              }
            }
            uint32_t pk[16];
            float a0 = 0.f, a1 = 0.f;
            exp_chunk<false>(t, c, -m_ref, a0, a1, pk);
            tmem_st16(p_addr, pk);
            l_sum += a0 + a1;
#if STAD_ATT_STAGGER
            if (kStaggerNow && slot == 0) asm volatile("bar.arrive 1, %0;" ::"n"(2 * BQ) : "memory");
#endif
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[slot]);
            have_s = false;
            continue;
          }
          // general case: chunk loop, S read twice (max pass, exp pass)
          float rmax = -INFINITY;
#pragma unroll 1
          for (int q = 0; q < last_chunks; ++q) {
            uint32_t t[32];
            tmem_ld32(s_addr + q * 32, t);
            tmem_ld_wait32(t);
            const int valid = last_valid - q * 32;
            if (valid < 32) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i >= valid) t[i] = 0xFF800000u;  // -inf
            }
            rmax = fmaxf(rmax, chunk_max(t));
          }
          const float mx = c * rmax;
          if (j == 0) {
            m_ref = mx;
          } else {
            mbar_wait(&o_full[slot], (g - 1) & 1);
            tc_fence_after();
            const bool grow = mx > m_ref + kRescaleThreshold;
            if (__any_sync(0xffffffffu, grow)) {
              const float alpha = grow ? ex2(m_ref - mx) : 1.0f;
              if (grow) {
                m_ref = mx;
                l_sum *= alpha;
              }
              rescale_o(alpha);
            }
          }
          const float neg_m = -m_ref;
          float a0 = 0.f, a1 = 0.f;
#pragma unroll 1
          for (int q = 0; q < last_chunks; ++q) {
            uint32_t t[32];
            tmem_ld32(s_addr + q * 32, t);
            tmem_ld_wait32(t);
            const int valid = last_valid - q * 32;
            if (valid < 32) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i >= valid) t[i] = 0xFF800000u;
            }
            uint32_t pk[16];
            exp_chunk<false>(t, c, neg_m, a0, a1, pk);
            tmem_st16(p_addr + q * 16, pk);
          }
          l_sum += a0 + a1;
#if STAD_ATT_STAGGER
          if (kStaggerNow && slot == 0) asm volatile("bar.arrive 1, %0;" ::"n"(2 * BQ) : "memory");
#endif
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[slot]);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[slot]);
          have_s = false;
        }
      }
      // ---- hand the row sum to the epilogue warps and go on with the next unit
      lsum_smem[slot * BQ + r] = l_sum;
      mbar_arrive(&l_ready[slot]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace

#ifdef STAD_ATT_TRACE
extern "C" __attribute__((visibility("default"))) int stad_debug_read_att_trace(unsigned long long* out, int* counts) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, att_trace, sizeof(unsigned long long) * 4 * kTraceCap);
  cudaMemcpyFromSymbol(counts, att_trace_n, sizeof(int) * 4);
  int zero[4] = {0, 0, 0, 0};
  cudaMemcpyToSymbol(att_trace_n, zero, sizeof(zero));
  return kTraceCap;
}
#endif

int attention_init() {
  STAD_CUDA_OK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  return STAD_OK;
}

int launch_attention(const bf16* qkv, bf16* out, int B, int H, int S, float scale, cudaStream_t stream) {
  STAD_CHECK_ARG(B > 0 && H > 0 && S > 0, "attention: empty problem B=%d H=%d S=%d", B, H, S);
  if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15)
    return fail(STAD_E_ALIGN, "attention: qkv and out must be 16-byte aligned");
  const uint64_t row = static_cast<uint64_t>(3) * H * HD;  // elements per token in the packed projection
  const uint64_t dims[3] = {row, (uint64_t)S, (uint64_t)B};
  const uint64_t strides[2] = {row * 2, row * 2 * (uint64_t)S};
  const uint32_t box[3] = {HD, BQ, 1};
  CUtensorMap tm, tm32;
  int rc = make_tmap_bf16(&tm, qkv, 3, dims, strides, box);
  if (rc) return rc;
  const uint32_t box32[3] = {HD, SPLIT_ROWS, 1};  // Q rows of a key-split tail unit, loaded four times
  if ((rc = make_tmap_bf16(&tm32, qkv, 3, dims, strides, box32))) return rc;
  AttArgs a;
  a.out = out;
  a.B = B;
  a.H = H;
  a.S = S;
  a.scale_log2 = scale * 1.4426950408889634f;
  const int n_q = ceil_div(S, BQ);
  const long long total_units = static_cast<long long>(B) * H * ((n_q + 1) / 2);
  STAD_CHECK_ARG(total_units < (1ll << 30), "attention: B*H*tiles = %lld too large", total_units);
  const int grid = total_units < sm_count() ? static_cast<int>(total_units) : sm_count();
  ProfScope prof(STAD_K_ATTENTION, 0, B, H, S, stream);
  STAD_CUDA_OK(launch_pdl(attention_kernel, dim3(grid), dim3(ATT_THREADS), SMEM_BYTES, stream, 1, tm, tm32, a));
  STAD_LAUNCH_OK("attention_kernel");
  return STAD_OK;
}

}  // namespace stad
