// Fused joint space-time attention for sm_100a:  out = softmax(q k^T * scale) v  per (clip, head), head dim 64.
//
// Replaces Attention._naive_attn (modeling_finetune.py:93-103) and Attention._flash_attn -> FlashAttention.forward ->
// flash_attn_varlen_qkvpacked_func (modeling_finetune.py:121-128, flash_attention_class.py:39-51).  Input is the
// packed projection qkv[B, S, 3, H, 64] (the layout FlashAttention.forward takes); output is [B, S, H*64], the
// layout Attention.proj consumes.  The S x S score matrix never leaves the SM.
//
// Persistent kernel, one CTA per SM.  A work unit = TWO 128-row query tiles ("slots") of one (b, h) that share every
// K/V tile (96 keys) streamed through shared memory.  Head dim 64 makes the kernel exp-bound (one exp per 256 MMA
// flops), so everything is arranged around keeping the eight softmax warps busy:
//
//   warp 12      : TMA producer: Q tiles of the unit, then K and V tiles through two 5-stage smem rings.
//   warps 13, 14 : MMA issuers, one per slot.  S_j = Q K_j^T (SS, fp32, 96 TMEM columns), O += P_j V_j (A = P from
//                  TMEM, B = V as MN-major smem operand).  Every slot owns TWO score buffers; P_j is written by the
//                  softmax threads over the columns S_j was read from.  Issue order per slot
//                      QK_0, QK_1, PV_0, QK_2, PV_1, QK_3, ...
//                  so S_{j+1} is complete long before the softmax warps finish tile j: inside a unit they never wait
//                  for the tensor core.  Q K_{j+2}^T re-uses the buffer of S_j / P_j; it is issued after P_j V_j by
//                  the same thread and tcgen05.mma executes in issue order, so no barrier is needed between them.
//                  The tiles of consecutive units form ONE stream: the first two Q K^T of unit u+1 are issued behind
//                  the last two P V of unit u (Q tiles are double-buffered per slot for this), so the softmax warps
//                  find S_0 of the next unit complete when they finish a unit — with two K/V tiles per unit (S = 160,
//                  the masked encoder) the hand-over at the unit boundary was most of the kernel.
//   warps 0..3   : softmax of slot 0, one query row per thread: tcgen05.ld S_j, exp2(s*c - m) (FFMA2 + MUFU.EX2, one
//   warps 4..7   : softmax of slot 1   pair in four on the FMA pipe), row sum (FADD2), P_j (bf16) -> TMEM.
//                  LAZY REFERENCE MAX: the exact row max is taken on the first tile of a unit only.  Later tiles use
//                  that reference as it is (softmax is invariant to the reference; bf16 P and fp32 O / row sum keep
//                  their relative precision at any magnitude) and only watch the tile's row sum: if it leaves
//                  [0, 2^64) (a score more than ~64 octaves above the reference, or an overflow), the tile is redone
//                  with its exact max and O / the row sum are rescaled.  This takes the row-max pass (64 FMNMX3 and
//                  ~330 clk of latency per tile and warp, none of it overlappable with the warp's own MUFU work) out of
//                  the steady state.
//   warps 8..11  : epilogue: at the end of a unit they take the row sums from the softmax warps, read O, normalise
//                  and store it, so the softmax warps start the next unit immediately.
//
// TMEM (512 columns), slot s at column 256 s:  SP0 [0, 96)  SP1 [96, 192)  O [192, 256).
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
constexpr int BKV = 96;   // keys per tile
constexpr int KV_STAGES = 5;
constexpr int Q_BYTES = BQ * HD * 2;    // 16 KB
constexpr int KV_BYTES = BKV * HD * 2;  // 12 KB: every K / V tile
constexpr int ATT_THREADS = 512;  // warpgroups 0, 1: softmax slots; 2: epilogue; 3: TMA + 2 MMA issuers (+1 idle warp)
constexpr int LSUM_BYTES = 3 * BQ * 4;  // row sums handed from the softmax warps to the epilogue warps: [slot][row], then
                                        // the reference maxima of a key-split unit's (key part, query) pairs
// Key-split tail units (see Unit::split): the four partial O rows of a query are combined through this scratch,
// [key part][column][query] fp32 (query fastest: conflict-free for writers and readers)
constexpr int SPLIT_ROWS = 32;
constexpr int SPLIT_SCRATCH_BYTES = 4 * HD * SPLIT_ROWS * 4;
constexpr int Q_BUFS = 2;           // Q tiles per slot: unit u+1's Q is resident while unit u is still being multiplied
constexpr int SMEM_BYTES = 2 * Q_BUFS * Q_BYTES + 2 * KV_STAGES * KV_BYTES + LSUM_BYTES + 512 + SPLIT_SCRATCH_BYTES + 1024;
// Warp roles.  The single-thread TMA / MMA issuers sit in the HIGHEST warps: the sub-partition arbiter favours high warp
// ids, and an issuer that has to queue behind two always-ready softmax warps paces the whole kernel (measured: ~135
// clk per tcgen05.mma issue and ~350 clk per already-complete mbarrier wait when the issuers were warps 0-2).
constexpr int kTmaWarp = 12;   // warps 0-3: softmax slot 0, 4-7: softmax slot 1, 8-11: epilogue
constexpr int kMmaWarp0 = 13;  // 13: MMA issuer of slot 0, 14: of slot 1, 15: idle
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t SLOT_COLS = 256;  // TMEM columns per slot
constexpr uint32_t O_COL = 2 * BKV;  // within the slot; the score / P buffers are at 0 and BKV
// a tile whose row sum (relative to the lazy reference) reaches this is redone with its exact max (see header)
constexpr float kSumGuard = 18446744073709551616.f;  // 2^64
constexpr float kRescaleThreshold = 8.0f;            // log2 units; key-split tail units only (exact running max)

// -DSTAD_ATT_TRACE: CTA 0 records (clock, tag) events of one softmax warp per slot and of the two MMA issuers into
// att_trace (development builds only; read through stad_debug_read_att_trace, see tools/att_trace.py).
#ifdef STAD_ATT_TRACE
constexpr int kTraceCap = 2048;
__device__ unsigned long long att_trace[6][kTraceCap];
__device__ int att_trace_n[6];
#define ATT_EV(role, tag)                                                                    \
  do {                                                                                       \
    if (blockIdx.x == 0 && tr_n < kTraceCap && (threadIdx.x & 31) == 0) {                    \
      att_trace[role][tr_n] = (static_cast<unsigned long long>(clock64()) << 8) | (tag);     \
      att_trace_n[role] = ++tr_n;                                                            \
    }                                                                                        \
  } while (0)
#define ATT_T(i) do { if (quarter == 0) ATT_EV(slot, i); } while (0)
#else
#define ATT_EV(role, tag) do {} while (0)
#define ATT_T(i) do {} while (0)
#endif

struct AttArgs {
  bf16* out;
  int B, H, S;
  float scale_log2;  // softmax scale * log2(e)
};

// A one-slot unit whose query tile holds at most 32 rows (the ragged end of the sequence: rows 1536..1567 of 1568)
// would keep ONE warp busy per tile while costing a whole unit's worth of MMAs and latencies (measured: 11 % of the
// kernel for 2 % of the rows).  It runs "key-split" instead: the 32 queries are replicated into all four 32-row groups
// of the Q tile, so every TMEM lane quarter holds their scores against all 96 keys of a tile, and the softmax warp of
// quarter k < 3 handles only keys [32k, 32k+32) of each tile with its own running max / sum (its P entries for the
// other keys are zero; quarter 3 contributes nothing).  The partial outputs of a query are merged by the epilogue warps.
struct Unit {
  int b, h, q0, slots;
  bool split;
};

// the part of a unit that depends only on its position t within the head (all the softmax warps and the MMA issuers
// need; they step t incrementally instead of dividing at every unit boundary, which is on their critical path)
STAD_DEVICE Unit unit_shape(int t, int S) {
  Unit w;
  w.b = w.h = 0;
  w.q0 = t * 2 * BQ;
  w.slots = (w.q0 + BQ < S) ? 2 : 1;
  w.split = w.slots == 1 && S - w.q0 <= SPLIT_ROWS;
  return w;
}

STAD_DEVICE Unit decode_unit(int u, int units_per_head, int H, int S) {
  const int bh = u / units_per_head;
  Unit w = unit_shape(u - bh * units_per_head, S);
  w.b = bh / H;
  w.h = bh - w.b * H;
  return w;
}

STAD_DEVICE void mask_from(uint32_t (&t)[32], int valid) {  // columns >= valid -> -inf
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i >= valid) t[i] = 0xFF800000u;
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                 const __grid_constant__ CUtensorMap tmap_q32, const AttArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_q = smem;                                   // [2 slots][Q_BUFS][16 KB]
  uint8_t* smem_k = smem_q + 2 * Q_BUFS * Q_BYTES;          // [KV_STAGES][12 KB]
  uint8_t* smem_v = smem_k + KV_STAGES * KV_BYTES;          // [KV_STAGES][12 KB]
  float* lsum_smem = reinterpret_cast<float*>(smem_v + KV_STAGES * KV_BYTES);  // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + KV_STAGES * KV_BYTES + LSUM_BYTES);
  uint64_t* q_full = bars;                   // [2 slots][Q_BUFS]  TMA -> MMA
  uint64_t* q_free = q_full + 2 * Q_BUFS;    // [2 slots][Q_BUFS]  MMA (last Q K^T of the unit) -> TMA
  uint64_t* k_full = q_free + 2 * Q_BUFS;    // [KV_STAGES]
  uint64_t* k_free = k_full + KV_STAGES;     // [KV_STAGES]
  uint64_t* v_full = k_free + KV_STAGES;     // [KV_STAGES]
  uint64_t* v_free = v_full + KV_STAGES;     // [KV_STAGES]
  uint64_t* s_full = v_free + KV_STAGES;     // [2 slots][2 buffers]  MMA -> softmax: S_j complete
  uint64_t* p_full = s_full + 4;             // [2 slots][2 buffers]  softmax (4 warps) -> MMA: P_j stored over S_j
  uint64_t* o_full = p_full + 4;             // [2]  MMA -> softmax / epilogue: P_j V_j complete
  uint64_t* l_ready = o_full + 2;            // [2]  softmax (128 threads) -> epilogue: unit done, row sums in smem
  uint64_t* o_free = l_ready + 2;            // [2]  epilogue (4 warps) -> MMA: O has been read, next unit may overwrite
  uint64_t* o_done = o_free + 2;             // [2]  MMA -> epilogue: the LAST P V of a unit is complete (one phase per unit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);
  float* split_scratch = reinterpret_cast<float*>(smem_v + KV_STAGES * KV_BYTES + LSUM_BYTES + 512);

  // warp index through a shuffle: the compiler then knows every value derived from it is warp-uniform (uniform
  // registers feed tcgen05.mma / TMA directly instead of a per-lane ELECT + R2UR.BROADCAST loop per instruction)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n_q = (p.S + BQ - 1) / BQ;
  const int units_per_head = (n_q + 1) / 2;
  const int total_units = p.B * p.H * units_per_head;
  const int n_kv = (p.S + BKV - 1) / BKV;
  // position within the head of this CTA's first unit, and of unit u + gridDim.x given that of unit u
  const int t_first = static_cast<int>(blockIdx.x) % units_per_head;
  const int t_step = static_cast<int>(gridDim.x) % units_per_head;
  auto next_t = [&](int t) {
    t += t_step;
    return t >= units_per_head ? t - units_per_head : t;
  };
  // The keys beyond the last multiple of 96 form a short ("ragged") tile.  It is processed FIRST (tile 0 of every
  // unit), the full tiles follow: the first tile of a unit is special anyway (it takes the exact row max that becomes
  // the lazy reference, with all its exps on the MUFU), and a separate ragged tile at the end costs nearly a full
  // tile's latency chain for a third of the work (S = 1568: 32 keys).  Folding the two special tiles into one short
  // first tile leaves 16 steady-state tiles per unit.
  const int last_valid = p.S - (n_kv - 1) * BKV;   // valid keys of the ragged tile, 1..96 (96: no ragged tile)
  const int last_chunks = (last_valid + 31) >> 5;  // its 32-column chunks that hold any valid key
  // first key of tile j
  auto kv_row = [&](int j) { return last_valid == BKV ? j * BKV : (j == 0 ? (n_kv - 1) * BKV : (j - 1) * BKV); };

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    tma_prefetch_desc(&tmap_q32);
    for (int s = 0; s < 2 * Q_BUFS; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_free[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&o_full[s], 1);
      mbar_init(&l_ready[s], BQ);
      mbar_init(&o_free[s], BQ);  // one arrive per epilogue thread
      mbar_init(&o_done[s], 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], BQ);  // one arrive per softmax thread
    }
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_free[s], 2);  // one arrive per slot
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
      // Tile order K_0, K_1, V_0, K_2, V_1, ... = the order in which the issuers consume them.
      uint32_t kst = 0, kph = 0, vst = 0, vph = 0;
      uint32_t ucnt0 = 0, ucnt1 = 0;  // units started per slot (phase of q_full / q_free)
      auto load_tile = [&](const CUtensorMap* map, uint64_t* full, uint8_t* dst, uint32_t bytes, int col, int row, int b) {
        if (elect_one()) {
          mbar_arrive_expect_tx(full, bytes);
          tma_load_3d(dst, map, full, col, row, b);
        }
        __syncwarp();
      };
      auto load_k = [&](int j, int b, int col_k) {
        mbar_wait(&k_free[kst], kph ^ 1);
        load_tile(&tmap_kv, &k_full[kst], smem_k + kst * KV_BYTES, KV_BYTES, col_k, kv_row(j), b);
        if (++kst == KV_STAGES) {
          kst = 0;
          kph ^= 1;
        }
      };
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const Unit w = decode_unit(u, units_per_head, p.H, p.S);
        const int col_q = w.h * HD;
        const int col_k = (p.H + w.h) * HD;
        const int col_v = (2 * p.H + w.h) * HD;
        // Q tile of slot s for its n-th unit -> buffer n % Q_BUFS (free once the last Q K^T of unit n - Q_BUFS is done)
        const uint32_t qb0 = ucnt0 % Q_BUFS;
        uint8_t* q0_dst = smem_q + qb0 * Q_BYTES;
        mbar_wait(&q_free[qb0], ((ucnt0 / Q_BUFS) & 1) ^ 1);
        if (w.split) {
          // the same 32 query rows into each of the four 32-row groups of the tile (4 KB apart in the swizzled layout)
          if (elect_one()) {
            mbar_arrive_expect_tx(&q_full[qb0], Q_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tma_load_3d(q0_dst + k * (SPLIT_ROWS * HD * 2), &tmap_q32, &q_full[qb0], col_q, w.q0, w.b);
          }
          __syncwarp();
        } else {
          load_tile(&tmap_q, &q_full[qb0], q0_dst, Q_BYTES, col_q, w.q0, w.b);
        }
        ++ucnt0;
        if (w.slots > 1) {
          const uint32_t qb1 = ucnt1 % Q_BUFS;
          mbar_wait(&q_free[Q_BUFS + qb1], ((ucnt1 / Q_BUFS) & 1) ^ 1);
          load_tile(&tmap_q, &q_full[Q_BUFS + qb1], smem_q + (Q_BUFS + qb1) * Q_BYTES, Q_BYTES, col_q, w.q0 + BQ, w.b);
          ++ucnt1;
        }
        load_k(0, w.b, col_k);
        for (int j = 0; j < n_kv; ++j) {
          if (j + 1 < n_kv) load_k(j + 1, w.b, col_k);
          mbar_wait(&v_free[vst], vph ^ 1);
          load_tile(&tmap_kv, &v_full[vst], smem_v + vst * KV_BYTES, KV_BYTES, col_v, kv_row(j), w.b);
          if (++vst == KV_STAGES) {
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
      // The K/V tiles of all units of this slot are ONE stream t = 0, 1, 2, ...: tile t uses score buffer t & 1, and
      // Q K_{t+2}^T is issued right behind P_t V_t, also when tile t + 2 belongs to the NEXT unit (its first two score
      // tiles are then complete when the softmax warps get there).
      // This warp sits on the hand-over P_t stored -> P_t V_t -> Q K_{t+2}^T -> S_{t+2}, which has little slack (one
      // more already-satisfied barrier test per tile costs 1 - 2 % of the kernel, measured): its per-tile path is kept
      // short — ring positions and phases are stepped, not divided; no per-tile unit bookkeeping.
      const int slot = warp - kMmaWarp0;
#ifdef STAD_ATT_TRACE
      int tr_n = 0;
#endif
      constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, HD, 0, 1);   // P from TMEM (K-major), V MN-major (d contiguous)
      constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0, 0);  // Q, K both K-major
      const uint32_t idesc_qk_ragged = make_idesc_bf16(BQ, static_cast<uint32_t>(last_chunks * 32), 0, 0);
      const int ksteps_ragged = last_chunks * 2;  // (no ragged tile: = BKV / 16)
      const bool has_ragged = last_valid != BKV;
      const uint64_t desc_q_base = make_smem_desc_sw128(smem_u32(smem_q + slot * Q_BUFS * Q_BYTES), 16, 1024);
      const uint64_t desc_k_base = make_smem_desc_sw128(smem_u32(smem_k), 16, 1024);
      const uint64_t desc_v_base = make_smem_desc_sw128(smem_u32(smem_v), 0, 1024);
      const uint32_t tmem_slot_base = tmem_base + slot * SLOT_COLS;
      const uint32_t tmem_o = tmem_slot_base + O_COL;
      uint64_t* const s_full_s = s_full + slot * 2;
      uint64_t* const p_full_s = p_full + slot * 2;
      uint64_t* const q_full_s = q_full + slot * Q_BUFS;
      uint64_t* const q_free_s = q_free + slot * Q_BUFS;
      // K ring position of the next Q K^T, V ring position of the next P V (stage, phase); tiles issued so far
      uint32_t kst = 0, kph = 0, vst = 0, vph = 0;
      uint32_t tq = 0, tp = 0;
      // a unit this slot sits out advances a ring position by n_kv tiles
      const uint32_t skip_st = static_cast<uint32_t>(n_kv) % KV_STAGES, skip_ph = (static_cast<uint32_t>(n_kv) / KV_STAGES) & 1u;
      auto skip_unit = [&](uint32_t& st, uint32_t& ph) {
        st += skip_st;
        ph ^= skip_ph;
        if (st >= KV_STAGES) {
          st -= KV_STAGES;
          ph ^= 1u;
        }
      };
      // S = Q K_j^T of unit number n (of this slot) into score buffer tq & 1; releases the K stage, and the Q tile
      // after the unit's last tile
      auto issue_qk = [&](uint32_t n, bool first, bool last, bool solo) {
        const uint32_t qb = n % Q_BUFS;
        if (first) mbar_wait(&q_full_s[qb], (n / Q_BUFS) & 1);
        mbar_wait(&k_full[kst], kph);
        tc_fence_after();
        const uint32_t buf = tq & 1;
        ATT_EV(2 + slot, 20);
        if (elect_one()) {
          const uint64_t desc_q = desc_q_base + static_cast<uint64_t>(qb * (Q_BYTES >> 4));
          const uint64_t desc_k = desc_k_base + static_cast<uint64_t>(kst * (KV_BYTES >> 4));
          const uint32_t tmem_s = tmem_slot_base + buf * BKV;
          const uint32_t idesc = (first && has_ragged) ? idesc_qk_ragged : idesc_qk;  // ragged tile: only its 32-key chunks
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_ss(tmem_s, desc_q + 2 * k, desc_k + 2 * k, idesc, k != 0);
          umma_commit(&s_full_s[buf]);
          if (last) umma_commit(&q_free_s[qb]);
          umma_commit(&k_free[kst]);
          if (solo) umma_commit(&k_free[kst]);
        }
        __syncwarp();
        ATT_EV(2 + slot, 21);
        ++tq;
        if (++kst == KV_STAGES) {
          kst = 0;
          kph ^= 1u;
        }
      };
      // O (+)= P_j V_j of unit number n.  V tile: one 128-byte row per key -> MN-major B operand; 16 keys =
      // 2 x 1024 B per MMA.
      auto issue_pv = [&](uint32_t n, bool first, bool last, bool solo) {
        mbar_wait(&v_full[vst], vph);
        if (first && n > 0) mbar_wait(&o_free[slot], (n - 1) & 1);  // epilogue has read the previous unit's O
        // The previous P V of this slot, issued a tile ago, is complete.  Observing its o_full phase here, every tile
        // and BEFORE the wait for P, means the barrier never runs a phase ahead of an observer: the parity waits of
        // the softmax warps' slow path, its only other waiters and rare, cannot alias, and compute-sanitizer's
        // synccheck (which rejects a barrier that completes phase after phase unobserved) stays clean.
        if (tp > 0) mbar_wait(&o_full[slot], (tp - 1) & 1);
        const uint32_t buf = tp & 1;
        ATT_EV(2 + slot, 22);
        mbar_wait(&p_full_s[buf], (tp >> 1) & 1);
        tc_fence_after();
        ATT_EV(2 + slot, 23);
        if (elect_one()) {
          const uint64_t desc_v = desc_v_base + static_cast<uint64_t>(vst * (KV_BYTES >> 4));
          const uint32_t tmem_p = tmem_slot_base + buf * BKV;
          if (!first) {
#pragma unroll
            for (int k = 0; k < BKV / 16; ++k) umma_ts(tmem_o, tmem_p + k * 8, desc_v + (k * 2048 >> 4), idesc_pv, 1u);
          } else {  // first tile of the unit: starts the accumulation; the ragged tile if there is one
            for (int k = 0; k < ksteps_ragged; ++k)
              umma_ts(tmem_o, tmem_p + k * 8, desc_v + (k * 2048 >> 4), idesc_pv, k != 0 ? 1u : 0u);
          }
          umma_commit(&o_full[slot]);
          if (last) umma_commit(&o_done[slot]);
          umma_commit(&v_free[vst]);
          if (solo) umma_commit(&v_free[vst]);
        }
        __syncwarp();
        ATT_EV(2 + slot, 24);
        ++tp;
        if (++vst == KV_STAGES) {
          vst = 0;
          vph ^= 1u;
        }
      };

      // Unit list of this CTA: blockIdx.x, + gridDim.x, ...; `ut` = position within the head.
      int u = blockIdx.x, ut = t_first;
      // moves (u, ut) to the first unit from (u, ut) on that this slot takes part in; counts the units it skips
      auto seek = [&](int& uu, int& uut, int& skipped, bool& solo) {
        skipped = 0;
        while (uu < total_units) {
          const Unit w = unit_shape(uut, p.S);
          if (slot < w.slots) {
            solo = w.slots == 1;
            return true;
          }
          ++skipped;
          uu += gridDim.x;
          uut = next_t(uut);
        }
        return false;
      };
      int skipped = 0;
      bool solo = false;
      bool have = seek(u, ut, skipped, solo);
      for (int i = 0; i < skipped; ++i) {
        skip_unit(kst, kph);
        skip_unit(vst, vph);
      }
      int pre = 0;      // Q K^T of the current unit already issued (by the look-ahead from the unit before)
      uint32_t n = 0;   // number of this unit among the units of this slot
      while (have) {
        // the next unit of this slot.  Look-ahead NOT across units this slot sits out: their K tiles come first in the
        // ring and are loaded only as V stages become free, and the V stages of the current unit are released by P V
        // this warp would be holding back while it waits for the K tile behind the gap (deadlock).  Behind a gap the
        // slot finishes its unit first, as it would without look-ahead.
        int u2 = u + gridDim.x, ut2 = next_t(ut), skipped2 = 0;
        bool solo2 = false;
        const bool have2 = seek(u2, ut2, skipped2, solo2);
        const bool look = have2 && skipped2 == 0;
        for (; pre < 2 && pre < n_kv; ++pre) issue_qk(n, pre == 0, pre + 1 == n_kv, solo);
        int pre2 = 0;
        for (int j = 0; j < n_kv; ++j) {
          issue_pv(n, j == 0, j + 1 == n_kv, solo);
          const int jq = j + 2;
          if (jq < n_kv) {
            issue_qk(n, false, jq + 1 == n_kv, solo);
          } else if (look && jq - n_kv == pre2) {  // (false with one tile per unit: the stream index is two units on)
            issue_qk(n + 1, pre2 == 0, pre2 + 1 == n_kv, solo2);
            ++pre2;
          }
        }
        for (int i = 0; i < skipped2; ++i) {  // (look-ahead implies skipped2 == 0)
          skip_unit(kst, kph);
          skip_unit(vst, vph);
        }
        pre = pre2;
        u = u2;
        ut = ut2;
        solo = solo2;
        have = have2;
        ++n;
      }
    }
#ifdef STAD_ATT_TRACE
    else {  // warp 15 (idle in product builds): when does each S tile of slot 0 really complete?
      int tr_n = 0;
      uint32_t t = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x)
        for (int j = 0; j < n_kv; ++j, ++t) {
          mbar_wait(&s_full[t & 1], (t >> 1) & 1);
          ATT_EV(4, 30 + (t & 1));
        }
    }
#endif
  } else if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    // ------------------------------------------------------------------ epilogue warps: O / l -> bf16 -> global
    // Takes the end-of-unit work off the softmax warps: they hand over the row sums and move on to the next unit.
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint32_t ucs[2] = {0, 0};   // units completed per slot
#ifdef STAD_ATT_TRACE
    int tr_n = 0;
#define ATT_E(i) do { if (quarter == 0 && slot == 0) ATT_EV(5, i); } while (0)
#else
#define ATT_E(i) do {} while (0)
#endif
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const Unit w = decode_unit(u, units_per_head, p.H, p.S);
#pragma unroll
      for (int slot = 0; slot < 2; ++slot) {
        if (slot >= w.slots) continue;
        // every softmax thread of the slot has finished the unit, and its last P V is complete.  (o_done, not the
        // per-tile o_full: the softmax warps can be through with a short unit before its FIRST P V has completed, and a
        // parity wait for tile n while tile n - 1 is still pending falls through.)
        ATT_E(39);
        mbar_wait(&l_ready[slot], ucs[slot] & 1);
        ATT_E(40);
        mbar_wait(&o_done[slot], ucs[slot] & 1);
        ATT_E(41);
        tc_fence_after();
        ++ucs[slot];
        const uint32_t o_addr = lane_addr + slot * SLOT_COLS + O_COL;
        if (w.split) {
          // ---- merge the key-part partials of each of the 32 queries:
          //   out = sum_k 2^(m_k - M) O_k / sum_k 2^(m_k - M) l_k,  M = max_k m_k
          float mk[4], M = -INFINITY;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            mk[k] = lsum_smem[2 * BQ + k * 32 + lane];
            M = fmaxf(M, mk[k]);
          }
          float L = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) L += (mk[k] == -INFINITY ? 0.f : ex2(mk[k] - M)) * lsum_smem[k * 32 + lane];
          const float wgt = (mk[quarter] == -INFINITY ? 0.f : ex2(mk[quarter] - M)) / L;
          named_bar_sync(2, 128);  // the previous split unit's combine has finished reading the scratch
          float* mine = split_scratch + quarter * (HD * SPLIT_ROWS) + lane;
#pragma unroll 1
          for (int q = 0; q < 4; ++q) {  // 16 columns at a time: these warps run with 56 registers
            uint32_t ov[16];
            tmem_ld16(o_addr + q * 16, ov);
            tmem_ld_wait16(ov);
            if (q == 3) {  // O is in registers: the next unit's first P V may overwrite it
              tc_fence_before();
              mbar_arrive(&o_free[slot]);
            }
            // a part that never saw a key (weight 0) holds whatever its all-zero P rows produced: exactly 0
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
            mbar_arrive(&o_free[slot]);
            ATT_E(42);
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
    const uint32_t slot_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + slot * SLOT_COLS;
    const uint32_t o_addr = slot_addr + O_COL;
    const float c = p.scale_log2;
    uint32_t g = 0;  // tiles completed by this slot (buffer g & 1, its (g >> 1)-th use; o_full phase g)
    uint32_t un = 0;  // units completed by this slot
#ifdef STAD_ATT_TRACE
    int tr_n = 0;
#endif
    // End of a unit: row sum (and, key-split units, the reference max) to the epilogue warps.  The scores of the next
    // unit are computed ahead of time (see the MMA issuers), so with few K/V tiles per unit these warps can finish
    // unit n before the epilogue warps have taken unit n - 1: wait until they have (o_free: O and the row sums read).
    // (With three or more tiles per unit the phase is always complete by now: S of the unit's last tile exists only
    // after the issuer has started the unit's first P V, for which it waited on the same o_free phase itself.)
    auto hand_over = [&](float l, float m, bool with_max) {
      if (un > 0) mbar_wait(&o_free[slot], (un - 1) & 1);
      lsum_smem[slot * BQ + r] = l;
      if (with_max) lsum_smem[2 * BQ + r] = m;
      mbar_arrive(&l_ready[slot]);
      ++un;
    };

    // Rescale of this thread's O row (rare: slow path of the lazy reference / growth in a key-split unit).
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
    // Barrier addresses as plain 32-bit shared addresses, computed once.
    const uint32_t s_full_a = smem_u32(&s_full[slot * 2]);
    const uint32_t p_full_a = smem_u32(&p_full[slot * 2]);
    // My P columns of this tile are in TMEM: let the issuer run P V.  EVERY thread arrives (the barrier counts the 128
    // threads of the slot): no warp barrier, no elected lane, no branch on the way out of the tile.
    auto publish_p = [&](uint32_t p_bar) {
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_a(p_bar);
    };

    for (int u = blockIdx.x, ut = t_first; u < total_units; u += gridDim.x, ut = next_t(ut)) {
      const Unit w = unit_shape(ut, p.S);
      if (slot >= w.slots) continue;
      if (w.split) {
        // ---- key-split tail unit (slot 0 only): lane i of EVERY quarter holds query q0 + i; the warp of quarter k < 3
        // handles the keys [32 k, +32) of each K/V tile (one 32-column chunk of S, one 16-column chunk of P) with an
        // exact running max; its other P columns are zero.
        float m_ref = -INFINITY, l_sum = 0.f;
        for (int j = 0; j < n_kv; ++j, ++g) {
          const uint32_t buf = g & 1;
          const uint32_t sp = slot_addr + buf * BKV;
          const int valid = (j == 0 ? last_valid : BKV) - quarter * 32;  // keys of this tile in my chunk
          mbar_wait(&s_full[slot * 2 + buf], (g >> 1) & 1);
          tc_fence_after();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = 0u;
          if (valid > 0) {
            uint32_t t[32];
            tmem_ld32(sp + quarter * 32, t);
            tmem_ld_wait32(t);
            if (valid < 32) mask_from(t, valid);
            const float mx = c * chunk_max(t);
            const bool grow = mx > m_ref + kRescaleThreshold;  // always true on the first tile with keys (m_ref = -inf)
            if (__any_sync(0xffffffffu, grow)) {
              const float alpha = (grow && m_ref != -INFINITY) ? ex2(m_ref - mx) : 1.0f;
              if (grow) {
                m_ref = mx;
                l_sum *= alpha;
              }
              if (j > 0) {
                mbar_wait(&o_full[slot], (g - 1) & 1);  // P V of the previous tile: O stable
                tc_fence_after();
                rescale_o(alpha);
              }
            }
            float a0 = 0.f, a1 = 0.f;
            exp_chunk<false>(t, c, -m_ref, a0, a1, pk);
            l_sum += a0 + a1;
          }
          // all 48 P columns of my lanes: my chunk, zeros elsewhere (the buffer held this tile's scores)
          uint32_t z[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            if (q == quarter) tmem_st16(sp + q * 16, pk);
            else tmem_st16(sp + q * 16, z);
          }
          publish_p(p_full_a + buf * 8);
        }
        // partial sum of (query lane, key part) and its reference max (-inf: this part never saw a key)
        hand_over(l_sum, m_ref, true);
        continue;
      }
      const int row0 = w.q0 + slot * BQ;
      // Phase offset between the two slots, re-established at the first tile of every two-slot unit: without it both
      // softmax warpgroups drift into lockstep (same phase of the tile at the same time), i.e. they fight for the
      // MUFU together and idle together.  Slot 1 starts when slot 0 has stored the first third of its first P tile.
      // (Only with >= KV_STAGES tiles per unit: the shared K ring then keeps slot 0 from reaching the next unit's arrive
      // before slot 1 has passed this unit's sync, so the arrive / sync counts of the named barrier cannot interleave.)
      const bool stagger = w.slots == 2 && n_kv >= KV_STAGES;
      // loop-carried tile state, toggled instead of re-derived from g (keeps the per-tile glue short): score buffer
      // address (the two buffers differ in one address bit pattern: BKV = 0x60 and the slot base has those bits clear),
      // barrier addresses (8 bytes apart, 16-byte aligned pairs), barrier phase
      uint32_t buf = g & 1;
      uint32_t sp = slot_addr + buf * BKV;
      uint32_t s_bar = s_full_a + buf * 8, p_bar = p_full_a + buf * 8;
      uint32_t ph = (g >> 1) & 1;
      // All four warps of slot 1 wait at this ONE bar.sync, whether their rows are valid or not (slot 0 of a two-slot
      // unit has no invalid rows, and arrives from the first tile below).
      if (stagger && slot == 1) named_bar_sync(1, 2 * BQ);
      if (row0 + quarter * 32 >= p.S) {
        // rows beyond S (warp-uniform): nothing to compute (their P / O rows are never stored); keep the pipeline
        // moving.  (The previous phase of p_full[buf] is complete: Q K_j^T was issued after P_{j-2} V_{j-2}.)
        for (int j = 0; j < n_kv; ++j, ++g) {
          mbar_wait_a(s_bar, ph);
          mbar_arrive_a(p_bar);
          ph ^= buf;
          buf ^= 1;
          s_bar ^= 8;
          p_bar ^= 8;
        }
        hand_over(1.f, 0.f, false);
        continue;
      }
      float m_ref = 0.f;  // reference max (log2 domain, already scaled): exact max of the unit's first tile
      float l_sum = 0.f;
      // Loop bounds in ordinary registers (laundered through an empty asm): kept in uniform registers the compiler
      // re-derives them from the kernel parameters at the top of EVERY tile (a chain of ~10 dependent uniform-datapath
      // instructions and a constant-bank load on the critical path of the tile).
      int n_tiles = n_kv;
      asm volatile("" : "+r"(n_tiles));

      for (int j = 0; j < n_tiles; ++j, ++g, ph ^= buf, buf ^= 1, sp ^= BKV, s_bar ^= 8, p_bar ^= 8) {
        ATT_T(7);
        mbar_wait_a(s_bar, ph);
        tc_fence_after();
        ATT_T(0);
        uint32_t sv[3][32];
        float t_sum;
        if (j != 0) {
          // ---- steady state: full tile, lazy reference.  The second and third chunk are in flight while the first
          // is processed; P chunk i overwrites columns [16 i, +16), which lie in S chunk i / 2 (already in registers).
          const float neg_m = -m_ref;
          float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
          uint32_t pk[16];
          tmem_ld32(sp, sv[0]);
          tmem_ld_wait32(sv[0]);
          tmem_ld32(sp + 32, sv[1]);
          tmem_ld32(sp + 64, sv[2]);
          exp_chunk<true>(sv[0], c, neg_m, a0, a1, pk);
          tmem_st16(sp, pk);
          tmem_ld_wait32(sv[1]);
          tmem_ld_wait32(sv[2]);
          ATT_T(1);
          exp_chunk<true>(sv[1], c, neg_m, b0, b1, pk);
          tmem_st16(sp + 16, pk);
          exp_chunk<true>(sv[2], c, neg_m, a0, a1, pk);
          tmem_st16(sp + 32, pk);
          t_sum = (a0 + a1) + (b0 + b1);
          ATT_T(3);
        } else {
          // ---- first tile of the unit: the ragged tile if there is one (masked, fewer chunks); exact max -> reference
          const int ncols = last_valid;  // valid keys of this tile
          const int nch = (ncols + 31) >> 5;
#pragma unroll
          for (int q = 0; q < 3; ++q)
            if (q < nch) tmem_ld32(sp + q * 32, sv[q]);
          float mx = -INFINITY;
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            if (q < nch) {
              tmem_ld_wait32(sv[q]);
              if (ncols - q * 32 < 32) mask_from(sv[q], ncols - q * 32);
              mx = fmaxf(mx, chunk_max(sv[q]));
            }
          }
          m_ref = c * mx;
          const float neg_m = -m_ref;
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            if (q < nch) {
              uint32_t pk[16];
              exp_chunk<false>(sv[q], c, neg_m, a0, a1, pk);
              tmem_st16(sp + q * 16, pk);
              if (q == 0 && stagger && slot == 0) asm volatile("bar.arrive 1, %0;" ::"n"(2 * BQ) : "memory");
            }
          }
          t_sum = a0 + a1;
        }
        if (j > 0) {
          const bool bad = !(t_sum < kSumGuard);
          if (__any_sync(0xffffffffu, bad)) {
            // ---- slow path (rare): some row of this warp met a score far above its reference.  Move that row's
            // reference to the exact max of this tile, rescale its O row and row sum, redo the tile.
            const float mx = c * max3(chunk_max(sv[0]), chunk_max(sv[1]), chunk_max(sv[2]));
            const float alpha = bad ? ex2(m_ref - mx) : 1.0f;  // 0 when the gap exceeds the fp32 range: O, l are then negligible
            if (bad) {
              m_ref = mx;
              l_sum *= alpha;
            }
            mbar_wait(&o_full[slot], (g - 1) & 1);  // P V of the previous tile: O stable
            tc_fence_after();
            rescale_o(alpha);
            const float neg_m = -m_ref;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int q = 0; q < 3; ++q) {  // absent / masked columns are -inf: their P entries are 0
              uint32_t pk[16];
              exp_chunk<false>(sv[q], c, neg_m, a0, a1, pk);
              tmem_st16(sp + q * 16, pk);
            }
            t_sum = a0 + a1;
          }
        }
        l_sum += t_sum;
        publish_p(p_bar);
        ATT_T(5);
      }
      // ---- hand the row sum to the epilogue warps and go on with the next unit
      hand_over(l_sum, 0.f, false);
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
  cudaMemcpyFromSymbol(out, att_trace, sizeof(unsigned long long) * 6 * kTraceCap);
  cudaMemcpyFromSymbol(counts, att_trace_n, sizeof(int) * 6);
  int zero[6] = {0, 0, 0, 0, 0, 0};
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
  CUtensorMap tm_q, tm_kv, tm_q32;
  const uint32_t box_q[3] = {HD, BQ, 1};
  const uint32_t box_kv[3] = {HD, BKV, 1};
  const uint32_t box_q32[3] = {HD, SPLIT_ROWS, 1};  // Q rows of a key-split tail unit, loaded four times
  int rc = make_tmap_bf16(&tm_q, qkv, 3, dims, strides, box_q);
  if (rc) return rc;
  if ((rc = make_tmap_bf16(&tm_kv, qkv, 3, dims, strides, box_kv))) return rc;
  if ((rc = make_tmap_bf16(&tm_q32, qkv, 3, dims, strides, box_q32))) return rc;
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
  STAD_CUDA_OK(launch_pdl(attention_kernel, dim3(grid), dim3(ATT_THREADS), SMEM_BYTES, stream, 1, tm_q, tm_kv, tm_q32, a));
  STAD_LAUNCH_OK("attention_kernel");
  return STAD_OK;
}

}  // namespace stad
