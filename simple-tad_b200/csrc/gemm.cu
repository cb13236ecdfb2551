// Persistent warp-specialised tcgen05 GEMM for sm_100a:  out[M,N] = epilogue(A[M,K] . W[N,K]^T)
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      : MMA issuer     (one elected thread, tcgen05.mma cta_group::1 kind::f16, M=128 x N=BN x K=16)
//                 + TMEM owner   (2 accumulator stages so the epilogue of tile i overlaps the mainloop of tile i+1)
//   warps 2..9  : epilogue       two warpgroups, each owning half of the BN columns: tcgen05.ld 32x32b -> registers ->
//                                bias / LN-fold / GELU / pos -> (+ residual sub-tile fetched by TMA into the staging
//                                buffer) -> bf16 -> 64B-swizzled smem staging -> TMA store (128 rows x 32 columns)
//
// Replaces every nn.Linear / F.linear on the reference path (modeling_finetune.py:48,52,92,104,119,128) and, in
// patch mode, the Conv3d of PatchEmbed (modeling_finetune.py:181-190): the A operand is then fetched with a 5-D tensor
// map straight out of the [.., H, W] clip planes (im2col-free), k order (c, dt, dh, dw) = Conv3d weight order.  One
// TMA box = the 16 dw pixels of one (c, dt, dh) line for every token of the tile = one UMMA K-slice, stored as
// 32-byte rows (SWIZZLE_32B, so the box is dense); four such boxes make one 64-wide k chunk.
//
// Roofline: dense BF16 tensor. Algorithmic FLOPs = 2 M N K; HBM bytes = 2 (M K + N K + M N) (+ 2 M N residual).
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace stad {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 8;
constexpr int kMaxFoldParts = kGemmMaxFoldParts;  // see kernels.h
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kSubCols = 32;                      // columns per epilogue sub-tile = one tcgen05.ld.x32 = one TMA store
constexpr int kSubTileBytes = BM * kSubCols * 2;  // 8 KB: 128 rows x 64 B, SWIZZLE_64B
constexpr int kWarpSlots = 4;                      // pair tile: staging slots per epilogue warp
constexpr int kWarpSlotBytes = 32 * kSubCols * 2;  // 2 KB: the 32 rows of one warp x 64 B

// kPair: the tile is 256 x BN, computed by the two CTAs of a cluster (one TPC) with tcgen05.mma.cta_group::2.  Each CTA
// stages its own 128 A rows and HALF of the B rows (the MMA reads the other half from the peer's shared memory), so
// per flop a CTA moves 2/3 of the operand bytes of the single-CTA tile through L2 -> smem -> tensor core.
template <int BN, bool kPair = false>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_ROWS = kPair ? BN / 2 : BN;   // B rows staged by this CTA
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = kPair ? 5 : BN == 256 ? 4 : BN == 192 ? 4 : BN == 128 ? 6 : 8;
  static constexpr int ACC_STRIDE = BN <= 64 ? 64 : BN <= 128 ? 128 : 256;  // TMEM columns per accumulator stage
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int BAR_BYTES = kPair ? 512 : 256;
  // single-CTA tile: 2 warpgroups x 2 buffers x (128 rows x 64 B), stored by the warpgroup's lead warp;
  // pair tile: every epilogue warp owns a ring of kWarpSlots (32 rows x 64 B) slots and issues its own TMA stores
  static constexpr int READY_BARS = kPair ? kEpiWarps * kWarpSlots : 4;
  static constexpr int STAGING_BYTES = kPair ? kEpiWarps * kWarpSlots * kWarpSlotBytes : 2 * 2 * kSubTileBytes;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB shared memory of one CTA");
};

struct KArgs {
  int M, N, K;
  int m_tiles, n_tiles;
  const float* bias;
  const float* colsum;
  const float2* stats;
  const float2* stat_parts;  // EPI_LN alternative to `stats`: [n_stat_parts, M] partial (sum, sumsq) of the input rows
  int n_stat_parts;
  float ln_inv_d, ln_eps;
  float2* stats_out;
  const bf16* residual;
  const float* pos;
  const int32_t* tok_idx;
  int pos_rows;
  bf16* out;
  PatchGeom pg;
};

// Row bookkeeping of one M-tile: which output rows the 128 accumulator lanes map to.
struct TileRows {
  int row0;        // output row of lane 0
  int valid_rows;  // lanes >= valid_rows are padding
};

template <bool kPatch>
STAD_DEVICE TileRows tile_rows(const KArgs& p, int m_tile) {
  TileRows t;
  if constexpr (kPatch) {
    const PatchGeom& g = p.pg;
    const int hh = m_tile % g.h_tiles;
    const int tp = (m_tile / g.h_tiles) % g.Tp;
    const int b = m_tile / (g.h_tiles * g.Tp);
    const int h0 = hh * g.hp_tile;
    const int hrows = min(g.hp_tile, g.Hp - h0);
    t.row0 = (b * g.Tp + tp) * g.Hp * g.Wp + h0 * g.Wp;
    t.valid_rows = hrows * g.Wp;
  } else {
    t.row0 = m_tile * BM;
    t.valid_rows = min(BM, p.M - t.row0);
  }
  return t;
}

template <int BN, int EPI, bool kPatch, bool kPair = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res, const KArgs p) {
  using C = Cfg<BN, kPair>;
  static_assert(!(kPair && kPatch), "the CTA-pair tile is not combined with the patch-embed A operand");
  // Pair mode: work item = (pair of consecutive M-tiles, n_tile); CTA rank r of the pair owns M-tile 2 * m_pair + r.
  const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int worker = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int n_workers = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned tiles; align in the shared address space.
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint8_t* staging = smem + C::STAGES * C::STAGE_BYTES;  // [warpgroup][buffer][128 rows x 64 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + C::STAGING_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full_bar = empty_bar + C::STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* buf_ready_bar = tmem_empty_bar + 2;  // [warpgroup][buffer]: staging buffer free (+ residual landed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(buf_ready_bar + C::READY_BARS);

  // warp index through a shuffle: the compiler then knows every value derived from it is warp-uniform, and the
  // single-lane TMA / tcgen05.mma issue below takes its operands straight from uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // pair mode with an odd number of M-tiles: the last pair's second M-tile lies wholly beyond row M — its TMA loads
  // return zeros, its stores are clipped and every per-row access is guarded by row_ok
  const int num_tiles = (kPair ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles;
  const int num_kb = p.K / BK;
  auto m_tile_of = [&](int tile) { return kPair ? 2 * (tile / p.n_tiles) + static_cast<int>(cta_rank) : tile / p.n_tiles; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
    if constexpr (EPI & EPI_RESID) tma_prefetch_desc(&tmap_res);
    for (int s = 0; s < C::READY_BARS; ++s) mbar_init(&buf_ready_bar[s], 1);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      // pair: the leader's barrier collects the epilogue warps of BOTH CTAs before an accumulator stage is reused
      mbar_init(&tmem_empty_bar[s], kPair ? 2 * kEpiWarps : kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (kPair) {
      tmem_alloc_pair<C::TMEM_COLS>(tmem_slot);
      tmem_relinquish_pair();
    } else {
      tmem_alloc<C::TMEM_COLS>(tmem_slot);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();  // the peer's barriers exist before anything arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above touched only this kernel's own shared memory / TMEM: under programmatic dependent launch it ran
  // while the previous kernel was still draining.  From here on global memory is read and written.
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // The whole warp runs the (warp-uniform) loop and the waits; one elected lane issues the copies.
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = worker; tile < num_tiles; tile += n_workers) {
      const int m_tile = m_tile_of(tile);
      const int n_tile = tile % p.n_tiles;
      int a_bytes = C::A_BYTES;
      int pe_b = 0, pe_tp = 0, pe_h0 = 0, pe_f0 = 0;
      if constexpr (kPatch) {
        const PatchGeom& g = p.pg;
        const int hh = m_tile % g.h_tiles;
        pe_tp = (m_tile / g.h_tiles) % g.Tp;
        pe_b = m_tile / (g.h_tiles * g.Tp);
        pe_h0 = hh * g.hp_tile;
        // first frame of clip pe_b in the frame buffer: an arithmetic progression, or the caller's list (ABI v6)
        pe_f0 = g.win_start != nullptr ? __ldg(g.win_start + pe_b) : g.start + pe_b * g.stride;
        a_bytes = g.Wp * g.hp_tile * BK * 2;  // 4 full boxes, out-of-range h' rows arrive as zeros
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          if constexpr (kPair) {
            // both CTAs' copies are credited to the LEADER's barrier: it expects the bytes of the whole 256-row stage
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (C::A_BYTES + C::B_BYTES));
            tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], kb * BK, m_tile * BM);
            tma_load_2d_pair(sb, &tmap_b, &full_bar[stage], kb * BK, n_tile * BN + static_cast<int>(cta_rank) * C::B_ROWS);
          } else {
          mbar_arrive_expect_tx(&full_bar[stage], a_bytes + C::B_BYTES);
          if constexpr (kPatch) {
            // k chunk kb = 64 consecutive k = (c, dt, dh0..dh0+3, dw 0..15)
            const PatchGeom& g = p.pg;
            const int c = kb / (g.tubelet * 4);
            const int dt = (kb >> 2) % g.tubelet;
            const int dh0 = (kb & 3) * 4;
            const int t = pe_tp * g.tubelet + dt;
            const int plane = g.mode == STAD_IN_CLIPS ? (pe_b * g.C + c) * g.T + t : (pe_f0 + t * g.fstep) * g.C + c;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)  // K-slice k = line dh0 + k: [tokens][16 dw] at sa + k * 4 KB
              tma_load_5d(sa + k * (BM * UMMA_K * 2), &tmap_a, &full_bar[stage], 0, 0, pe_h0, dh0 + k, plane);
          } else {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m_tile * BM);
          }
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n_tile * BN);
          }
        }
        __syncwarp();
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && (!kPair || leader)) {
    // ------------------------------------------------------------------ MMA issuer (pair mode: the leader CTA only)
    // Whole warp in the loop (waits, bookkeeping); one elected lane issues tcgen05.mma / tcgen05.commit.
    constexpr uint32_t idesc = make_idesc_bf16(kPair ? 2 * BM : BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int tile = worker; tile < num_tiles; tile += n_workers, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * C::ACC_STRIDE;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          const uint64_t db = make_smem_desc_sw128(sb, 16, 1024);
          if constexpr (kPatch) {
            // A: four K-slices of [128 tokens][32 B] rows (SWIZZLE_32B, 8-row groups 256 B apart), 4 KB each
            const uint64_t da = make_smem_desc(sa, 16, 256, kLayoutSw32);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_ss(tmem_d, da + k * (BM * UMMA_K * 2 >> 4), db + 2 * k, idesc, (kb | k) != 0);
          } else {
            const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 16 elements (32 bytes) along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
              if constexpr (kPair) umma_ss_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
              else umma_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            }
          }
          if constexpr (kPair) {
            umma_commit_pair(&empty_bar[stage]);  // frees the stage in BOTH CTAs
            if (kb == num_kb - 1) umma_commit_pair(&tmem_full_bar[acc]);  // both CTAs' epilogues
          } else {
          umma_commit(&empty_bar[stage]);  // smem slot is free once these MMAs have read it
          if (kb == num_kb - 1) umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
          }
        }
        __syncwarp();
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;       // 0..7
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int wg = ew >> 2;        // warpgroup = which half of the BN columns
    constexpr int COLS = BN / 2;
    constexpr int CHUNKS = COLS / kSubCols;
    const bool lead_warp = (ew & 3) == 0;            // its elected lane issues the TMA traffic of this warpgroup
    const int r = quarter * 32 + lane;               // accumulator row (TMEM lane) of this thread
    uint8_t* stage_buf = kPair ? staging + ew * kWarpSlots * kWarpSlotBytes : staging + wg * 2 * kSubTileBytes;
    uint64_t* ready = kPair ? buf_ready_bar + ew * kWarpSlots : buf_ready_bar + wg * 2;
    // 64-byte swizzle of the staging tile: 16-byte chunk c of row r lives at chunk position c ^ ((r >> 1) & 3)
    // (pair tile: rows are counted within the warp's own 32-row slot)
    const uint32_t srow = kPair ? static_cast<uint32_t>(lane) : static_cast<uint32_t>(r);
    const uint32_t row_off = srow * 64u;
    const uint32_t swz = (srow >> 1) & 3u;

    // Sub-tile coordinates of running chunk index `ci` of this CTA/warpgroup (tile-major, then column chunk).
    auto chunk_coords = [&](int ci, int& col, int& row0) -> bool {
      const int tile = worker + (ci / CHUNKS) * n_workers;
      if (tile >= num_tiles) return false;
      const int m_tile = m_tile_of(tile);
      const int n_tile = tile % p.n_tiles;
      col = n_tile * BN + wg * COLS + (ci % CHUNKS) * kSubCols;
      row0 = tile_rows<kPatch>(p, m_tile).row0 + (kPair ? quarter * 32 : 0);
      return true;
    };
    // Pair tile: residual sub-tile (this warp's 32 rows x 32 columns) of running chunk `ci` -> slot ci % kWarpSlots.
    // The caller has made sure the TMA store that last used the slot has finished reading it.
    auto prefetch_residual = [&](int ci) {
      if constexpr (kPair && (EPI & EPI_RESID)) {
        int col, row0;
        if (!chunk_coords(ci, col, row0)) return;
        uint64_t* bar = &ready[ci & (kWarpSlots - 1)];
        mbar_arrive_expect_tx(bar, kWarpSlotBytes);
        tma_load_2d(stage_buf + (ci & (kWarpSlots - 1)) * kWarpSlotBytes, &tmap_res, bar, col, row0);
      }
    };
    // Leader: make staging buffer (ci & 1) usable for chunk ci: the TMA store issued from it two chunks ago must have
    // finished READING it; then either start fetching the residual sub-tile into it or just mark it free.
    auto prepare_buffer = [&](int ci) {
      int col, row0;
      if (!chunk_coords(ci, col, row0)) return;
      tma_store_wait_read<1>();
      uint64_t* bar = &ready[ci & 1];
      if constexpr (EPI & EPI_RESID) {
        mbar_arrive_expect_tx(bar, kSubTileBytes);
        tma_load_2d(stage_buf + (ci & 1) * kSubTileBytes, &tmap_res, bar, col, row0);
      } else {
        mbar_arrive(bar);
      }
    };

    int ci = 0;
    if constexpr (kPair) {
      if (elect_one()) {  // residual sub-tiles are requested two chunks ahead of their use
        prefetch_residual(0);
        prefetch_residual(1);
      }
      __syncwarp();
    } else if (lead_warp) {
      if (elect_one()) prepare_buffer(0);
      __syncwarp();
    }
    // EPI_LN with stat_parts: the partial (sum, sumsq) of this thread's row are fetched one tile AHEAD (registers), so
    // their L2 latency is hidden behind the previous tile's epilogue instead of sitting at the head of every tile's
    // (the dynamic loop of dependent loads it replaces made folding a loss beyond ~8 k rows; now every LayerNorm-folded
    // GEMM finishes its own statistics and the stats_finalize launches are gone from the model paths).
    float2 pre_parts[kMaxFoldParts];
    auto fetch_parts = [&](int tile) {
      if constexpr (EPI & EPI_LN) {
        if (p.stat_parts != nullptr && p.n_stat_parts <= kMaxFoldParts && tile < num_tiles) {
          const TileRows t = tile_rows<kPatch>(p, m_tile_of(tile));
          const bool ok = r < t.valid_rows;
          const float2* src = p.stat_parts + (t.row0 + r);
#pragma unroll
          for (int i = 0; i < kMaxFoldParts; ++i)
            pre_parts[i] = (ok && i < p.n_stat_parts) ? __ldg(src + static_cast<size_t>(i) * p.M) : make_float2(0.f, 0.f);
        }
      }
    };
    fetch_parts(worker);
    int local = 0;
    for (int tile = worker; tile < num_tiles; tile += n_workers, ++local) {
      const int m_tile = m_tile_of(tile);
      const int n_tile = tile % p.n_tiles;
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const TileRows tr = tile_rows<kPatch>(p, m_tile);
      const bool row_ok = r < tr.valid_rows;
      const int row = tr.row0 + r;

      float mean = 0.f, rstd = 1.f;
      if constexpr (EPI & EPI_LN) {
        if (row_ok) {
          if (p.stat_parts != nullptr) {
            // LayerNorm statistics straight from the partial sums the producing GEMM emitted: the same index-order
            // sum and the same arithmetic as stats_finalize_kernel (bit-identical; absent parts add 0), without that
            // launch
            float s1 = 0.f, s2 = 0.f;
            if (p.n_stat_parts <= kMaxFoldParts) {
#pragma unroll
              for (int i = 0; i < kMaxFoldParts; ++i) {
                s1 += pre_parts[i].x;
                s2 += pre_parts[i].y;
              }
            } else {  // many narrow column tiles (small batches): few tiles per CTA, nothing to hide the loads behind
              for (int i = 0; i < p.n_stat_parts; ++i) {
                const float2 v = __ldg(&p.stat_parts[static_cast<size_t>(i) * p.M + row]);
                s1 += v.x;
                s2 += v.y;
              }
            }
            mean = s1 * p.ln_inv_d;
            const float var = fmaxf(fmaf(-mean, mean, s2 * p.ln_inv_d), 0.f);
            rstd = rsqrtf(var + p.ln_eps);
          } else {
            const float2 st = __ldg(&p.stats[row]);
            mean = st.x;
            rstd = st.y;
          }
        }
        fetch_parts(tile + n_workers);
      }
      float st_sum = 0.f, st_sq = 0.f;  // EPI_STATS: running (sum, sumsq) of this thread's part of the row
      const float* pos_row = nullptr;
      if constexpr (EPI & EPI_POS) {
        if (row_ok && p.pos != nullptr) {  // no table: the accumulator is stored as it is (tubelet embeddings, api.cu)
          const int pr = p.tok_idx ? __ldg(&p.tok_idx[row]) : row % p.pos_rows;
          pos_row = p.pos + static_cast<size_t>(pr) * p.N;
        }
      }

      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * C::ACC_STRIDE + wg * COLS;
      uint32_t v[32];
      tmem_ld32(taddr, v);
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c, ++ci) {
        tmem_ld_wait32(v);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (c + 1 < CHUNKS) {
          tmem_ld32(taddr + (c + 1) * kSubCols, v);  // next chunk streams in under this chunk's math
        } else {
          // accumulator stage fully in registers: hand it back to the MMA issuer (pair mode: in the leader CTA)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (kPair) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
            else mbar_arrive(&tmem_empty_bar[acc]);
          }
        }
        const int n0 = n_tile * BN + wg * COLS + c * kSubCols;
        if constexpr (EPI & EPI_LN) {
          // rstd * (acc - mean * colsum) + bias  as two packed FMAs per column pair
          const float nmean = -mean;
          const bool has_bias = p.bias != nullptr;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 cs = __ldg(reinterpret_cast<const float4*>(p.colsum + n0 + j));
            const float4 bv = has_bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
            fma2(f[j + 0], f[j + 1], cs.x, cs.y, nmean, nmean, f[j + 0], f[j + 1]);
            fma2(f[j + 2], f[j + 3], cs.z, cs.w, nmean, nmean, f[j + 2], f[j + 3]);
            fma2(f[j + 0], f[j + 1], f[j + 0], f[j + 1], rstd, rstd, bv.x, bv.y);
            fma2(f[j + 2], f[j + 3], f[j + 2], f[j + 3], rstd, rstd, bv.z, bv.w);
          }
        } else if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
            add2(f[j + 0], f[j + 1], f[j + 0], f[j + 1], bv.x, bv.y);
            add2(f[j + 2], f[j + 3], f[j + 2], f[j + 3], bv.z, bv.w);
          }
        }
        if constexpr (EPI & EPI_GELU) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) gelu_erf2(f[j], f[j + 1], f[j], f[j + 1]);
        }
        if constexpr (EPI & EPI_POS) {
          if (pos_row != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 pv = __ldg(reinterpret_cast<const float4*>(pos_row + n0 + j));
              f[j + 0] += pv.x;
              f[j + 1] += pv.y;
              f[j + 2] += pv.z;
              f[j + 3] += pv.w;
            }
          }
        }

        // staging buffer of this chunk is free (and holds the residual sub-tile, if any)
        uint8_t* buf;
        if constexpr (kPair) {
          // this warp's ring of kWarpSlots (32 rows x 64 B) slots: chunk ci uses slot ci % kWarpSlots
          const int slot = ci & (kWarpSlots - 1);
          buf = stage_buf + slot * kWarpSlotBytes;
          if constexpr (EPI & EPI_RESID) {
            mbar_wait(&ready[slot], (ci / kWarpSlots) & 1);  // residual landed; its prefetch knew the slot to be free
          } else {
            // stores leave in batches of two chunks (see below): slots {0,1} / {2,3} are free once the batch that last
            // used them, two batches ago, has been read
            if ((c & 1) == 0) {
              if (elect_one()) tma_store_wait_read<1>();
              __syncwarp();
            }
          }
        } else {
          buf = stage_buf + (ci & 1) * kSubTileBytes;
          mbar_wait(&ready[ci & 1], (ci >> 1) & 1);
        }
        uint4* my_row = reinterpret_cast<uint4*>(buf + row_off);
        if constexpr (EPI & EPI_RESID) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 rv = my_row[q ^ swz];
            f[q * 8 + 0] += bf16_lo(rv.x);
            f[q * 8 + 1] += bf16_hi(rv.x);
            f[q * 8 + 2] += bf16_lo(rv.y);
            f[q * 8 + 3] += bf16_hi(rv.y);
            f[q * 8 + 4] += bf16_lo(rv.z);
            f[q * 8 + 5] += bf16_hi(rv.z);
            f[q * 8 + 6] += bf16_lo(rv.w);
            f[q * 8 + 7] += bf16_hi(rv.w);
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack_bf16(f[q * 8 + 0], f[q * 8 + 1]);
          o.y = pack_bf16(f[q * 8 + 2], f[q * 8 + 3]);
          o.z = pack_bf16(f[q * 8 + 4], f[q * 8 + 5]);
          o.w = pack_bf16(f[q * 8 + 6], f[q * 8 + 7]);
          my_row[q ^ swz] = o;
          if constexpr (EPI & EPI_STATS) {
            // statistics of the values as stored (bf16), which is what the next LayerNorm sees
            const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float lo = bf16_lo(w[t]), hi = bf16_hi(w[t]);
              st_sum += lo + hi;
              st_sq = fmaf(lo, lo, fmaf(hi, hi, st_sq));
            }
          }
        }
        if constexpr (kPair) {
          // Every warp stores its own 32 x 32 sub-tile.  The proxy fence is executed by the ISSUING lane only, after
          // the warp barrier that makes the other lanes' shared-memory writes visible to it: executed by all 32 lanes
          // (the single-CTA path does that before its 128-thread barrier) it costs 35 us of a 242 us qkv GEMM, the
          // lanes stalling on it while bulk stores are in flight (ordering: writes -> __syncwarp -> fence.proxy.async
          // -> cp.async.bulk of the same thread; tests/kernel_checks.py::check_gemm_pair fails reliably without it).
          __syncwarp();
          if constexpr (EPI & EPI_RESID) {
            if (elect_one()) {
              fence_proxy_async_smem();
              tma_store_2d(&tmap_out, buf, n0, tr.row0 + quarter * 32);
              tma_store_commit();
              tma_store_wait_read<2>();   // slot of chunk ci + 2 was last stored from by chunk ci - 2
              prefetch_residual(ci + 2);
            }
          } else if ((c & 1) || c == CHUNKS - 1) {
            // Without a residual the stores leave two chunks at a time: fence.proxy.async stalls the issuing lane while
            // bulk stores of the CTA are still in flight, and half a tile after the previous batch they no longer are.
            // (Three chunks per warpgroup, BN = 192: a batch of two, then the last chunk alone.)
            if (elect_one()) {
              fence_proxy_async_smem();
              if (c & 1)
                tma_store_2d(&tmap_out, stage_buf + ((ci - 1) & (kWarpSlots - 1)) * kWarpSlotBytes, n0 - kSubCols,
                             tr.row0 + quarter * 32);
              tma_store_2d(&tmap_out, buf, n0, tr.row0 + quarter * 32);
              tma_store_commit();
            }
          }
          __syncwarp();
          continue;
        }
        fence_proxy_async_smem();         // generic-proxy writes -> visible to the TMA (async proxy)
        named_bar_sync(1 + wg, 128);      // whole sub-tile staged
        if (lead_warp) {
          if (elect_one()) {  // deterministic: always the same lane, which owns the bulk async-groups
            tma_store_2d(&tmap_out, buf, n0, tr.row0);  // rows beyond M (or beyond the box) are clipped by the TMA
            tma_store_commit();
            prepare_buffer(ci + 1);
          }
          __syncwarp();
        }
      }
      if constexpr (EPI & EPI_STATS) {
        // one partial per (column tile, warpgroup), partial-major so that a warp's 32 rows are contiguous
        if (row_ok)
          p.stats_out[static_cast<size_t>(n_tile * 2 + wg) * p.M + row] = make_float2(st_sum, st_sq);
      }
    }
    if (kPair || lead_warp) {
      if (elect_one()) tma_store_wait<0>();  // all output tiles of this CTA are in global memory
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();  // neither CTA frees TMEM / exits while the peer's MMAs or arrivals are in flight
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if constexpr (kPair) tmem_dealloc_pair<C::TMEM_COLS>(tmem_base);
    else tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

template <int BN, int EPI, bool kPatch>
int launch_one(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr,
               const KArgs& ka, cudaStream_t stream) {
  using C = Cfg<BN>;
  const int tiles = ka.m_tiles * ka.n_tiles;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  ProfScope prof(STAD_K_GEMM, EPI | (kPatch ? 32 : 0), ka.M, ka.N, ka.K, stream);
  STAD_CUDA_OK(launch_pdl(gemm_kernel<BN, EPI, kPatch, false>, dim3(grid), dim3(kThreads), C::SMEM_BYTES, stream, 1, ta, tb,
                          to, tr, ka));
  STAD_LAUNCH_OK("gemm_kernel");
  return STAD_OK;
}

template <int BN, int EPI, bool kPatch>
int set_smem() {
  STAD_CUDA_OK(cudaFuncSetAttribute(gemm_kernel<BN, EPI, kPatch>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg<BN>::SMEM_BYTES));
  return STAD_OK;
}

// CTA-pair variant (256 x 256 tiles, cta_group::2): launched as clusters of two CTAs, one pair per TPC.
template <int BN, int EPI>
int launch_pair_bn(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr,
                   const KArgs& ka, cudaStream_t stream) {
  using C = Cfg<BN, true>;
  const int tiles = ((ka.m_tiles + 1) / 2) * ka.n_tiles;
  const int pairs = tiles < sm_count() / 2 ? tiles : sm_count() / 2;
  ProfScope prof(STAD_K_GEMM, EPI | 64, ka.M, ka.N, ka.K, stream);
  STAD_CUDA_OK(launch_pdl(gemm_kernel<BN, EPI, false, true>, dim3(2 * pairs), dim3(kThreads), C::SMEM_BYTES, stream, 2, ta,
                          tb, to, tr, ka));
  STAD_LAUNCH_OK("gemm_kernel (CTA pair)");
  return STAD_OK;
}

// 256 x 256 tiles; 256 x 192 where 256 does not divide N (N = 384, 1152: ViT-S, the MAE decoder) only on request
template <int EPI>
int launch_pair(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr,
                const KArgs& ka, cudaStream_t stream) {
  if (bn == 192) return launch_pair_bn<192, EPI>(ta, tb, to, tr, ka, stream);
  return launch_pair_bn<256, EPI>(ta, tb, to, tr, ka, stream);
}

template <int EPI>
int set_smem_pair() {
  STAD_CUDA_OK(cudaFuncSetAttribute(gemm_kernel<256, EPI, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg<256, true>::SMEM_BYTES));
  STAD_CUDA_OK(cudaFuncSetAttribute(gemm_kernel<192, EPI, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg<192, true>::SMEM_BYTES));
  return STAD_OK;
}

// Which shapes run CTA-pair (256-row) tiles — settled by A/B measurements on the B200, recorded under profiles/:
//   * 256 x 256 pair tiles for every GEMM with N % 256 == 0 (B = 64 ViT-B, single-CTA -> pair: qkv 259 -> 232 us, fc1
//     387 -> 346, fc2 337 -> 318, proj 118 -> 118);
//   * an odd number of M-tiles is padded with a virtual, fully clipped M-tile rather than falling back to single-CTA
//     tiles (masked encoder 32.4 k -> 34.2 k clips/s, MVD + class token 2417 -> 2568; profiles/r1c_pair_odd_m_tiles_ab.txt);
//   * 256 x 192 pair tiles (N = 384 / 1152: ViT-S, the MAE decoder) are NOT used: measured slower than the single-CTA
//     192-column tile on every shape (ViT-S 6407 -> 6260 clips/s, full MAE forward 11.48 k -> 11.09 k;
//     profiles/r1c_pair192_ab.txt) — with K = 384 the mainloop is six k-blocks long and the tile is bound by its
//     epilogue, which the pair does not shorten.  The kernel stays generic in BN (tests/kernel_checks.py keeps the
//     192-wide single-CTA tiles covered).
constexpr bool kPairTiles192 = false;

template <int EPI, bool kPatch>
int set_smem_all_bn() {
  int rc;
  if ((rc = set_smem<64, EPI, kPatch>())) return rc;
  if ((rc = set_smem<128, EPI, kPatch>())) return rc;
  if ((rc = set_smem<192, EPI, kPatch>())) return rc;
  if ((rc = set_smem<256, EPI, kPatch>())) return rc;
  return STAD_OK;
}

template <int EPI, bool kPatch>
int dispatch_bn(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr,
                const KArgs& ka, cudaStream_t stream) {
  switch (bn) {
    case 64: return launch_one<64, EPI, kPatch>(ta, tb, to, tr, ka, stream);
    case 128: return launch_one<128, EPI, kPatch>(ta, tb, to, tr, ka, stream);
    case 192: return launch_one<192, EPI, kPatch>(ta, tb, to, tr, ka, stream);
    case 256: return launch_one<256, EPI, kPatch>(ta, tb, to, tr, ka, stream);
  }
  return fail(STAD_E_SHAPE, "gemm: no tile width for N");
}

// Column-tile width: the divisor of N that minimises  rounds x (BN + 64),  rounds = ceil(tiles / SMs) being the number of
// tile rounds of the persistent grid and 64 columns standing for the fixed cost of a tile (pipeline fill, accumulator
// hand-over, epilogue tail).  With many tiles per SM this is the widest tile; at small batch it avoids a mostly empty
// last round (batch 1, fc1: 156 tiles of 256 columns = 2 rounds -> 208 tiles of 192) and too-narrow tiles (proj: 156
// tiles of 64 = 2 rounds -> 78 tiles of 128 = 1 round).  Ties go to the wider tile.
int pick_bn(int m_tiles, int N) {
  const int cand[4] = {256, 192, 128, 64};
  int best = 0;
  long best_cost = 0;
  for (int i = 0; i < 4; ++i) {
    if (N % cand[i] != 0) continue;
    const long tiles = static_cast<long>(m_tiles) * (N / cand[i]);
    const long rounds = (tiles + sm_count() - 1) / sm_count();
    const long cost = rounds * (cand[i] + 64);
    if (best == 0 || cost < best_cost) {
      best = cand[i];
      best_cost = cost;
    }
  }
  return best;
}

}  // namespace

int gemm_init() {
  int rc;
  if ((rc = set_smem_all_bn<0, false>())) return rc;
  if ((rc = set_smem_all_bn<EPI_LN, false>())) return rc;
  if ((rc = set_smem_all_bn<EPI_LN | EPI_GELU, false>())) return rc;
  if ((rc = set_smem_all_bn<EPI_RESID, false>())) return rc;
  if ((rc = set_smem_all_bn<EPI_POS, false>())) return rc;
  if ((rc = set_smem_all_bn<EPI_POS, true>())) return rc;
  if ((rc = set_smem_all_bn<EPI_RESID | EPI_STATS, false>())) return rc;
  if ((rc = set_smem_all_bn<EPI_POS | EPI_STATS, false>())) return rc;
  if ((rc = set_smem_all_bn<EPI_POS | EPI_STATS, true>())) return rc;
  if ((rc = set_smem_all_bn<EPI_LN | EPI_POS, false>())) return rc;
  if ((rc = set_smem_pair<0>())) return rc;
  if ((rc = set_smem_pair<EPI_LN>())) return rc;
  if ((rc = set_smem_pair<EPI_LN | EPI_GELU>())) return rc;
  if ((rc = set_smem_pair<EPI_RESID>())) return rc;
  if ((rc = set_smem_pair<EPI_RESID | EPI_STATS>())) return rc;
  return STAD_OK;
}

int gemm_stat_parts(int M, int N, bool patch, const PatchGeom* pg) {
  const int m_tiles = patch ? (M / (pg->Tp * pg->Hp * pg->Wp)) * pg->Tp * pg->h_tiles : ceil_div(M, BM);
  return 2 * (N / pick_bn(m_tiles, N));
}

int launch_gemm(const GemmArgs& g, cudaStream_t stream) {
  STAD_CHECK_ARG(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  STAD_CHECK_ARG(g.K % BK == 0, "gemm: K=%d must be a multiple of %d", g.K, BK);
  STAD_CHECK_ARG(g.N % 64 == 0, "gemm: N=%d must be a multiple of 64", g.N);
  if ((reinterpret_cast<uintptr_t>(g.a) | reinterpret_cast<uintptr_t>(g.w) | reinterpret_cast<uintptr_t>(g.out) |
       reinterpret_cast<uintptr_t>(g.residual) | reinterpret_cast<uintptr_t>(g.bias) |
       reinterpret_cast<uintptr_t>(g.colsum) | reinterpret_cast<uintptr_t>(g.pos)) & 15)
    return fail(STAD_E_ALIGN, "gemm: all operands must be 16-byte aligned");
  if (g.epi & EPI_LN)
    STAD_CHECK_ARG((g.stats || (g.stat_parts && g.n_stat_parts >= 1 && g.n_stat_parts <= kMaxStatParts)) && g.colsum,
                   "gemm: LN epilogue needs stats (or partial sums) and colsum");
  if (g.epi & EPI_RESID) STAD_CHECK_ARG(g.residual, "gemm: residual epilogue needs a residual");
  if (g.epi & EPI_POS) STAD_CHECK_ARG(g.pos == nullptr || g.tok_idx || g.pos_rows > 0, "gemm: pos epilogue needs pos_rows");
  if (g.epi & EPI_STATS) STAD_CHECK_ARG(g.stats_out, "gemm: statistics epilogue needs an output buffer");
  if ((reinterpret_cast<uintptr_t>(g.stats) | reinterpret_cast<uintptr_t>(g.stats_out)) & 7)
    return fail(STAD_E_ALIGN, "gemm: statistics buffers must be 8-byte aligned");

  KArgs ka{};
  ka.M = g.M;
  ka.N = g.N;
  ka.K = g.K;
  ka.bias = g.bias;
  ka.colsum = g.colsum;
  ka.stats = g.stats;
  ka.stat_parts = g.stats ? nullptr : g.stat_parts;
  ka.n_stat_parts = g.n_stat_parts;
  ka.ln_inv_d = 1.0f / static_cast<float>(g.K);
  ka.ln_eps = g.ln_eps;
  ka.stats_out = g.stats_out;
  ka.residual = g.residual;
  ka.pos = g.pos;
  ka.tok_idx = g.tok_idx;
  ka.pos_rows = g.pos_rows;
  ka.out = g.out;

  CUtensorMap ta, tb, to, tr;
  int rc;
  int out_box_rows = BM;
  if (g.patch) {
    const PatchGeom& pg = *g.patch;
    ka.pg = pg;
    ka.m_tiles = (g.M / (pg.Tp * pg.Hp * pg.Wp)) * pg.Tp * pg.h_tiles;
    // dims innermost-first: dw(16) w'(Wp) h'(Hp) dh(16) plane; box = one (dh, plane) line of every token of the tile
    const uint64_t dims[5] = {16, (uint64_t)pg.Wp, (uint64_t)pg.Hp, 16, (uint64_t)pg.n_planes};
    const uint64_t strides[4] = {32, (uint64_t)pg.img_w * 32, (uint64_t)pg.img_w * 2,
                                 (uint64_t)pg.img_h * pg.img_w * 2};
    const uint32_t box[5] = {16, (uint32_t)pg.Wp, (uint32_t)pg.hp_tile, 1, 1};
    if ((rc = make_tmap_bf16(&ta, g.a, 5, dims, strides, box, 32))) return rc;
    // every patch tile holds exactly hp_tile * Wp tokens (make_geom picks hp_tile | Hp): the store box is that tall
    STAD_CHECK_ARG(pg.Hp % pg.hp_tile == 0, "patch mode: hp_tile=%d must divide Hp=%d", pg.hp_tile, pg.Hp);
    out_box_rows = pg.hp_tile * pg.Wp;
  } else {
    ka.m_tiles = ceil_div(g.M, BM);
    const uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)g.M};
    const uint64_t strides[1] = {(uint64_t)g.K * 2};
    const uint32_t box[2] = {BK, BM};
    if ((rc = make_tmap_bf16(&ta, g.a, 2, dims, strides, box))) return rc;
  }
  const int bn = pick_bn(ka.m_tiles, g.N);
  ka.n_tiles = g.N / bn;
  // CTA-pair tiles (256 x 256) when the shape allows: full-width column tiles, at least one tile per pair, and an
  // epilogue that has a pair instantiation (an odd M-tile count is padded with a virtual, fully clipped M-tile)
  const bool pair_epi = g.epi == 0 || g.epi == EPI_LN || g.epi == (EPI_LN | EPI_GELU) || g.epi == EPI_RESID ||
                        g.epi == (EPI_RESID | EPI_STATS);
  // (measured, B = 64 ViT-B, single-CTA -> pair tile with the warp-private epilogue: qkv 259 -> 232 us, fc1 387 -> 346,
  // fc2 337 -> 318, proj 118 -> 118)
  const bool pair = !g.patch && (bn == 256 || (bn == 192 && kPairTiles192)) && pair_epi &&
                    ((ka.m_tiles + 1) / 2) * ka.n_tiles >= sm_count() / 2;
  {
    const uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)g.N};
    const uint64_t strides[1] = {(uint64_t)g.K * 2};
    const uint32_t box[2] = {BK, (uint32_t)(pair ? bn / 2 : bn)};
    if ((rc = make_tmap_bf16(&tb, g.w, 2, dims, strides, box))) return rc;
  }

  {
    // output (and residual) sub-tiles: 32 columns x out_box_rows rows, 64-byte swizzle
    const uint64_t dims[2] = {(uint64_t)g.N, (uint64_t)g.M};
    const uint64_t strides[1] = {(uint64_t)g.N * 2};
    const uint32_t box[2] = {kSubCols, (uint32_t)(pair ? 32 : out_box_rows)};  // pair tile: one warp's rows per store
    if ((rc = make_tmap_bf16(&to, g.out, 2, dims, strides, box, 64))) return rc;
    tr = to;
    if (g.epi & EPI_RESID)
      if ((rc = make_tmap_bf16(&tr, g.residual, 2, dims, strides, box, 64))) return rc;
  }

  if (g.patch) {
    STAD_CHECK_ARG((g.epi & ~EPI_STATS) == EPI_POS, "gemm: patch mode supports only the pos epilogue");
    if (g.epi & EPI_STATS) return dispatch_bn<EPI_POS | EPI_STATS, true>(bn, ta, tb, to, tr, ka, stream);
    return dispatch_bn<EPI_POS, true>(bn, ta, tb, to, tr, ka, stream);
  }
  if (pair) {
    switch (g.epi) {
      case 0: return launch_pair<0>(bn, ta, tb, to, tr, ka, stream);
      case EPI_LN: return launch_pair<EPI_LN>(bn, ta, tb, to, tr, ka, stream);
      case EPI_LN | EPI_GELU: return launch_pair<EPI_LN | EPI_GELU>(bn, ta, tb, to, tr, ka, stream);
      case EPI_RESID: return launch_pair<EPI_RESID>(bn, ta, tb, to, tr, ka, stream);
      case EPI_RESID | EPI_STATS: return launch_pair<EPI_RESID | EPI_STATS>(bn, ta, tb, to, tr, ka, stream);
    }
  }
  switch (g.epi) {
    case EPI_RESID | EPI_STATS: return dispatch_bn<EPI_RESID | EPI_STATS, false>(bn, ta, tb, to, tr, ka, stream);
    case EPI_POS | EPI_STATS: return dispatch_bn<EPI_POS | EPI_STATS, false>(bn, ta, tb, to, tr, ka, stream);
    case 0: return dispatch_bn<0, false>(bn, ta, tb, to, tr, ka, stream);
    case EPI_LN: return dispatch_bn<EPI_LN, false>(bn, ta, tb, to, tr, ka, stream);
    case EPI_LN | EPI_GELU: return dispatch_bn<EPI_LN | EPI_GELU, false>(bn, ta, tb, to, tr, ka, stream);
    case EPI_RESID: return dispatch_bn<EPI_RESID, false>(bn, ta, tb, to, tr, ka, stream);
    case EPI_POS: return dispatch_bn<EPI_POS, false>(bn, ta, tb, to, tr, ka, stream);
    case EPI_LN | EPI_POS: return dispatch_bn<EPI_LN | EPI_POS, false>(bn, ta, tb, to, tr, ka, stream);
  }
  return fail(STAD_E_SHAPE, "gemm: unsupported epilogue combination %d", g.epi);
}

}  // namespace stad
