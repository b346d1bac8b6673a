// Persistent "stream-K" variant of the warp-specialised forward kernel (non-causal, Nq % 256 == 0).
//
// Why: fa_fwd_ws_kernel runs one CTA per 256-row query block, so a launch with U blocks on G = 148 SMs
// takes ceil(U / G) rounds: B=1, H=16 gives 128 blocks at N=2048 (one round on 128 of 148 SMs), 256 at
// N=4096 (2 rounds for 1.73 rounds of work) and 512 at N=8192 (4 rounds for 3.46).  Here the work is
// linearised as (unit, KV tile) items, unit = (batch, head, 256-row block) with T = ceil(Nkv / 128) items
// each, and the grid is min(G, items) persistent CTAs.  The first sk_dp rounds are data-parallel (CTA c takes
// unit round * G + c: the CTAs of a round work on neighbouring heads, so K/V stays L2-resident as in the
// one-shot kernel - splitting ALL units contiguously made every CTA stream a different head and was
// HBM-bound at N=16384).  The remaining units are split evenly: CTA c takes the contiguous range
// [c W / G, (c+1) W / G) of their W items, so every SM gets the same number of KV tiles (+-1), and every
// CTA pays the prologue (TMEM allocation, barrier initialisation, descriptor prefetch) once.
//
// A range is a sequence of SEGMENTS, one per unit it touches:
//   * a segment that does not start at KV tile 0 (a tail or interior part; only ever the FIRST segment of a
//     CTA) is a PRODUCER part: unnormalised O, m and l go to the CTA's workspace slot, then a flag is raised;
//   * a segment that starts at tile 0 but stops short of T (only ever the LAST segment of a CTA) is the
//     CONSUMER part: it waits for the flags of the CTAs that own the rest of the unit (cta+1, cta+2, ... -
//     ranges shorter than a unit are allowed, so there can be several), merges
//       O = sum_i O_i 2^((m_i-M)c),  l likewise,  M = max_i m_i,
//     normalises, stores, and lowers the flags again (so launches and CUDA-graph replays start clean);
//   * anything else is a whole unit.
// No deadlock: producers never wait on another CTA; consumers wait at the very end of their range.
//
// Unit boundaries inside a CTA are pipelined (round 2): Q has its own buffers and O its own staging tile,
// so the next unit's Q is fetched as soon as the last S product of the current unit has read the old one
// (bar_q_free, a tcgen05.commit), its first S is issued right behind the current unit's last PV, and the only
// thing the tensor cores wait for at a boundary is the read-out of O from tensor memory (bar_tile_free).
// Partials are laid out so that every warp-level access of the exchange is one contiguous 512-byte run
// ([tile][half][float4 index][row]): the row-major layout of round 1 cost 8x the LSU wavefronts.
//
// Everything per KV tile - roles, barriers, TMEM layout, the softmax step - is that of
// fa_fwd_ws.cuh; barrier parities run on counters that continue across the units of a CTA.
#pragma once
#include "fa_fwd_ws.cuh"
#include "fa_fwd_ws3.cuh"  // ws3_softmax_step: the early-S protocol at head dim 64

// Head-dim-128 path of the persistent kernel: when Nkv is a multiple of the KV tile there is no ragged tile anywhere, and
// every step after a segment's first runs the mask-free instantiation in ONE loop (+0.35 % at N=8192 and on config 5's shards;
// the first | interior | last split of the other kernels costs this path 2 %, see the loop).
#ifndef FA_SK128_NOTAIL_LOOP
#define FA_SK128_NOTAIL_LOOP 1
#endif

namespace fa {

// Debug-only timeline (-DFA_TRACE): the leader thread of tile 0 stamps globaltimer into p.trace[cta * 32 + slot]:
// slot 0 = start, then per segment 1 + 3 seg = KV loop done, 2 + 3 seg = partials awaited, 3 + 3 seg = epilogue done
// (tools/trace_sk.py)
#ifdef FA_TRACE
#define FA_SK_TR(slot)                                                                                   \
  do {                                                                                                   \
    if (p.trace != nullptr && tile_leader && t == 0 && (slot) < 32) p.trace[cta * 32 + (slot)] = globaltimer_ns(); \
  } while (0)
#else
#define FA_SK_TR(slot) do { } while (0)
#endif

// workspace slot of one CTA: unnormalised O of both tiles (fp32, [tile][half][kDP/8 float4s][128 rows]),
// then (m, l) per row
template <int kDP>
struct SkSlot {
  static constexpr int kOFloats = 2 * kTileM * kDP;
  static constexpr int kFloats = kOFloats + 2 * kTileM * 2;
};
constexpr int kSkSlotFloatsMax = SkSlot<128>::kFloats;

// Shared memory: 2 Q tiles, the K/V ring, ONE O staging tile (shared by the two Q tiles, which finish half a
// step apart), barriers and a single-buffered row-max exchange.  No alignment slack: the dynamic
// shared-memory window starts 1024-byte aligned (checked at kernel start).
template <int kDP>
struct SkCfg {
  static constexpr int kTileBytes = kTileM * kDP * 2;
  static constexpr int kStages = (kDP == 128) ? 4 : 8;
  static constexpr int kQ = 0;
  static constexpr int kKV = kQ + 2 * kTileBytes;
  static constexpr int kStage = kKV + kStages * kTileBytes;  // O staging
  static constexpr int kBars = kStage + kTileBytes;
  static constexpr int kNumBars = 27 + 2 * kStages;
  static constexpr int kMax = kBars + 8 * kNumBars + 16;  // float [2 tile][2 half][128]; also row sums
  static constexpr int kTotal = kMax + 2 * 2 * 128 * 4;
  static_assert(kTotal <= 232448, "shared memory budget");
};

// Segment walker: the same sequence in every role.  First `dp` whole units in round-robin order, then the
// stream-K range of this CTA over the remaining units.
struct SkWalker {
  int dp, round, u, t0, rem;  // (CTA index, grid size and tiles per unit are passed in: they cost no registers)
  __device__ __forceinline__ void init(const TcParams& p, int cta, int G) {
    const int T = p.sk_T;
    dp = p.sk_dp;
    round = 0;
    const long long pos_begin = p.sk_W * cta / G;
    const long long pos_end = p.sk_W * (cta + 1) / G;
    const int u_rel = static_cast<int>(pos_begin / T);
    u = p.sk_dp * G + u_rel;
    t0 = static_cast<int>(pos_begin - static_cast<long long>(u_rel) * T);
    rem = static_cast<int>(pos_end - pos_begin);
  }
  __device__ __forceinline__ bool more() const { return dp > 0 || rem > 0; }
  __device__ __forceinline__ int unit(int cta, int G) const { return dp > 0 ? round * G + cta : u; }
  __device__ __forceinline__ int first() const { return dp > 0 ? 0 : t0; }
  __device__ __forceinline__ int count(int T) const { return dp > 0 ? T : min(T - t0, rem); }
  __device__ __forceinline__ void next(int T) {
    if (dp > 0) {
      --dp;
      ++round;
    } else {
      rem -= min(T - t0, rem);
      u += 1;
      t0 = 0;
    }
  }
};

// A zero the compiler cannot see through: values computed from it are computed AFTER this point in program
// order (volatile asm is not moved across the volatile asm of the KV loop), i.e. they are not kept live in
// registers across the loop, where the 64 scores of the softmax step need every register there is.
__device__ __forceinline__ int opaque_zero() {
  int z;
  asm volatile("mov.u32 %0, 0;" : "=r"(z));
  return z;
}

template <int kDP, bool kBF16>
__global__ void __launch_bounds__(kWsThreads, 1)
fa_fwd_sk_kernel(const __grid_constant__ CUtensorMap tmap_q,
                 const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v,
                 const __grid_constant__ CUtensorMap tmap_o, const TcParams p) {
  using C = SkCfg<kDP>;
  constexpr int kS = C::kStages;
  constexpr int kDBlocks = kDP / 64;
  constexpr int kKSteps = kDP / 16;
  constexpr int kOHalf = kDP / 2;
  constexpr int kQ4 = kOHalf / 4;  // float4s of a partial row-half
  auto col_s = [](int t) -> uint32_t { return static_cast<uint32_t>(t) * 128u; };
  auto col_o = [](int t) -> uint32_t { return 256u + static_cast<uint32_t>(t) * 128u; };
  constexpr bool kEarlyS = (kDP == 64);  // P outside the S columns, S_t(g+1) issued ahead of PV_t(g) (see the barrier map)
  auto col_p = [](int t) -> uint32_t { return 256u + static_cast<uint32_t>(t) * 128u + 64u; };  // kEarlyS only

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();  // SkCfg has no alignment slack
  const uint32_t sQ = smem_u32(smem + C::kQ);
  const uint32_t sKV = smem_u32(smem + C::kKV);
  const uint32_t sStage = smem_u32(smem + C::kStage);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::kBars + 8 * C::kNumBars);
  float* sMax = reinterpret_cast<float*>(smem + C::kMax);
  float* sFinal = sMax;  // row sums are exchanged between KV passes, row maxima inside them

  // barrier map.  "per item": one phase per KV tile this CTA processes; "per segment": one per unit part
  auto bar_q_full = [&](int t) { return smem_u32(&bars[t]); };           // tx, count 1; per segment
  auto bar_s_full = [&](int t) { return smem_u32(&bars[2 + t]); };       // tcgen05.commit; per item
  auto bar_p_early = [&](int t) { return smem_u32(&bars[4 + t]); };      // 8 softmax warps; per item
  auto bar_p_late = [&](int t) { return smem_u32(&bars[6 + t]); };       // 8 softmax warps; per item
  auto bar_o_final = [&](int t) { return smem_u32(&bars[8 + t]); };      // tcgen05.commit; per segment
  // 8 softmax warps, per segment: "O_t has been read out of tensor memory" - gates the first PV of the
  // next segment
  auto bar_tile_free = [&](int t) { return smem_u32(&bars[10 + t]); };
  auto bar_p_mid = [&](int t) { return smem_u32(&bars[12 + t]); };       // 8 softmax warps; per item
  // tcgen05.commit behind the last S_t product of a segment: the Q_t buffer may be reloaded; per segment
  auto bar_q_free = [&](int t) { return smem_u32(&bars[14 + t]); };
  // The O staging tile is used in turns, numbered k = 2 * (final segments so far) + tile (producer parts
  // do not use it): the 8 softmax warps of the tile fill it and arrive on bar_stage_full (phase k); the store
  // warp (warp 18) issues the TMA store, waits until it has read the tile and arrives on bar_stage_free(tile).
  const uint32_t bar_stage_full = smem_u32(&bars[16]);                   // 8 softmax warps
  // "the store of tile t's turn has read the staging tile": count 1, one phase per final segment
  auto bar_stage_free = [&](int t) { return smem_u32(&bars[t == 0 ? 17 : 22]); };
  // producer part (at most one per CTA): the 8 softmax warps of tile t have stored their partial; the store
  // warp then publishes it (flag, release at gpu scope) while the softmax warps move on
  auto bar_part_done = [&](int t) { return smem_u32(&bars[18 + t]); };   // 8 softmax warps, one phase
  // consumer part (at most one per CTA): the store warp has seen the flags of every producer part of tile t
  auto bar_parts_ready = [&](int t) { return smem_u32(&bars[20 + t]); };  // count 1, one phase
  // Head dim 64 (kEarlyS): tensor memory has 128 spare columns, so P_t gets a region of its own next to O_t instead of
  // overwriting S_t - the protocol of fa_fwd_ws3.cuh: the softmax warps signal "S_t(g) is in registers" (bar_s_read, 8 warps,
  // per item) and the MMA thread issues S_t(g+1) at once, AHEAD of PV_t(g); a commit behind every PV_t(g) (bar_pv_done, per
  // item) frees the P region and is what the rare O rescale waits for.
  auto bar_s_read = [&](int t) { return smem_u32(&bars[23 + t]); };
  auto bar_pv_done = [&](int t) { return smem_u32(&bars[25 + t]); };
  auto bar_kv_full = [&](int s) { return smem_u32(&bars[27 + s]); };     // tx, count 1
  auto bar_kv_empty = [&](int s) { return smem_u32(&bars[27 + kS + s]); };  // tcgen05.commit

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int cta = blockIdx.x;
  const int G = gridDim.x;
  const int T = p.sk_T;
  SkWalker wk;
  wk.init(p, cta, G);
  auto unit_coords = [&](int unit, int& row0, int& hh, int& bb) {
    row0 = (unit % p.sk_P) * 2 * kTileM;
    hh = (unit / p.sk_P) % p.H;
    bb = (unit / p.sk_P) / p.H;
  };

  if (warp == 16 && lane == 0) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_q_full(t), 1);
      mbar_init(bar_s_full(t), 1);
      mbar_init(bar_p_early(t), 8);
      mbar_init(bar_p_late(t), 8);
      mbar_init(bar_o_final(t), 1);
      mbar_init(bar_tile_free(t), 8);
      mbar_init(bar_p_mid(t), 8);
      mbar_init(bar_q_free(t), 1);
      mbar_init(bar_part_done(t), 8);
      mbar_init(bar_parts_ready(t), 1);
      mbar_init(bar_s_read(t), 8);
      mbar_init(bar_pv_done(t), 1);
    }
    mbar_init(bar_stage_full, 8);
    mbar_init(bar_stage_free(0), 1);
    mbar_init(bar_stage_free(1), 1);
#pragma unroll
    for (int s = 0; s < kS; ++s) {
      mbar_init(bar_kv_full(s), 1);
      mbar_init(bar_kv_empty(s), 1);
    }
    fence_mbar_init();
  }
  if (warp == 17 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_o);
    if (wk.more()) {  // first segment's Q tiles and first ring-full of K/V -> L2 before pdl_wait() (fa_fwd_ws.cuh)
      int row0, hh, bb;
      unit_coords(wk.unit(cta, G), row0, hh, bb);
      const int t0 = wk.first(), n = wk.count(T);
#pragma unroll
      for (int db = 0; db < kDBlocks; ++db) {
        tma_prefetch_l2_4d(&tmap_q, db * 64, row0, hh, bb);
        tma_prefetch_l2_4d(&tmap_q, db * 64, row0 + kTileM, hh, bb);
#pragma unroll
        for (int j = 0; j < kS / 2; ++j) {
          if (j < n) {
            tma_prefetch_l2_4d(&tmap_k, db * 64, (t0 + j) * kTileN, hh, bb);
            tma_prefetch_l2_4d(&tmap_v, db * 64, (t0 + j) * kTileN, hh, bb);
          }
        }
      }
    }
  }
  if (warp == 16) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // PDL: everything above overlapped the previous kernel's tail; global memory (inputs, the workspace and
  // its flags, which consecutive launches on a stream share) is touched only below
  pdl_wait();
  pdl_launch_dependents();
  if (*tmem_slot != 0u) __trap();  // see fa_fwd_ws.cuh: constant TMEM addresses
  constexpr uint32_t tmem = 0u;
  const float c = p.scale_log2;

  if (warp >= 16) {
    setmaxnreg_dec<64>();  // 512 x 104 + 128 x 64 = 640 x 96 (the walker state does not fit in 32)
    if (warp == 17) {
      // =======================================================================================
      // TMA producer
      // =======================================================================================
      if (elect_one()) {
        int kvi = 0;  // running K/V ring index (K and V alternate)
        auto load_kv = [&](int x, int t0, int hh, int bb) {  // element x of a segment's K V K V ... sequence
          const int slot = kvi % kS;
          const uint32_t use = kvi / kS;
          mbar_wait(bar_kv_empty(slot), (use & 1) ^ 1, 20);
          mbar_arrive_expect_tx(bar_kv_full(slot), C::kTileBytes);
          const CUtensorMap* map = (x & 1) ? &tmap_v : &tmap_k;
#pragma unroll
          for (int db = 0; db < kDBlocks; ++db)
            tma_load_4d(sKV + slot * C::kTileBytes + db * 16384, map, bar_kv_full(slot), db * 64,
                        (t0 + (x >> 1)) * kTileN, hh, bb);
          ++kvi;
        };
        bool k0_loaded = false;  // the first K tile of this segment went out behind the previous segment's tiles
        for (int seg = 0; wk.more(); ++seg) {
          const int n = wk.count(T);
          const int t0 = wk.first();
          int row0, hh, bb;
          unit_coords(wk.unit(cta, G), row0, hh, bb);
          // Q tiles of this segment; the buffers were last read by the final S products of segment seg - 1
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (seg > 0) mbar_wait(bar_q_free(t), (seg - 1) & 1, 21);
            mbar_arrive_expect_tx(bar_q_full(t), C::kTileBytes);
#pragma unroll
            for (int db = 0; db < kDBlocks; ++db)
              tma_load_4d(sQ + t * C::kTileBytes + db * 16384, &tmap_q, bar_q_full(t), db * 64,
                          row0 + t * kTileM, hh, bb);
          }
#pragma unroll 1
          for (int x = k0_loaded ? 1 : 0; x < 2 * n; ++x) load_kv(x, t0, hh, bb);
          wk.next(T);
          // The next segment's first S needs its K tile as much as its Q tiles, and the Q loads wait for the
          // last S products of this segment: send that K tile first (its ring slot is free long before).
          k0_loaded = wk.more();
          if (k0_loaded) {
            int row0n, hhn, bbn;
            unit_coords(wk.unit(cta, G), row0n, hhn, bbn);
            load_kv(0, wk.first(), hhn, bbn);
          }
        }
      }
      __syncwarp();
    } else if (warp == 16) {
      // =======================================================================================
      // MMA issuer.  One continuous sequence over all KV tiles of the CTA:
      //   S0 S1 | PV0(g) S0(g+1) PV1(g) S1(g+1) | ...
      // where item g+1 may belong to the next segment (its S then reads the next unit's Q tiles).
      // =======================================================================================
      if (elect_one()) {
        constexpr uint32_t idesc_s = make_idesc_f16(kTileM, kTileN, kBF16, false, false);
        constexpr uint32_t idesc_o = make_idesc_f16(kTileM, kDP, kBF16, false, true);
        constexpr uint32_t desc_hi = smem_desc_hi_sw128(1024);
        auto wait_kv = [&](int idx) {
          mbar_wait(bar_kv_full(idx % kS), (idx / kS) & 1, 30);
          tc_fence_after();
        };
        auto release_kv = [&](int idx) { tc_commit(bar_kv_empty(idx % kS)); };
        // S_t = Q_t K^T, K in ring position kidx; `last_of_seg`: no later product of this segment reads Q_t
        auto issue_s = [&](int t, int kidx, bool last_of_seg) {
          const uint32_t k_lo = smem_desc_lo(sKV + (kidx % kS) * C::kTileBytes, 16);
          const uint32_t q_lo = smem_desc_lo(sQ + t * C::kTileBytes, 16);
#pragma unroll
          for (int k = 0; k < kKSteps; ++k) {
            const uint32_t off = ((k >> 2) * 16384 + (k & 3) * 32) >> 4;
            umma_ss2(tmem + col_s(t), q_lo + off, desc_hi, k_lo + off, desc_hi, idesc_s, k > 0);
          }
          tc_commit(bar_s_full(t));
          if (last_of_seg) tc_commit(bar_q_free(t));
        };
        auto pv_step = [&](int t, uint32_t v_lo, int ks, uint32_t acc) {
          const uint32_t p_col = kEarlyS ? col_p(t) + ks * 8 : col_s(t) + (ks >> 2) * 64 + (ks & 3) * 8;
          umma_ts2(tmem + col_o(t), tmem + p_col, v_lo + ((ks * 2048) >> 4), desc_hi, idesc_o, acc);
        };
        // g = running KV-tile count of this CTA (parity source of the per-item barriers)
        auto issue_pv = [&](int t, int vidx, int g, bool first, bool last) {
          const uint32_t v_lo = smem_desc_lo(sKV + (vidx % kS) * C::kTileBytes, 16384);
          mbar_wait(bar_p_early(t), g & 1, 31 + t);
          tc_fence_after();
          pv_step(t, v_lo, 0, first ? 0u : 1u);
          pv_step(t, v_lo, 1, 1);
          pv_step(t, v_lo, 4, 1);
          pv_step(t, v_lo, 5, 1);
          if (kPvParts == 3) {
            mbar_wait(bar_p_mid(t), g & 1, 37 + t);
            tc_fence_after();
            pv_step(t, v_lo, 2, 1);
            pv_step(t, v_lo, 6, 1);
            mbar_wait(bar_p_late(t), g & 1, 35 + t);
            tc_fence_after();
            pv_step(t, v_lo, 3, 1);
            pv_step(t, v_lo, 7, 1);
          } else {
            mbar_wait(bar_p_late(t), g & 1, 35 + t);
            tc_fence_after();
            pv_step(t, v_lo, 2, 1);
            pv_step(t, v_lo, 3, 1);
            pv_step(t, v_lo, 6, 1);
            pv_step(t, v_lo, 7, 1);
          }
          if (kEarlyS) tc_commit(bar_pv_done(t));
          if (last) tc_commit(bar_o_final(t));
        };

        int g = 0;  // KV tiles processed so far by this CTA; ring indices are 2g (K) and 2g+1 (V)
        if (wk.more()) {
          const int n0 = wk.count(T);
          wait_kv(0);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            mbar_wait(bar_q_full(t), 0, 33);
            tc_fence_after();
            issue_s(t, 0, n0 == 1);
          }
          release_kv(0);
        }
        for (int seg = 0; wk.more(); ++seg) {
          const int n = wk.count(T);
          wk.next(T);
          const bool more_segs = wk.more();
          const int n_next = more_segs ? wk.count(T) : 0;
#pragma unroll 1
          for (int j = 0; j < n; ++j, ++g) {
            const bool first = (j == 0), last = (j == n - 1);
            // what follows item g: item j+1 of this segment, item 0 of the next one, or nothing
            const bool has_next = !last || more_segs;
            const bool next_is_new_seg = last && more_segs;
            const bool next_last_s = last ? (n_next == 1) : (j + 1 == n - 1);
            // S_t(g+1): behind PV_t(g), or (kEarlyS) ahead of it as soon as the softmax warps hold S_t(g) in registers
            auto next_s = [&](int t) {
              if (t == 0) wait_kv(2 * g + 2);
              if (next_is_new_seg) {
                mbar_wait(bar_q_full(t), (seg + 1) & 1, 33 + t);
                tc_fence_after();
              }
              if (kEarlyS) {
                mbar_wait(bar_s_read(t), g & 1, 38 + t);
                tc_fence_after();
              }
              issue_s(t, 2 * g + 2, next_last_s);
              if (t == 1) release_kv(2 * g + 2);
            };
            if (kEarlyS && has_next) next_s(0);
            wait_kv(2 * g + 1);
            // O_t of the previous segment must have left TMEM before the first PV overwrites it
            if (first && seg > 0) {
              mbar_wait(bar_tile_free(0), (seg - 1) & 1, 36);
              tc_fence_after();
            }
            issue_pv(0, 2 * g + 1, g, first, last);
            if (has_next) next_s(kEarlyS ? 1 : 0);
            if (first && seg > 0) {
              mbar_wait(bar_tile_free(1), (seg - 1) & 1, 37);
              tc_fence_after();
            }
            issue_pv(1, 2 * g + 1, g, first, last);
            release_kv(2 * g + 1);
            if (!kEarlyS && has_next) next_s(1);
          }
        }
      }
      __syncwarp();
    } else if (warp == 18) {
      // =======================================================================================
      // store warp: O staging tile -> global (TMA), off the softmax warps' critical path
      // =======================================================================================
      if (elect_one()) {
        int k = 0;  // staging turn
        for (; wk.more(); wk.next(T)) {
          const int n = wk.count(T);
          const int t0 = wk.first();
          if (t0 > 0) {
            // producer part: publish each tile's partial once its 8 warps have stored it.  mbarrier arrive /
            // wait order the warps' global stores before this thread's release store, which is cumulative.
#pragma unroll 1
            for (int t = 0; t < 2; ++t) {
              mbar_wait(bar_part_done(t), 0, 59);
              asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.sk_flags + cta * 2 + t), "r"(1)
                           : "memory");
            }
            continue;
          }
          const int unit = wk.unit(cta, G);
          int row0, hh, bb;
          unit_coords(unit, row0, hh, bb);
          int parts = 0;
          if (n < T) {
            // consumer part (the last thing this CTA does): watch the flags of the rest of the unit while the
            // KV loop runs, so that the softmax warps find them raised
            const long long unit_end = static_cast<long long>(unit - p.sk_dp * G + 1) * T;
            for (int c2 = cta + 1; c2 < G && p.sk_W * c2 / G < unit_end; ++c2) ++parts;
#pragma unroll 1
            for (int t = 0; t < 2; ++t) {
              for (int i = 1; i <= parts; ++i) {
                const int* flag = p.sk_flags + (cta + i) * 2 + t;
                int v;
                do {
                  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                } while (v == 0);
              }
              mbar_arrive(bar_parts_ready(t));
            }
          }
#pragma unroll 1
          for (int t = 0; t < 2; ++t, ++k) {
            mbar_wait(bar_stage_full, k & 1, 58);
#pragma unroll
            for (int db = 0; db < kDBlocks; ++db)
              tma_store_4d(&tmap_o, sStage + db * 16384, db * 64, row0 + t * kTileM, hh, bb);
            tma_store_commit();
            // every thread of the tile has read the partials before it arrived: lower the flags
            for (int i = 1; i <= parts; ++i)
              asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p.sk_flags + (cta + i) * 2 + t), "r"(0)
                           : "memory");
            tma_store_wait_read();  // the staging tile may be rewritten once the store has read it
            mbar_arrive(bar_stage_free(t));
          }
        }
      }
      __syncwarp();
    }
  } else {
    // =========================================================================================
    // softmax warps (0-7: tile 0, 8-15: tile 1)
    // =========================================================================================
    setmaxnreg_inc<104>();
    const int t = warp >> 3;
    const int half = (warp >> 2) & 1;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_base + col_s(t) + half * 64;
    const uint32_t tO = tmem + lane_base + col_o(t) + half * kOHalf;
    const int pair_bar = 1 + t * 4 + (warp & 3);
    float* my_max = sMax + (t * 2 + half) * 128 + r;
    const float* other_max = sMax + (t * 2 + (half ^ 1)) * 128 + r;
    [[maybe_unused]] const bool tile_leader = ((warp & 7) == 0) && lane == 0;  // FA_TRACE stamps

    int g = 0;
    int stage_use = 0;  // staging turns taken so far by final (non-producer) segments, two per segment
    FA_SK_TR(0);
    for (int seg = 0; wk.more(); ++seg) {
      const int n = wk.count(T);
      const int t0 = wk.first();
      float m_run = -INFINITY;
      float l_run = 0.f;

      auto kv_step = [&](int j, auto first_tag, auto nomask_tag) {
        constexpr bool kFirstStep = decltype(first_tag)::value;
        constexpr bool kNoMaskStep = decltype(nomask_tag)::value;
        mbar_wait_warp(bar_s_full(t), g & 1, 40 + t);
        tc_fence_after();
        float s[64];
        tmem_ld_x32(tS, reinterpret_cast<uint32_t*>(s));
        tmem_ld_x32(tS + 32, reinterpret_cast<uint32_t*>(s) + 32);
        tmem_wait_ld();
        if constexpr (kEarlyS) {
          Ws3StepArgs a;  // (shared::cta addresses are valid shared::cluster addresses of the executing CTA)
          a.bar_early = bar_p_early(t);
          a.bar_mid = bar_p_mid(t);
          a.bar_late = bar_p_late(t);
          a.bar_s_read = bar_s_read(t);
          a.bar_pv_done = bar_pv_done(t);
          a.p_row = nullptr;
          a.swz = 0;
          a.tP = tmem + lane_base + col_p(t) + half * 32;
          ws3_softmax_step<kDP, kBF16, kFirstStep, kNoMaskStep>(s, tS, tO, lane, (t0 + j) * kTileN + half * 64, p.Nkv, c, m_run, l_run, g,
                                                   my_max, other_max, pair_bar, a, false, 64,
                                                   FA_PEEL_FIRST ? (kFirstStep ? 0 : 1) : (j > 0 ? 1 : 0));
        } else {
          ws_softmax_step<kDP, kBF16, false, kFirstStep, kNoMaskStep>(s, tS, tO, half, r, lane, (t0 + j) * kTileN + half * 64, p.Nkv,
                                                         false, c, m_run, l_run, FA_PEEL_FIRST ? !kFirstStep : (j > 0), my_max,
                                                         other_max, pair_bar, bar_p_early(t), bar_p_late(t), 0u,
                                                         bar_p_mid(t));
        }
        ++g;
      };
      // (non-causal: only a unit's last KV tile can be ragged, and that can only be a segment's last step.  The mask-free
      // interior step is used on the head-dim-64 path only: measured +5.7 % / +6.6 % there at N = 4096 / 8192, but -2.3 % on
      // the head-dim-128 path at N = 8192 and -2 % on config 5 - the second loop body changes ptxas' schedule of the hot one)
#if FA_PEEL_FIRST && FA_PEEL_MASK
      kv_step(0, std::true_type{}, std::false_type{});  // (a segment has at least one KV tile)
      if constexpr (kEarlyS) {
#pragma unroll 1
        for (int j = 1; j < n - 1; ++j) kv_step(j, std::false_type{}, std::true_type{});
        if (n > 1) kv_step(n - 1, std::false_type{}, std::false_type{});
      } else {
#if FA_SK128_NOTAIL_LOOP
        if (p.Nkv % kTileN == 0) {  // no ragged tile anywhere: every further step of the segment is mask-free
#pragma unroll 1
          for (int j = 1; j < n; ++j) kv_step(j, std::false_type{}, std::true_type{});
        } else
#endif
        {
#pragma unroll 1
          for (int j = 1; j < n; ++j) kv_step(j, std::false_type{}, std::false_type{});
        }
      }
#elif FA_PEEL_FIRST
      kv_step(0, std::true_type{}, std::false_type{});  // (a segment has at least one KV tile)
#pragma unroll 1
      for (int j = 1; j < n; ++j) kv_step(j, std::false_type{}, std::false_type{});
#else
#pragma unroll 1
      for (int j = 0; j < n; ++j) kv_step(j, std::false_type{}, std::false_type{});
#endif

      // ---- end of the pass over this unit's KV range.  Everything the epilogue needs is derived here, behind
      // an opaque zero, so that none of it occupies registers during the KV loop.
      FA_SK_TR(1 + 3 * seg);
      const int z = opaque_zero();
      const int rz = r + z;
      // Row sums cross over through the row-max slots (single-buffered here: no shared memory left at head dim 128),
      // each thread writing into its PARTNER's slot and reading its own: the last accesses to the partner's slot were
      // the partner's store before the last pair barrier and my own load after it, so this store is ordered behind
      // both by the barrier and program order (a store into my own slot would only be ordered behind the partner's
      // load of my maximum by a margin of time - compute-sanitizer racecheck reported exactly that pair).
      sFinal[(t * 2 + (half ^ 1)) * 128 + rz] = l_run;
      named_bar_sync(pair_bar, 64);
      float l_tot = l_run + sFinal[(t * 2 + half) * 128 + rz];
      const int kind = (t0 > 0) ? 1 : (n < T ? 2 : 0);
      const int unit = wk.unit(cta, G) + z;
      int row0, hh, bb;
      unit_coords(unit, row0, hh, bb);
      const int tile_row0 = row0 + t * kTileM;
      const int row = tile_row0 + rz;
      // my float4 column q of a partial lives at part_o[q * 128] (see SkSlot)
      const size_t part_o_off = (static_cast<size_t>((t * 2 + half) * kQ4) * 128 + rz) * 4;
      const size_t part_ml_off = SkSlot<kDP>::kOFloats + (t * kTileM + rz) * 2;
      int parts = 0;  // consumer part: the rest of my unit is held by CTAs cta+1 .. cta+parts
      if (kind == 2) {
        const long long unit_end = static_cast<long long>(unit - p.sk_dp * G + 1) * T;
        for (int c2 = cta + 1; c2 < G && p.sk_W * c2 / G < unit_end; ++c2) ++parts;
      }
      // my turn on the O staging tile (producer parts take none)
      const int use = stage_use + t;
      if (kind != 1) stage_use += 2;

      if (kind == 1) {
        // producer: unnormalised O row-half, and (m, l) by the half-0 thread, to my workspace slot
        mbar_wait(bar_o_final(t), seg & 1, 54 + t);
        tc_fence_after();
        float* slot = p.sk_ws + static_cast<size_t>(cta) * SkSlot<kDP>::kFloats;
        float4* o_dst = reinterpret_cast<float4*>(slot + part_o_off);
#pragma unroll
        for (int cidx = 0; cidx < kOHalf / 32; ++cidx) {
          uint32_t o[32];
          tmem_ld_x32(tO + cidx * 32, o);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            __stcg(o_dst + (cidx * 8 + (e >> 2)) * 128,
                   make_float4(__uint_as_float(o[e]), __uint_as_float(o[e + 1]),
                               __uint_as_float(o[e + 2]), __uint_as_float(o[e + 3])));
        }
        // O_t has left tensor memory: the next segment's first PV may overwrite it
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tile_free(t));
        if (half == 0) __stcg(reinterpret_cast<float2*>(slot + part_ml_off), make_float2(m_run, l_tot));
        // the store warp raises the flag (release at gpu scope) once all 8 warps of the tile are here;
        // nobody on the softmax side waits for that
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_part_done(t));
      } else {
        float f_mine = 1.f;
        if (kind == 2) {
          // the partials of the rest of the unit (the store warp watched the flags during the KV loop):
          // common maximum and row sum
          mbar_wait(bar_parts_ready(t), 0, 55);
          FA_SK_TR(2 + 3 * seg);
          float m_all = m_run;
          for (int i = 1; i <= parts; ++i) {
            const float* slot = p.sk_ws + static_cast<size_t>(cta + i) * SkSlot<kDP>::kFloats;
            m_all = fmaxf(m_all, __ldcg(reinterpret_cast<const float2*>(slot + part_ml_off)).x);
          }
          f_mine = ex2_approx((m_run - m_all) * c);
          l_tot *= f_mine;
          for (int i = 1; i <= parts; ++i) {
            const float* slot = p.sk_ws + static_cast<size_t>(cta + i) * SkSlot<kDP>::kFloats;
            const float2 ml = __ldcg(reinterpret_cast<const float2*>(slot + part_ml_off));
            l_tot = fmaf(ml.y, ex2_approx((ml.x - m_all) * c), l_tot);
          }
          m_run = m_all;
        }
        if (half == 0 && p.lse != nullptr && row < p.Nq)
          p.lse[static_cast<int64_t>(unit / p.sk_P) * p.Nq + row] = m_run * c + log2f(l_tot);
        const float inv_l = 1.f / l_tot;
        f_mine *= inv_l;
        mbar_wait(bar_o_final(t), seg & 1, 54 + t);
        tc_fence_after();
        // my turn on the staging tile: the previous user's - the OTHER tile's - TMA store has read it.  One "free"
        // barrier per tile: tile 0 in final segment fs waits for tile 1's turn of segment fs - 1, tile 1 for tile 0's
        // turn of segment fs.  Each tile observed the other's preceding phase one turn earlier, so a parity wait can
        // never mistake an older phase for the one it needs (with a single shared barrier, phases use - 2 and use
        // have the same parity).
        {
          const int fs = use >> 1;
          if (t == 0) {
            if (fs > 0) mbar_wait(bar_stage_free(1), (fs - 1) & 1, 57);
          } else {
            mbar_wait(bar_stage_free(0), fs & 1, 57);
          }
        }
        uint8_t* stage = smem + C::kStage;
#pragma unroll
        for (int cidx = 0; cidx < kOHalf / 32; ++cidx) {
          uint32_t o[32];
          tmem_ld_x32(tO + cidx * 32, o);
          tmem_wait_ld();
          if (cidx == kOHalf / 32 - 1) {  // O_t has left TMEM: the next segment's first PV may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tile_free(t));
          }
          float of[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) of[e] = __uint_as_float(o[e]) * f_mine;
          if (kind == 2) {
            for (int i = 1; i <= parts; ++i) {
              const float* slot = p.sk_ws + static_cast<size_t>(cta + i) * SkSlot<kDP>::kFloats;
              const float f_i =
                  ex2_approx((__ldcg(reinterpret_cast<const float2*>(slot + part_ml_off)).x - m_run) * c) * inv_l;
              const float4* o_src = reinterpret_cast<const float4*>(slot + part_o_off) + cidx * 8 * 128;
#pragma unroll
              for (int e = 0; e < 32; e += 4) {
                const float4 x = __ldcg(o_src + (e >> 2) * 128);
                of[e] = fmaf(x.x, f_i, of[e]);
                of[e + 1] = fmaf(x.y, f_i, of[e + 1]);
                of[e + 2] = fmaf(x.z, f_i, of[e + 2]);
                of[e + 3] = fmaf(x.w, f_i, of[e + 3]);
              }
            }
          }
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint4 val;
            val.x = pack2<kBF16>(of[ch * 8 + 0], of[ch * 8 + 1]);
            val.y = pack2<kBF16>(of[ch * 8 + 2], of[ch * 8 + 3]);
            val.z = pack2<kBF16>(of[ch * 8 + 4], of[ch * 8 + 5]);
            val.w = pack2<kBF16>(of[ch * 8 + 6], of[ch * 8 + 7]);
            *reinterpret_cast<uint4*>(stage + sw128_offset_16bit(rz, half * kOHalf + cidx * 32 + ch * 8)) = val;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_stage_full);  // the store warp takes it from here
      }
      FA_SK_TR(3 + 3 * seg);
      wk.next(T);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 512);
}

}  // namespace fa
