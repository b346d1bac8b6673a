// Persistent "stream-K" variant of the warp-specialised forward kernel (non-causal, Nq % 256 == 0).
//
// Why: fa_fwd_ws_kernel runs one CTA per 256-row query block, so a launch with U blocks on G = 148 SMs
// takes ceil(U / G) rounds: B=1, H=16 gives 256 blocks at N=4096 (2 rounds for 1.73 rounds of work,
// 15 % lost) and 512 at N=8192 (4 rounds for 3.46).  Here the work is linearised as
// (unit, KV tile) items, unit = (batch, head, 256-row block) with T = ceil(Nkv / 128) items each, and
// the grid is G persistent CTAs.  The first floor(U / G) - 1 rounds are data-parallel (CTA c takes unit
// round * G + c: the CTAs of a round work on neighbouring heads, so K/V stays L2-resident as in the
// one-shot kernel - splitting ALL units contiguously made every CTA stream a different head and was
// HBM-bound at N=16384).  The last G + U % G units are split evenly: CTA c takes the contiguous range
// [c W / G, (c+1) W / G) of their W items, so every SM gets the same number of KV tiles (+-1); and
// every CTA pays the prologue (TMEM allocation, barrier initialisation, descriptor prefetch) once.  A range covers a tail part of one unit, whole units,
// and a head part of another (ranges are at least T long, checked by the launcher), so a unit is
// shared by at most two CTAs:
//   * the CTA that owns the TAIL part (KV tiles t0..T-1, the first thing it does) stores its
//     unnormalised O, m and l to a per-CTA workspace slot and raises a flag;
//   * the CTA that owns the HEAD part (KV tiles 0..t1-1, the last thing it does) waits for the flag of
//     the next CTA, merges  O = O_a 2^((m_a-M)c) + O_b 2^((m_b-M)c),  l likewise, M = max(m_a, m_b),
//     normalises, stores, and lowers the flag again (so launches and CUDA-graph replays start clean).
// No deadlock: producers never wait on anything; consumers wait at the very end of their range.
//
// Everything per KV tile - roles, barriers, TMEM layout, the softmax step - is that of
// fa_fwd_ws.cuh; barrier parities run on counters that continue across the units of a CTA.
#pragma once
#include "fa_fwd_ws.cuh"

namespace fa {

// workspace slot of one CTA: unnormalised O of both tiles (fp32), then (m, l) per row
template <int kDP>
struct SkSlot {
  static constexpr int kOFloats = 2 * kTileM * kDP;
  static constexpr int kFloats = kOFloats + 2 * kTileM * 2;
};

// Shared memory: Q tiles (optionally double-buffered, see FA_SK_QDOUBLE), the K/V ring, barriers and
// a single-buffered row-max exchange.  No alignment slack: the dynamic shared-memory window starts
// 1024-byte aligned (checked at kernel start).
// FA_SK_QDOUBLE=1 double-buffers the Q tiles (next unit's Q and first S under the current unit's
// epilogue) at the price of a 3-deep instead of 4-deep K/V ring at D = 128.  Measured with the
// three-part P hand-off: the deeper ring wins by 1-3 % at every sweep length, so the default is 0.
#ifndef FA_SK_QDOUBLE
#define FA_SK_QDOUBLE 0
#endif
constexpr bool kSkQDouble = FA_SK_QDOUBLE != 0;

template <int kDP>
struct SkCfg {
  static constexpr int kTileBytes = kTileM * kDP * 2;
  static constexpr int kStages = (kDP == 128) ? (kSkQDouble ? 3 : 4) : 8;
  static constexpr int kQ = 0;                          // [2 buffers][2 tiles]; also O staging
  static constexpr int kKV = kQ + (kSkQDouble ? 4 : 2) * kTileBytes;
  static constexpr int kBars = kKV + kStages * kTileBytes;
  static constexpr int kNumBars = 18 + 2 * kStages;
  static constexpr int kMax = kBars + 8 * kNumBars + 16;  // float [2 tile][2 half][128]; also row sums
  static constexpr int kTotal = kMax + 2 * 2 * 128 * 4;
};

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
  auto col_s = [](int t) -> uint32_t { return static_cast<uint32_t>(t) * 128u; };
  auto col_o = [](int t) -> uint32_t { return 256u + static_cast<uint32_t>(t) * 128u; };

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();  // SkCfg has no alignment slack
  const uint32_t sQ = smem_u32(smem + C::kQ);
  const uint32_t sKV = smem_u32(smem + C::kKV);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::kBars + 8 * C::kNumBars);
  float* sMax = reinterpret_cast<float*>(smem + C::kMax);
  float* sFinal = sMax;  // row sums are exchanged between KV passes, row maxima inside them

  // barrier map
  auto bar_q_full = [&](int buf, int t) { return smem_u32(&bars[12 + buf * 2 + t]); };  // tx, count 1
  auto bar_s_full = [&](int t) { return smem_u32(&bars[2 + t]); };      // tcgen05.commit
  auto bar_p_early = [&](int t) { return smem_u32(&bars[4 + t]); };     // 8 softmax warps
  auto bar_p_late = [&](int t) { return smem_u32(&bars[6 + t]); };      // 8 softmax warps
  auto bar_o_final = [&](int t) { return smem_u32(&bars[8 + t]); };     // tcgen05.commit
  // per tile, 8 softmax warps: "O_t has been read out of TMEM and the Q_t buffer (O staging) is
  // free again" - gates the Q load and the first PV of the next unit
  auto bar_tile_free = [&](int t) { return smem_u32(&bars[10 + t]); };
  auto bar_p_mid = [&](int t) { return smem_u32(&bars[16 + 2 * kS + t]); };  // 8 softmax warps
  auto bar_kv_full = [&](int s) { return smem_u32(&bars[16 + s]); };    // tx, count 1
  auto bar_kv_empty = [&](int s) { return smem_u32(&bars[16 + kS + s]); };  // tcgen05.commit

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int cta = blockIdx.x;
  const int G = gridDim.x;
  const int T = p.sk_T;
  // Segment walker, the same sequence in every role (32-bit state).  First sk_dp whole units in
  // round-robin order (unit = round * G + cta: the CTAs of a round share heads, so K/V tiles are
  // re-used in L2 exactly as in the one-shot kernel), then the stream-K range of this CTA over the
  // remaining units.  kind 0 = whole unit, 1 = tail part (producer), 2 = head part (consumer).
  int w_dp = p.sk_dp, w_round = 0;
  int w_u, w_t0, w_rem;
  {
    const long long pos_begin = p.sk_W * cta / G;
    const long long pos_end = p.sk_W * (cta + 1) / G;
    const int u_rel = static_cast<int>(pos_begin / T);
    w_u = p.sk_dp * G + u_rel;
    w_t0 = static_cast<int>(pos_begin - static_cast<long long>(u_rel) * T);
    w_rem = static_cast<int>(pos_end - pos_begin);
  }
  auto seg_more = [&]() { return w_dp > 0 || w_rem > 0; };
  auto seg_unit = [&]() { return w_dp > 0 ? w_round * G + cta : w_u; };
  auto seg_t0 = [&]() { return w_dp > 0 ? 0 : w_t0; };
  auto seg_n = [&]() { return w_dp > 0 ? T : min(T - w_t0, w_rem); };
  auto seg_next = [&](int n) {
    if (w_dp > 0) {
      --w_dp;
      ++w_round;
    } else {
      w_rem -= n;
      w_u += 1;
      w_t0 = 0;
    }
  };

  if (warp == 16 && lane == 0) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_q_full(0, t), 1);
      mbar_init(bar_q_full(1, t), 1);
      mbar_init(bar_s_full(t), 1);
      mbar_init(bar_p_early(t), 8);
      mbar_init(bar_p_late(t), 8);
      mbar_init(bar_o_final(t), 1);
      mbar_init(bar_tile_free(t), 8);
      mbar_init(bar_p_mid(t), 8);
    }
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
  }
  if (warp == 16) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (*tmem_slot != 0u) __trap();  // see fa_fwd_ws.cuh: constant TMEM addresses
  constexpr uint32_t tmem = 0u;
  const float c = p.scale_log2;

  if (warp >= 16) {
    setmaxnreg_dec<56>();  // 512 x 104 + 128 x 56 <= 640 x 96 (the walker state does not fit in 32)
    if (warp == 17) {
      // =======================================================================================
      // TMA producer
      // =======================================================================================
      if (elect_one()) {
        int kvi = 0;  // running K/V ring index (K and V alternate)
        // Q tiles of segment `sgi` (its unit is `unit`) -> buffer sgi & 1.  The buffer was the O
        // staging of segment sgi - 2: wait until that store has read it.
        auto load_q = [&](int sgi, int unit) {
          const int buf = kSkQDouble ? (sgi & 1) : 0;
          const int lag = kSkQDouble ? 2 : 1;  // the buffer was the O staging of segment sgi - lag
          const int row0 = (unit % p.sk_P) * 2 * kTileM;
          const int hh = (unit / p.sk_P) % p.H;
          const int bb = (unit / p.sk_P) / p.H;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (sgi >= lag) mbar_wait(bar_tile_free(t), (sgi - lag) & 1, 21);
            mbar_arrive_expect_tx(bar_q_full(buf, t), C::kTileBytes);
#pragma unroll
            for (int db = 0; db < kDBlocks; ++db)
              tma_load_4d(sQ + (buf * 2 + t) * C::kTileBytes + db * 16384, &tmap_q, bar_q_full(buf, t),
                          db * 64, row0 + t * kTileM, hh, bb);
          }
        };
        if (kSkQDouble && seg_more()) load_q(0, seg_unit());
        for (int seg = 0; seg_more(); ++seg) {
          const int n = seg_n();
          const int t0 = seg_t0();
          const int unit = seg_unit();
          const int hh = (unit / p.sk_P) % p.H;
          const int bb = (unit / p.sk_P) / p.H;
          auto load_kv = [&](int x) {  // ring order K V K V ...
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
          // a ring-full of this unit's K/V first, then the NEXT unit's Q (its buffer frees up when
          // the previous unit's epilogue, which runs as this unit starts, has been stored)
          const int pre = min(2 * n, kS);
#pragma unroll 1
          for (int x = 0; x < pre; ++x) load_kv(x);
          if (!kSkQDouble) load_q(seg, unit);
          seg_next(n);
          if (kSkQDouble && seg_more()) load_q(seg + 1, seg_unit());
#pragma unroll 1
          for (int x = pre; x < 2 * n; ++x) load_kv(x);
        }
      }
      __syncwarp();
    } else if (warp == 16) {
      // =======================================================================================
      // MMA issuer
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
        auto issue_s = [&](int qbuf, int t, int kidx) {  // S_t = Q_t K^T, K in ring position kidx
          const uint32_t k_lo = smem_desc_lo(sKV + (kidx % kS) * C::kTileBytes, 16);
          const uint32_t q_lo = smem_desc_lo(sQ + (qbuf * 2 + t) * C::kTileBytes, 16);
#pragma unroll
          for (int k = 0; k < kKSteps; ++k) {
            const uint32_t off = ((k >> 2) * 16384 + (k & 3) * 32) >> 4;
            umma_ss2(tmem + col_s(t), q_lo + off, desc_hi, k_lo + off, desc_hi, idesc_s, k > 0);
          }
          tc_commit(bar_s_full(t));
        };
        auto pv_step = [&](int t, uint32_t v_lo, int ks, uint32_t acc) {
          umma_ts2(tmem + col_o(t), tmem + col_s(t) + (ks >> 2) * 64 + (ks & 3) * 8,
                   v_lo + ((ks * 2048) >> 4), desc_hi, idesc_o, acc);
        };
        // g = running KV-tile count of this CTA (parity source of the per-tile barriers)
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
          if (last) tc_commit(bar_o_final(t));
        };

        int g = 0;  // KV tiles processed so far by this CTA; ring indices are 2g (K) and 2g+1 (V)
        for (int seg = 0; seg_more(); ++seg) {
          const int n = seg_n();
          const int qb = kSkQDouble ? (seg & 1) : 0;
          wait_kv(2 * g);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            mbar_wait(bar_q_full(qb, t), kSkQDouble ? ((seg >> 1) & 1) : (seg & 1), 33);
            tc_fence_after();
            issue_s(qb, t, 2 * g);
          }
          release_kv(2 * g);
#pragma unroll 1
          for (int j = 0; j < n; ++j, ++g) {
            const bool first = (j == 0), last = (j == n - 1);
            wait_kv(2 * g + 1);
            // O_t of the previous unit must have left TMEM before the first PV overwrites it
            if (first && seg > 0) {
              mbar_wait(bar_tile_free(0), (seg - 1) & 1, 36);
              tc_fence_after();
            }
            issue_pv(0, 2 * g + 1, g, first, last);
            if (!last) {
              wait_kv(2 * g + 2);
              issue_s(qb, 0, 2 * g + 2);
            }
            if (first && seg > 0) {
              mbar_wait(bar_tile_free(1), (seg - 1) & 1, 37);
              tc_fence_after();
            }
            issue_pv(1, 2 * g + 1, g, first, last);
            release_kv(2 * g + 1);
            if (!last) {
              issue_s(qb, 1, 2 * g + 2);
              release_kv(2 * g + 2);
            }
          }
          seg_next(n);
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
    const int tile_bar = 9 + t;  // the 8 warps of my tile
    float* my_max = sMax + (t * 2 + half) * 128 + r;
    const float* other_max = sMax + (t * 2 + (half ^ 1)) * 128 + r;
    const bool tile_leader = ((warp & 7) == 0) && lane == 0;

    int g = 0;
    for (int seg = 0; seg_more(); ++seg) {
      const int n = seg_n();
      const int t0 = seg_t0();
      const int kind = (t0 > 0) ? 1 : (n < T ? 2 : 0);
      const int unit = seg_unit();
      const int tile_row0 = (unit % p.sk_P) * 2 * kTileM + t * kTileM;
      float m_run = -INFINITY;
      float l_run = 0.f;

#pragma unroll 1
      for (int j = 0; j < n; ++j, ++g) {
        mbar_wait_warp(bar_s_full(t), g & 1, 40 + t);
        tc_fence_after();
        float s[64];
        tmem_ld_x32(tS, reinterpret_cast<uint32_t*>(s));
        tmem_ld_x32(tS + 32, reinterpret_cast<uint32_t*>(s) + 32);
        tmem_wait_ld();
        ws_softmax_step<kDP, kBF16>(s, tS, tO, half, r, lane, (t0 + j) * kTileN + half * 64, p.Nkv,
                                    false, c, m_run, l_run, j > 0, my_max, other_max, pair_bar, bar_p_early(t),
                                    bar_p_late(t), 0u, bar_p_mid(t));
      }

      // ---- end of the pass over this unit's KV range
      sFinal[(t * 2 + half) * 128 + r] = l_run;
      named_bar_sync(pair_bar, 64);
      float l_tot = l_run + sFinal[(t * 2 + (half ^ 1)) * 128 + r];
      const int row = tile_row0 + r;
      mbar_wait(bar_o_final(t), seg & 1, 54 + t);
      tc_fence_after();

      if (kind == 1) {
        // producer: unnormalised O row-half, and (m, l) by the half-0 thread, to my workspace slot
        float* slot = p.sk_ws + static_cast<size_t>(cta) * SkSlot<kDP>::kFloats;
        float* o_dst = slot + (t * kTileM + r) * kDP + half * kOHalf;
#pragma unroll
        for (int cidx = 0; cidx < kOHalf / 32; ++cidx) {
          uint32_t o[32];
          tmem_ld_x32(tO + cidx * 32, o);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            __stcg(reinterpret_cast<float4*>(o_dst + cidx * 32 + e),
                   make_float4(__uint_as_float(o[e]), __uint_as_float(o[e + 1]),
                               __uint_as_float(o[e + 2]), __uint_as_float(o[e + 3])));
        }
        if (half == 0) {
          float* ml = slot + SkSlot<kDP>::kOFloats + (t * kTileM + r) * 2;
          __stcg(reinterpret_cast<float2*>(ml), make_float2(m_run, l_tot));
        }
        __threadfence();
        tc_fence_before();
        named_bar_sync(tile_bar, 256);
        if (tile_leader) {
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.sk_flags + cta * 2 + t), "r"(1)
                       : "memory");
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tile_free(t));
      } else {
        float scale_mine = 1.f, scale_other = 0.f;
        const float* o_src = nullptr;
        if (kind == 2) {
          // consumer: wait for the partial of the tail part (owned by the next CTA), merge
          int* flag = p.sk_flags + (cta + 1) * 2 + t;
          if (tile_leader) {
            int v;
            do {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            } while (v == 0);
          }
          named_bar_sync(tile_bar, 256);
          const float* slot = p.sk_ws + static_cast<size_t>(cta + 1) * SkSlot<kDP>::kFloats;
          const float2 ml = __ldcg(reinterpret_cast<const float2*>(slot + SkSlot<kDP>::kOFloats +
                                                                   (t * kTileM + r) * 2));
          const float m_new = fmaxf(m_run, ml.x);
          scale_mine = ex2_approx((m_run - m_new) * c);
          scale_other = ex2_approx((ml.x - m_new) * c);
          l_tot = l_tot * scale_mine + ml.y * scale_other;
          m_run = m_new;
          o_src = slot + (t * kTileM + r) * kDP + half * kOHalf;
        }
        if (half == 0 && p.lse != nullptr && row < p.Nq)
          p.lse[static_cast<int64_t>(unit / p.sk_P) * p.Nq + row] = m_run * c + log2f(l_tot);
        const float inv_l = 1.f / l_tot;
        const float f_mine = scale_mine * inv_l, f_other = scale_other * inv_l;
        const int qb = kSkQDouble ? (seg & 1) : 0;
        uint8_t* stage = smem + C::kQ + (qb * 2 + t) * C::kTileBytes;
#pragma unroll
        for (int cidx = 0; cidx < kOHalf / 32; ++cidx) {
          uint32_t o[32];
          tmem_ld_x32(tO + cidx * 32, o);
          tmem_wait_ld();
          float of[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) of[e] = __uint_as_float(o[e]) * f_mine;
          if (kind == 2) {
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const float4 x = __ldcg(reinterpret_cast<const float4*>(o_src + cidx * 32 + e));
              of[e] = fmaf(x.x, f_other, of[e]);
              of[e + 1] = fmaf(x.y, f_other, of[e + 1]);
              of[e + 2] = fmaf(x.z, f_other, of[e + 2]);
              of[e + 3] = fmaf(x.w, f_other, of[e + 3]);
            }
          }
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint4 val;
            val.x = pack2<kBF16>(of[ch * 8 + 0], of[ch * 8 + 1]);
            val.y = pack2<kBF16>(of[ch * 8 + 2], of[ch * 8 + 3]);
            val.z = pack2<kBF16>(of[ch * 8 + 4], of[ch * 8 + 5]);
            val.w = pack2<kBF16>(of[ch * 8 + 6], of[ch * 8 + 7]);
            *reinterpret_cast<uint4*>(stage + sw128_offset_16bit(r, half * kOHalf + cidx * 32 + ch * 8)) = val;
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        named_bar_sync(tile_bar, 256);
        if (tile_leader) {
#pragma unroll
          for (int db = 0; db < kDBlocks; ++db)
            tma_store_4d(&tmap_o, sQ + (qb * 2 + t) * C::kTileBytes + db * 16384, db * 64, tile_row0,
                         (unit / p.sk_P) % p.H, (unit / p.sk_P) / p.H);
          tma_store_commit();
          tma_store_wait_read();  // the Q_t buffer may be reloaded once the store has read it
          if (kind == 2) {        // every thread of the tile has read the partial (barrier above)
            asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p.sk_flags + (cta + 1) * 2 + t),
                         "r"(0)
                         : "memory");
          }
        }
        __syncwarp();
        // warp 0 of the tile arrives after its leader's wait_read: Q_t / O_t are free for the next unit
        if (lane == 0) mbar_arrive(bar_tile_free(t));
      }
      seg_next(n);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 512);
}

}  // namespace fa
