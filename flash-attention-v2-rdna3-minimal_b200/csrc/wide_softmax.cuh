// Software-pipelined softmax step for the one-Q-tile kernels (fa_fwd_wide2.cuh): because the score
// tile is double-buffered there, S(j+1) is usually complete while the softmax warps are still working
// on S(j).  The step therefore
//   - starts with the scores of tile j already in registers and its row maximum already known (computed
//     and exchanged with the partner thread during the PREVIOUS step), so every exponential is taken
//     against the final maximum: no speculation, no redo, and the rare O rescale is decided up front;
//   - loads S(j+1) into a second register array right after the first P hand-off, and reduces / exchanges
//     its row maximum in the shadow of the second half of the exponentials (the MUFU is the busy pipe).
// Motivation: with ws_softmax_step one softmax group needs ~1630 cycles per 128x128 tile even when S is
// always ready, against 896 cycles of MUFU work; the difference is the serial prologue of each tile
// (barrier wait, TMEM load, max, pair exchange), which this arrangement was meant to overlap.
//
// RESULT (B200, fp16, N=16384, pair kernel): it is SLOWER than ws_softmax_step - 607 vs 669 TFLOPS at D=64,
// 1186 vs 1284 at D=128, 1697 vs 1769 at D=256 - and issuing the prefetch before the first half (so that
// the max reduction sits inside the MUFU-bound second half) is worse still (499 / 876 / 1391): two live
// 64-register score arrays and the extra tcgen05.wait::ld points cost more issue slots and scheduling
// freedom than the hidden latency returns.  Kept, off by default (-DFA_WIDE2_PIPELINED=1), as a measured
// dead end; the parity tests pass with it on.
//
// The thread layout, the P hand-off in three parts, the lazy rescale threshold and the exp2 FMA-pipe
// share are those of ws_softmax_step (fa_fwd_ws.cuh), whose header comment explains them.
#pragma once
#include "fa_fwd_ws.cuh"

namespace fa {

// p = 2^(s*c + nmc) for columns [i, i+4): kEmuPairs of every 8 element pairs on the FMA pipes
__device__ __forceinline__ void wide_exp4(float (&s)[64], int i, float c, float nmc) {
  ffma2(s[i], s[i + 1], s[i], s[i + 1], c, c, nmc, nmc);
  ffma2(s[i + 2], s[i + 3], s[i + 2], s[i + 3], c, c, nmc, nmc);
  if ((((i >> 1) * kEmuPairs) & 7) < kEmuPairs) {
    ex2_fma2(s[i], s[i + 1]);
  } else {
    s[i] = ex2_approx(s[i]);
    s[i + 1] = ex2_approx(s[i + 1]);
  }
  if (((((i >> 1) + 1) * kEmuPairs) & 7) < kEmuPairs) {
    ex2_fma2(s[i + 2], s[i + 3]);
  } else {
    s[i + 2] = ex2_approx(s[i + 2]);
    s[i + 3] = ex2_approx(s[i + 3]);
  }
}

// mask my 64 columns of a tile (KV tail, causal) and return their maximum
//   col0   index of the first key of my half;   r_lim  row limit for the causal mask: columns i with
//   i >= r_lim - (my half's offset inside the tile) are hidden (r_lim <= 0 hides the whole tile)
__device__ __forceinline__ float wide_mask_max(float (&s)[64], int col0, int Nkv, bool causal_tile, int r_lim_half) {
  const bool tail = (col0 + 64 > Nkv);
  if (tail || causal_tile) {
    const int valid = tail ? (Nkv - col0) : 64;
    const int lim = causal_tile ? min(valid, r_lim_half) : valid;
#pragma unroll
    for (int i = 0; i < 64; ++i)
      if (i >= lim) s[i] = -INFINITY;
  }
  float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    mx0 = fmaxf(mx0, fmaxf(s[i], s[i + 32]));
    mx1 = fmaxf(mx1, fmaxf(s[i + 1], s[i + 33]));
    mx2 = fmaxf(mx2, fmaxf(s[i + 2], s[i + 34]));
    mx3 = fmaxf(mx3, fmaxf(s[i + 3], s[i + 35]));
  }
  return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
}

struct WideStepArgs {
  uint32_t tS;        // TMEM address of my 64 S columns of tile j (P goes over the first 32)
  uint32_t tS_next;   // same for tile j+1 (the other S buffer)
  uint32_t tO;        // my O columns
  uint32_t bar_early, bar_mid, bar_late;  // P hand-off barriers of tile j's buffer
  uint32_t bar_s_next;                    // "S(j+1) is ready" (this CTA's copy), with its parity
  uint32_t s_next_parity;
  uint32_t bar_o, o_parity;               // "PV(j-1) has left the tensor cores"
  bool has_next, have_o;
  int next_col0;        // first key of my half in tile j+1
  bool next_causal;     // tile j+1 needs the causal mask
  int next_r_lim_half;  // its row limit for my half
  int Nkv;
  float c;
  float* my_max;        // exchange slots for tile j+1's row maximum (parity (j+1)&1)
  const float* other_max;
  int pair_bar;
};

// Blocking fetch of a tile: wait for S, load my 64 columns, mask, row maximum over both halves.
__device__ __forceinline__ float wide_fetch_tile(float (&s)[64], uint32_t tS, uint32_t bar_s, uint32_t parity,
                                                 int col0, int Nkv, bool causal_tile, int r_lim_half,
                                                 float* my_max, const float* other_max, int pair_bar) {
  mbar_wait_warp(bar_s, parity, 40);
  tc_fence_after();
  tmem_ld_x32(tS, reinterpret_cast<uint32_t*>(s));
  tmem_ld_x32(tS + 32, reinterpret_cast<uint32_t*>(s) + 32);
  tmem_wait_ld();
  const float mx = wide_mask_max(s, col0, Nkv, causal_tile, r_lim_half);
  *my_max = mx;
  named_bar_sync(pair_bar, 64);
  return fmaxf(mx, *other_max);
}

// One step: `cur` = raw (masked) scores of tile j, `mx_cur` = its row maximum over both halves.
// Returns true if `nxt` / `mx_next` hold the same for tile j+1 on return: the prefetch happens only when
// S(j+1) is already complete at the first hand-off (always, when the softmax is the slower side); when the
// tensor cores are the slower side, waiting for S(j+1) here would hold back the rest of P(j), so the
// caller fetches tile j+1 at the start of the next step instead (wide_fetch_tile).
template <int kDP, bool kBF16, bool kPairArrive>
__device__ __forceinline__ bool wide_softmax_step(float (&cur)[64], float (&nxt)[64], float mx_cur, float& mx_next,
                                                  float& m_run, float& l_run, int lane, const WideStepArgs& a) {
  constexpr int kOHalf = kDP / 2;
  auto arrive = [](uint32_t bar) {
    if constexpr (kPairArrive) mbar_arrive_cluster(bar); else mbar_arrive(bar);
  };
  const float c = a.c;

  // ---- lazy rescale, decided before any exponential (both threads of the row see the same numbers)
  const float m_cand = fmaxf(mx_cur, m_run);
  const bool grow = (m_cand - m_run) * c > kRescaleThreshold;  // always true on the first tile
  float alpha = 1.f;
  if (__any_sync(0xffffffffu, grow)) {
    if (grow) {
      alpha = ex2_approx((m_run - m_cand) * c);
      m_run = m_cand;
    }
    if (a.have_o) {
      mbar_wait(a.bar_o, a.o_parity, 44);  // S(j) was issued before PV(j-1): wait for PV(j-1) itself
      tc_fence_after();
#pragma unroll 1
      for (int c8 = 0; c8 < kOHalf; c8 += 8) {
        uint32_t o[8];
        tmem_ld_x8(a.tO + c8, o);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st_x8(a.tO + c8, o);
      }
    }
  }
  const float nmc = -m_run * c;

  // ---- columns [0,32): exponentials -> P -> "early" hand-off
#pragma unroll
  for (int i = 0; i < 32; i += 4) wide_exp4(cur, i, c, nmc);
  {
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = pack2<kBF16>(cur[2 * i], cur[2 * i + 1]);
    tmem_st_x16(a.tS, pk);
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) arrive(a.bar_early);
  }

  // ---- prefetch the scores of tile j+1 (the other S buffer) if they are complete already; consumed
  // after the late hand-off
  const bool pre = a.has_next && __all_sync(0xffffffffu, mbar_try_wait(a.bar_s_next, a.s_next_parity));
  if (pre) {
    tc_fence_after();
    tmem_ld_x32(a.tS_next, reinterpret_cast<uint32_t*>(nxt));
    tmem_ld_x32(a.tS_next + 32, reinterpret_cast<uint32_t*>(nxt) + 32);
  }

  // ---- columns [32,64), the row sum of the first half in the MUFU shadow, "mid" and "late" hand-offs
  float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
  {
    uint32_t pk[16];
#pragma unroll
    for (int i = 32; i < 64; i += 4) {
      wide_exp4(cur, i, c, nmc);
      fadd2(sum0, sum1, sum0, sum1, cur[i - 32], cur[i - 31]);
      fadd2(sum2, sum3, sum2, sum3, cur[i - 30], cur[i - 29]);
      pk[(i - 32) >> 1] = pack2<kBF16>(cur[i], cur[i + 1]);
      pk[((i - 32) >> 1) + 1] = pack2<kBF16>(cur[i + 2], cur[i + 3]);
      if (i == 44) {  // columns [32,48) leave as the "mid" part
        uint32_t lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) lo[e] = pk[e];
        tmem_st_x8(a.tS + 16, lo);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive(a.bar_mid);
      }
    }
    uint32_t hi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) hi[e] = pk[8 + e];
    tmem_st_x8(a.tS + 24, hi);
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) arrive(a.bar_late);
  }

  // ---- row maximum of tile j+1, exchanged with the partner thread of my row
  if (pre) {
    tmem_wait_ld();
    const float mx = wide_mask_max(nxt, a.next_col0, a.Nkv, a.next_causal, a.next_r_lim_half);
    *a.my_max = mx;
    named_bar_sync(a.pair_bar, 64);
    mx_next = fmaxf(mx, *a.other_max);
  }
#pragma unroll
  for (int i = 32; i < 64; i += 4) {
    fadd2(sum0, sum1, sum0, sum1, cur[i], cur[i + 1]);
    fadd2(sum2, sum3, sum2, sum3, cur[i + 2], cur[i + 3]);
  }
  l_run = l_run * alpha + ((sum0 + sum1) + (sum2 + sum3));
  return pre;
}

}  // namespace fa
