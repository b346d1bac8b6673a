// One-Q-tile forward kernel on CTA pairs with FOUR threads per query row ("quad2"): head dims <= 128.
//
// The pair arrangement is that of fa_fwd_wide2.cuh (cluster of two CTAs, tcgen05 cta_group::2, each CTA
// one 128-row Q tile with its S tile double-buffered in tensor memory, half of every K/V tile per SM,
// the leader CTA issuing every MMA).  What changes is the softmax side: with two threads per row a single
// softmax group needs ~1630 cycles per 128x128 tile even when S is always ready - 896 cycles of MUFU work
// plus the serial prologue of a 64-element stream - and one group per SM is then the limit (1284 TFLOPS
// at D=128).  Here 16 softmax warps share the one tile: warp w owns query rows 32 (w % 4).. and the 32-key
// QUARTER w / 4 of every S tile, so four threads on the same SM sub-partition share a row, each with a
// 32-element stream (max, exp2, pack), and four warps per sub-partition keep the MUFU fed.
//
//   warps 0-15  softmax (quarter q = w / 4); the four partial row maxima meet through shared memory
//   warp  16    MMA issuer (leader CTA only)        warp 17   TMA producer (each CTA)
//
// TMEM: S buffer 0 [0,128)  S buffer 1 [128,256)  O [256,256+D).  P (16 bit) of quarter q overwrites S
// columns [32q, 32q+16); it is handed over in two parts (first / second 16 keys of every quarter), i.e.
// k-steps {0,2,4,6} then {1,3,5,7} of O += P V.
// Tensor-core issue order: S(0) S(1) | PV(0) S(2) | PV(1) S(3) | ...  (see fa_fwd_wide.cuh).
//
// Replaces /root/reference/rocwmma_fattn/kernel_fp16.cu:306-544 (fwd_kernel) for head dims <= 128.
#pragma once
#include "fa_fwd_wide2.cuh"

// Debug-only timeline (-DFA_TRACE): lane 0 of softmax warp 0 (role 0), of softmax warp 5 (role 1) and the
// MMA thread (role 2) of one leader CTA stamp clock64() into p.trace[(role * 128 + j % 128) * 8 + event];
// tools/trace_quad2.py prints the averages.
#ifdef FA_TRACE
#define FA_QTR(role, j, ev)                                                                        \
  do {                                                                                             \
    if (tr_cta && lane == 0) p.trace[((role) * 128 + ((j) & 127)) * 8 + (ev)] = clock64();          \
  } while (0)
#else
#define FA_QTR(role, j, ev) do { } while (0)
#endif

namespace fa {

constexpr int kQuadThreads = 576;  // 16 softmax warps + MMA warp + TMA warp

template <int kDP_>
struct Quad2Cfg {
  static_assert(kDP_ == 64 || kDP_ == 128, "quad2 kernel: padded head dim 64 or 128");
  static constexpr int kDP = kDP_;
  static constexpr int kQBytes = kTileM * kDP * 2;
  static constexpr int kKHalfBytes = (kTileN / 2) * kDP * 2;   // 64 keys x kDP: kDP/64 blocks of 8 KB
  static constexpr int kVHalfBytes = kTileN * 64 * 2;          // 128 keys x kDP/2 columns in one 64-column block
  static constexpr int kSlotBytes = 16384;
  static constexpr int kStages = 8;
  static constexpr int kQ = 0;
  static constexpr int kKV = kQ + kQBytes;
  static constexpr int kBars = kKV + kStages * kSlotBytes;
  static constexpr int kNumBars = 9 + 2 * kStages;
  static constexpr int kMax = kBars + 8 * kNumBars + 16;   // float [2 parity][4 quarters][128]
  static constexpr int kFinal = kMax + 2 * 4 * 128 * 4;    // float [4 quarters][128] row sums
  static constexpr int kTotal = kFinal + 4 * 128 * 4 + 1024;
};

// One softmax step of one thread: `s` = its 32 raw scores of the current S tile.
//   tS        TMEM address of my 32 S columns (P goes over the first 16)
//   tO        my kDP/4 O columns
//   col0      index of the first key of my quarter;  lim_c  causal limit: columns i >= lim_c are hidden
//   maxes     4 exchange slots of my row (stride 128 floats), mine is maxes[128 * quarter]
template <int kDP, bool kBF16>
__device__ __forceinline__ void quad_softmax_step(float (&s)[32], uint32_t tS, uint32_t tO, int quarter, int lane,
                                                  int col0, int Nkv, bool causal_tile, int lim_c, float c,
                                                  float& m_run, float& l_run, bool have_o, float* maxes,
                                                  int group_bar, uint32_t bar_early, uint32_t bar_late,
                                                  uint32_t bar_o, uint32_t o_parity) {
  constexpr int kOQ = kDP / 4;
  const bool tail = (col0 + 32 > Nkv);
  const bool masked = tail || causal_tile;
  int lim = 32;
  if (masked) {
    const int valid = tail ? (Nkv - col0) : 32;
    lim = causal_tile ? min(valid, lim_c) : valid;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i >= lim) s[i] = -INFINITY;
  }
  auto exp4 = [&](int i, float nmc_) {
    ffma2(s[i], s[i + 1], s[i], s[i + 1], c, c, nmc_, nmc_);
    ffma2(s[i + 2], s[i + 3], s[i + 2], s[i + 3], c, c, nmc_, nmc_);
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
  };

  // columns [0,16) against the max of the previous tiles while this tile's max is reduced and exchanged
  // (exact whenever the lazy-rescale rule keeps m_run; otherwise redone below)
  float nmc = -m_run * c;
  float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    mx0 = fmaxf(mx0, fmaxf(s[i], s[i + 16]));
    mx1 = fmaxf(mx1, fmaxf(s[i + 1], s[i + 17]));
    mx2 = fmaxf(mx2, fmaxf(s[i + 2], s[i + 18]));
    mx3 = fmaxf(mx3, fmaxf(s[i + 3], s[i + 19]));
    exp4(i, nmc);
  }
  const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
  maxes[128 * quarter] = mx;
  named_bar_sync(group_bar, 128);
  const float m_cand = fmaxf(fmaxf(fmaxf(maxes[0], maxes[128]), fmaxf(maxes[256], maxes[384])), m_run);
  const bool grow = (m_cand - m_run) * c > kRescaleThreshold;  // always true on the first tile
  float alpha = 1.f;
  if (__any_sync(0xffffffffu, grow)) {
    if (grow) {
      alpha = ex2_approx((m_run - m_cand) * c);
      m_run = m_cand;
    }
    if (have_o) {
      mbar_wait(bar_o, o_parity, 44);  // S(j) was issued before PV(j-1): wait for PV(j-1) itself
      tc_fence_after();
#pragma unroll 1
      for (int c8 = 0; c8 < kOQ; c8 += 8) {
        uint32_t o[8];
        tmem_ld_x8(tO + c8, o);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st_x8(tO + c8, o);
      }
    }
    nmc = -m_run * c;
    tmem_ld_x16(tS, reinterpret_cast<uint32_t*>(s));  // S is still intact: no P stored yet
    tmem_wait_ld();
    if (masked) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i >= lim) s[i] = -INFINITY;
    }
#pragma unroll
    for (int i = 0; i < 16; i += 4) exp4(i, nmc);
  }

  // ---- first 16 keys of my quarter -> TMEM -> "early" hand-off
  {
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pk[i] = pack2<kBF16>(s[2 * i], s[2 * i + 1]);
    tmem_st_x8(tS, pk);
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(bar_early);
  }
  // ---- second 16 keys, with the row sum of the first 16 in the MUFU shadow -> "late" hand-off
  float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
  {
    uint32_t pk[8];
#pragma unroll
    for (int i = 16; i < 32; i += 4) {
      exp4(i, nmc);
      fadd2(sum0, sum1, sum0, sum1, s[i - 16], s[i - 15]);
      fadd2(sum2, sum3, sum2, sum3, s[i - 14], s[i - 13]);
      pk[(i - 16) >> 1] = pack2<kBF16>(s[i], s[i + 1]);
      pk[((i - 16) >> 1) + 1] = pack2<kBF16>(s[i + 2], s[i + 3]);
    }
    tmem_st_x8(tS + 8, pk);
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(bar_late);
  }
#pragma unroll
  for (int i = 16; i < 32; i += 4) {
    fadd2(sum0, sum1, sum0, sum1, s[i], s[i + 1]);
    fadd2(sum2, sum3, sum2, sum3, s[i + 2], s[i + 3]);
  }
  l_run = l_run * alpha + ((sum0 + sum1) + (sum2 + sum3));
}

template <int kDP, bool kBF16, bool kCausal>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kQuadThreads, 1)
fa_fwd_quad2_kernel(const __grid_constant__ CUtensorMap tmap_q,
                    const __grid_constant__ CUtensorMap tmap_k64,  // box {64 head-dim columns, 64 keys}
                    const __grid_constant__ CUtensorMap tmap_v,
                    const __grid_constant__ CUtensorMap tmap_o, const TcParams p) {
  using C = Quad2Cfg<kDP>;
  constexpr int kS = C::kStages;
  constexpr int kDBlocks = kDP / 64;
  constexpr int kKSteps = kDP / 16;
  constexpr int kOQ = kDP / 4;  // O columns each of the four threads of a row owns
  constexpr uint32_t kColO = 256u;
  constexpr int kMmaWarp = 16, kTmaWarp = 17;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t sQ = smem_u32(smem + C::kQ);
  const uint32_t sKV = smem_u32(smem + C::kKV);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::kBars + 8 * C::kNumBars);
  float* sMax = reinterpret_cast<float*>(smem + C::kMax);
  float* sFinal = reinterpret_cast<float*>(smem + C::kFinal);

  // "leader": only the copy in cluster rank 0 is used; "each": one per CTA (multicast commits)
  const uint32_t bar_q_full = smem_u32(&bars[0]);                           // leader: tx of both Q tiles
  const uint32_t bar_o = smem_u32(&bars[1]);                                // each: commit after PV(j)
  auto bar_s_full = [&](int buf) { return smem_u32(&bars[2 + buf]); };      // each
  auto bar_p_early = [&](int buf) { return smem_u32(&bars[4 + buf]); };     // leader: 32 softmax warps
  auto bar_p_late = [&](int buf) { return smem_u32(&bars[6 + buf]); };
  const uint32_t bar_o_final = smem_u32(&bars[8]);                          // each
  auto bar_kv_full = [&](int s) { return smem_u32(&bars[9 + s]); };         // leader: tx of both halves
  auto bar_kv_empty = [&](int s) { return smem_u32(&bars[9 + kS + s]); };   // each

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = kCausal ? (static_cast<int>(gridDim.x / 2) - 1 - static_cast<int>(blockIdx.x / 2))
                           : static_cast<int>(blockIdx.x / 2);
  const int qtile = 2 * pair + static_cast<int>(rank);
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int row0 = qtile * kTileM;
  int n = (p.Nkv + kTileN - 1) / kTileN;
  if (kCausal) n = min(n, 2 * pair + 2);  // lock step: both CTAs visit the tiles of the later Q tile

  auto idx_k = [](int j) { return j == 0 ? 0 : 2 * j - 1; };
  auto idx_v = [n](int j) { return (j + 1 < n) ? 2 * j + 2 : 2 * j + 1; };
#ifdef FA_TRACE
  const bool tr_cta = p.trace != nullptr && blockIdx.x == (gridDim.x / 4) * 2 && blockIdx.y == 0 && blockIdx.z == 0;
#endif

  if (warp == kMmaWarp && lane == 0) {
    mbar_init(bar_q_full, 1);
    mbar_init(bar_o, 1);
    mbar_init(bar_o_final, 1);
#pragma unroll
    for (int buf = 0; buf < 2; ++buf) {
      mbar_init(bar_s_full(buf), 1);
      mbar_init(bar_p_early(buf), 32);
      mbar_init(bar_p_late(buf), 32);
    }
#pragma unroll
    for (int s = 0; s < kS; ++s) {
      mbar_init(bar_kv_full(s), 1);
      mbar_init(bar_kv_empty(s), 1);
    }
    fence_mbar_init();
  }
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k64);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_o);
  }
  if (warp == kMmaWarp) {
    tmem_alloc_2cta(smem_u32(tmem_slot), 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (*tmem_slot != 0u) __trap();
  constexpr uint32_t tmem = 0u;
  const float c = p.scale_log2;

  if (warp == kTmaWarp) {
    // ========================================================================= TMA producer (each CTA)
    if (elect_one()) {
      const uint32_t q_full_leader = mapa_shared(bar_q_full, 0);
      if (leader) mbar_arrive_expect_tx(bar_q_full, 2 * C::kQBytes);
#pragma unroll
      for (int db = 0; db < kDBlocks; ++db)
        tma_load_4d_2cta(sQ + db * 16384, &tmap_q, q_full_leader, db * 64, row0, h, b);
      auto load = [&](bool is_v, int j, int idx) {
        const int slot = idx % kS;
        mbar_wait(bar_kv_empty(slot), ((idx / kS) & 1) ^ 1, 20);
        if (leader) mbar_arrive_expect_tx(bar_kv_full(slot), 2 * (is_v ? C::kVHalfBytes : C::kKHalfBytes));
        const uint32_t full_leader = mapa_shared(bar_kv_full(slot), 0);
        const uint32_t dst = sKV + slot * C::kSlotBytes;
        if (!is_v) {
#pragma unroll
          for (int db = 0; db < kDBlocks; ++db)
            tma_load_4d_2cta(dst + db * 8192, &tmap_k64, full_leader, db * 64, j * kTileN + rank * 64, h, b);
        } else {
          tma_load_4d_2cta(dst, &tmap_v, full_leader, rank * (kDP / 2), j * kTileN, h, b);
        }
      };
      load(false, 0, 0);
#pragma unroll 1
      for (int j = 0; j < n; ++j) {
        if (j + 1 < n) load(false, j + 1, idx_k(j + 1));
        load(true, j, idx_v(j));
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ========================================================================= MMA issuer (leader only)
    if (leader && elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(2 * kTileM, kTileN, kBF16, false, false);
      constexpr uint32_t idesc_o = make_idesc_f16(2 * kTileM, kDP, kBF16, false, true);
      constexpr uint32_t desc_hi = smem_desc_hi_sw128(1024);
      auto wait_kv = [&](int idx) {
        mbar_wait(bar_kv_full(idx % kS), (idx / kS) & 1, 30);
        tc_fence_after();
      };
      auto release_kv = [&](int idx) { tc_commit_2cta(bar_kv_empty(idx % kS), 0b11); };
      auto issue_s = [&](int j) {
        const int idx = idx_k(j);
        wait_kv(idx);
        const uint32_t k_lo = smem_desc_lo(sKV + (idx % kS) * C::kSlotBytes, 16);
        const uint32_t q_lo = smem_desc_lo(sQ, 16);
#pragma unroll
        for (int k = 0; k < kKSteps; ++k) {
          const uint32_t q_off = ((k >> 2) * 16384 + (k & 3) * 32) >> 4;
          const uint32_t k_off = ((k >> 2) * 8192 + (k & 3) * 32) >> 4;
          umma_ss2_2cta(tmem + (j & 1) * 128, q_lo + q_off, desc_hi, k_lo + k_off, desc_hi, idesc_s, k > 0);
        }
        tc_commit_2cta(bar_s_full(j & 1), 0b11);
        release_kv(idx);
      };
      auto issue_pv = [&](int j) {
        const int idx = idx_v(j);
        const int buf = j & 1;
        const uint32_t par = (j >> 1) & 1;
        wait_kv(idx);
        const uint32_t v_lo = smem_desc_lo(sKV + (idx % kS) * C::kSlotBytes, 16384);
        // k-step ks covers keys [16 ks, 16 ks + 16): quarter ks / 2, P columns at S column 32 (ks/2) + 8 (ks%2)
        auto pv_step = [&](int ks, uint32_t acc) {
          umma_ts2_2cta(tmem + kColO, tmem + buf * 128 + (ks >> 1) * 32 + (ks & 1) * 8,
                        v_lo + ((ks * 2048) >> 4), desc_hi, idesc_o, acc);
        };
        FA_QTR(2, j, 0);
        if (j > 0) mbar_wait(bar_o, (j - 1) & 1, 35);  // see fa_fwd_wide.cuh
        FA_QTR(2, j, 1);
        mbar_wait(bar_p_early(buf), par, 31);
        tc_fence_after();
        FA_QTR(2, j, 2);
        pv_step(0, j > 0);
        pv_step(2, 1);
        pv_step(4, 1);
        pv_step(6, 1);
        mbar_wait(bar_p_late(buf), par, 33);
        tc_fence_after();
        FA_QTR(2, j, 3);
        pv_step(1, 1);
        pv_step(3, 1);
        pv_step(5, 1);
        pv_step(7, 1);
        tc_commit_2cta(bar_o, 0b11);
        release_kv(idx);
        if (j == n - 1) tc_commit_2cta(bar_o_final, 0b11);
        FA_QTR(2, j, 4);
      };

      mbar_wait(bar_q_full, 0, 34);
      tc_fence_after();
      issue_s(0);
      if (n > 1) issue_s(1);
#pragma unroll 1
      for (int j = 0; j < n; ++j) {
        issue_pv(j);
        if (j + 2 < n) issue_s(j + 2);
        FA_QTR(2, j, 5);
      }
    }
    __syncwarp();
  } else {
    // ========================================================================= softmax warps 0-15 (each CTA)
    const int quarter = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tO = tmem + lane_base + kColO + quarter * kOQ;
    const int group_bar = 1 + (warp & 3);  // named barrier of the four warps sharing my rows
    float* maxes = sMax + r;               // + parity * 512 + quarter * 128
    const uint32_t p_early0 = mapa_shared(bar_p_early(0), 0);
    const uint32_t p_late0 = mapa_shared(bar_p_late(0), 0);

    float m_run = -INFINITY;
    float l_run = 0.f;  // partial row sum over my key quarter

#pragma unroll 1
    for (int j = 0; j < n; ++j) {
      const int buf = j & 1;
      const uint32_t tS = tmem + lane_base + buf * 128 + quarter * 32;
#ifdef FA_TRACE
      const int tr_role = (warp == 0) ? 0 : ((warp == 5) ? 1 : -1);
      if (tr_role >= 0) FA_QTR(tr_role, j, 0);
#endif
      mbar_wait_warp(bar_s_full(buf), (j >> 1) & 1, 40);
      tc_fence_after();
#ifdef FA_TRACE
      if (tr_role >= 0) FA_QTR(tr_role, j, 1);
#endif
      float s[32];
      tmem_ld_x32(tS, reinterpret_cast<uint32_t*>(s));
      tmem_wait_ld();
#ifdef FA_TRACE
      if (tr_role >= 0) FA_QTR(tr_role, j, 2);
#endif
      // causal: in tile j >= qtile the keys [128 j + i] with i > r - 128 (j - qtile) are hidden
      const int lim_c = r - (j - qtile) * kTileN + 1 - quarter * 32;
      quad_softmax_step<kDP, kBF16>(s, tS, tO, quarter, lane, j * kTileN + quarter * 32, p.Nkv,
                                    kCausal && j >= qtile, lim_c, c, m_run, l_run, j > 0, maxes + buf * 512,
                                    group_bar, p_early0 + buf * 8, p_late0 + buf * 8, bar_o,
                                    static_cast<uint32_t>((j - 1) & 1));
#ifdef FA_TRACE
      if (tr_role >= 0) FA_QTR(tr_role, j, 3);
#endif
    }

    // ---- epilogue: O / l -> 16 bit -> swizzled smem (my Q buffer) -> TMA store
    sFinal[quarter * 128 + r] = l_run;
    named_bar_sync(group_bar, 128);
    const float l_tot = (sFinal[r] + sFinal[128 + r]) + (sFinal[256 + r] + sFinal[384 + r]);
    const int row = row0 + r;
    if (quarter == 0 && p.lse != nullptr && row < p.Nq)
      p.lse[(static_cast<int64_t>(b) * p.H + h) * p.Nq + row] = m_run * c + log2f(l_tot);
    const float inv_l = 1.f / l_tot;
    mbar_wait(bar_o_final, 0, 54);
    tc_fence_after();
    uint8_t* stage = smem + C::kQ;
    {
      uint32_t o[kOQ];
      if constexpr (kOQ == 32) {
        tmem_ld_x32(tO, o);
      } else {
        tmem_ld_x16(tO, o);
      }
      tmem_wait_ld();
#pragma unroll
      for (int ch = 0; ch < kOQ / 8; ++ch) {
        uint4 val;
        val.x = pack2<kBF16>(__uint_as_float(o[ch * 8 + 0]) * inv_l, __uint_as_float(o[ch * 8 + 1]) * inv_l);
        val.y = pack2<kBF16>(__uint_as_float(o[ch * 8 + 2]) * inv_l, __uint_as_float(o[ch * 8 + 3]) * inv_l);
        val.z = pack2<kBF16>(__uint_as_float(o[ch * 8 + 4]) * inv_l, __uint_as_float(o[ch * 8 + 5]) * inv_l);
        val.w = pack2<kBF16>(__uint_as_float(o[ch * 8 + 6]) * inv_l, __uint_as_float(o[ch * 8 + 7]) * inv_l);
        *reinterpret_cast<uint4*>(stage + sw128_offset_16bit(r, quarter * kOQ + ch * 8)) = val;
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(5, 512);
    if (warp == 0 && lane == 0) {
#pragma unroll
      for (int db = 0; db < kDBlocks; ++db)
        tma_store_4d(&tmap_o, sQ + db * 16384, db * 64, row0, h, b);
      tma_store_commit();
      tma_store_wait_read();
    }
    __syncwarp();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == kMmaWarp) tmem_dealloc_2cta(tmem, 512);
}

}  // namespace fa
