// Wide-head tcgen05 forward kernel ("wide"): head dims 129..256 (padded to 192 or 256 columns).
//
// At D > 128 two Q tiles no longer fit in tensor memory (2 x (128 S + D O) columns > 512), so the
// two-tile ping-pong of fa_fwd_ws.cuh is replaced by ONE 128-row Q tile per CTA with the SCORE tile
// double-buffered instead: while the softmax warps work on S(j) in registers the tensor cores compute
// S(j+1) into the other buffer and O += P(j-1) V(j-1).  With D = 256 one KV tile costs the tensor
// cores 2 x 1024 cycles against ~1024 MUFU cycles of softmax, so a single softmax group keeps up.
//
//   warps 0-7  softmax: warp w owns query rows 32*(w%4).. and the 64-key half w/4 of every S tile
//              (the same two-threads-per-row arrangement, stale-max speculation, lazy rescale and
//              three-part P hand-off as the ws kernel: ws_softmax_step is shared)
//   warp  8    MMA issuer (one elected thread)
//   warp  9    TMA producer (Q once, then the K/V ring in consumption order K0 K1 V0 K2 V1 K3 ...)
//
// TMEM: S buffer 0 [0,128)  S buffer 1 [128,256)  O [256,256+D).  P(j) overwrites columns
// [64h, 64h+32) of S buffer j%2, so S(j+2) is issued after O += P(j) V(j).
// Tensor-core issue order:  S(0) S(1) | PV(0) S(2) | PV(1) S(3) | ...
// Because S(j) is issued BEFORE PV(j-1), "S(j) is ready" does not imply "O holds PV(j-1)"; the rare
// O rescale therefore waits on the barrier PV(j-1) commits to (bar_o).
//
// Shared memory: Q 128 x D x 2 B, K/V ring of whole tiles (2 slots at D=256: 192 KB in all; 3 at 192).
//
// Replaces /root/reference/rocwmma_fattn/kernel_fp16.cu:306-544 for the head dims the reference
// reaches through its host-side padding (kernel_fp16.cu:763-779; sweep in bench_with_sdpa.py:259-261).
#pragma once
#include "fa_fwd_ws.cuh"

namespace fa {

constexpr int kWideThreads = 320;

template <int kDP>
struct WideCfg {
  static_assert(kDP == 64 || kDP == 128 || kDP == 192 || kDP == 256, "wide kernel: padded head dim");
  static constexpr int kTileBytes = kTileM * kDP * 2;
  // K/V ring slots (one K or one V tile each): what fits next to the Q tile
  static constexpr int kStages = (kDP == 256) ? 2 : (kDP == 192) ? 3 : (kDP == 128) ? 5 : 8;
  static constexpr int kQ = 0;                           // Q tile (re-used as O staging)
  static constexpr int kKV = kQ + kTileBytes;
  static constexpr int kBars = kKV + kStages * kTileBytes;
  static constexpr int kNumBars = 11 + 2 * kStages;
  static constexpr int kMax = kBars + 8 * kNumBars + 16;   // float [2 parity][2 half][128]
  static constexpr int kFinal = kMax + 2 * 2 * 128 * 4;    // float [2 half][128] row sums
  static constexpr int kTotal = kFinal + 2 * 128 * 4 + 1024;  // + alignment slack
};

template <int kDP, bool kBF16, bool kCausal>
__global__ void __launch_bounds__(kWideThreads, 1)
fa_fwd_wide_kernel(const __grid_constant__ CUtensorMap tmap_q,
                   const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_v,
                   const __grid_constant__ CUtensorMap tmap_o, const TcParams p) {
  using C = WideCfg<kDP>;
  constexpr int kS = C::kStages;
  constexpr int kDBlocks = kDP / 64;
  constexpr int kKSteps = kDP / 16;
  constexpr int kOHalf = kDP / 2;  // O columns each of the two threads of a row owns
  constexpr uint32_t kColO = 256u;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t sQ = smem_u32(smem + C::kQ);
  const uint32_t sKV = smem_u32(smem + C::kKV);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::kBars + 8 * C::kNumBars);
  float* sMax = reinterpret_cast<float*>(smem + C::kMax);
  float* sFinal = reinterpret_cast<float*>(smem + C::kFinal);

  const uint32_t bar_q_full = smem_u32(&bars[0]);                           // tx, count 1
  const uint32_t bar_o = smem_u32(&bars[1]);                                // tcgen05.commit after PV(j)
  auto bar_s_full = [&](int buf) { return smem_u32(&bars[2 + buf]); };      // tcgen05.commit
  auto bar_p_early = [&](int buf) { return smem_u32(&bars[4 + buf]); };     // 8 softmax warps
  auto bar_p_mid = [&](int buf) { return smem_u32(&bars[6 + buf]); };
  auto bar_p_late = [&](int buf) { return smem_u32(&bars[8 + buf]); };
  const uint32_t bar_o_final = smem_u32(&bars[10]);                         // tcgen05.commit after the last PV
  auto bar_kv_full = [&](int s) { return smem_u32(&bars[11 + s]); };        // tx, count 1
  auto bar_kv_empty = [&](int s) { return smem_u32(&bars[11 + kS + s]); };  // tcgen05.commit

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  int qtile, h, b;  // causal: longest tiles first across the whole launch (work_coords)
  work_coords<kCausal>((p.Nq + kTileM - 1) / kTileM, p.H, 1, qtile, h, b);
  const int row0 = qtile * kTileM;

  int n = (p.Nkv + kTileN - 1) / kTileN;  // KV tiles this Q tile visits (>= 1)
  if (kCausal) n = min(n, qtile + 1);

  // position of a tile in the ring (= consumption) order K0 K1 V0 K2 V1 ... K(n-1) V(n-2) V(n-1)
  auto idx_k = [](int j) { return j == 0 ? 0 : 2 * j - 1; };
  auto idx_v = [n](int j) { return (j + 1 < n) ? 2 * j + 2 : 2 * j + 1; };

  if (warp == 8 && lane == 0) {
    mbar_init(bar_q_full, 1);
    mbar_init(bar_o, 1);
    mbar_init(bar_o_final, 1);
#pragma unroll
    for (int buf = 0; buf < 2; ++buf) {
      mbar_init(bar_s_full(buf), 1);
      mbar_init(bar_p_early(buf), 8);
      mbar_init(bar_p_mid(buf), 8);
      mbar_init(bar_p_late(buf), 8);
    }
#pragma unroll
    for (int s = 0; s < kS; ++s) {
      mbar_init(bar_kv_full(s), 1);
      mbar_init(bar_kv_empty(s), 1);
    }
    fence_mbar_init();
  }
  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_o);
    // Q, K0, K1, V0 -> L2 before pdl_wait() (see fa_fwd_ws.cuh)
#pragma unroll
    for (int db = 0; db < kDBlocks; ++db) {
      tma_prefetch_l2_4d(&tmap_q, db * 64, row0, h, b);
      tma_prefetch_l2_4d(&tmap_k, db * 64, 0, h, b);
      tma_prefetch_l2_4d(&tmap_v, db * 64, 0, h, b);
      if (n > 1) tma_prefetch_l2_4d(&tmap_k, db * 64, kTileN, h, b);
    }
  }
  if (warp == 8) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // PDL: everything above overlapped the previous kernel's tail; global memory is touched only below
  pdl_wait();
  pdl_launch_dependents();
  if (*tmem_slot != 0u) __trap();  // one CTA per SM owns all of TMEM: constant addresses (see ws kernel)
  constexpr uint32_t tmem = 0u;
  const float c = p.scale_log2;

  if (warp == 9) {
    // ========================================================================= TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_q_full, C::kTileBytes);
#pragma unroll
      for (int db = 0; db < kDBlocks; ++db)
        tma_load_4d(sQ + db * 16384, &tmap_q, bar_q_full, db * 64, row0, h, b);
      auto load = [&](const CUtensorMap* map, int j, int idx) {
        const int slot = idx % kS;
        mbar_wait(bar_kv_empty(slot), ((idx / kS) & 1) ^ 1, 20);
        mbar_arrive_expect_tx(bar_kv_full(slot), C::kTileBytes);
#pragma unroll
        for (int db = 0; db < kDBlocks; ++db)
          tma_load_4d(sKV + slot * C::kTileBytes + db * 16384, map, bar_kv_full(slot), db * 64,
                      j * kTileN, h, b);
      };
      load(&tmap_k, 0, 0);
#pragma unroll 1
      for (int j = 0; j < n; ++j) {
        if (j + 1 < n) load(&tmap_k, j + 1, idx_k(j + 1));
        load(&tmap_v, j, idx_v(j));
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ========================================================================= MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(kTileM, kTileN, kBF16, false, false);
      constexpr uint32_t idesc_o = make_idesc_f16(kTileM, kDP, kBF16, false, true);
      constexpr uint32_t desc_hi = smem_desc_hi_sw128(1024);
      auto wait_kv = [&](int idx) {
        mbar_wait(bar_kv_full(idx % kS), (idx / kS) & 1, 30);
        tc_fence_after();
      };
      auto release_kv = [&](int idx) { tc_commit(bar_kv_empty(idx % kS)); };
      auto issue_s = [&](int j) {  // S(j) = Q K_j^T into buffer j % 2
        const int idx = idx_k(j);
        wait_kv(idx);
        const uint32_t k_lo = smem_desc_lo(sKV + (idx % kS) * C::kTileBytes, 16);
        const uint32_t q_lo = smem_desc_lo(sQ, 16);
#pragma unroll
        for (int k = 0; k < kKSteps; ++k) {
          const uint32_t off = ((k >> 2) * 16384 + (k & 3) * 32) >> 4;
          umma_ss2(tmem + (j & 1) * 128, q_lo + off, desc_hi, k_lo + off, desc_hi, idesc_s, k > 0);
        }
        tc_commit(bar_s_full(j & 1));
        release_kv(idx);
      };
      auto issue_pv = [&](int j) {  // O += P(j) V_j, P in S buffer j % 2
        const int idx = idx_v(j);
        const int buf = j & 1;
        const uint32_t par = (j >> 1) & 1;
        wait_kv(idx);
        const uint32_t v_lo = smem_desc_lo(sKV + (idx % kS) * C::kTileBytes, 16384);
        // k-step ks covers keys [16 ks, 16 ks + 16): P columns of half ks/4 at S column
        // 64 (ks/4) + 8 (ks%4); V rows 16 ks of the tile (2048 bytes apart in the MN-major tile)
        auto pv_step = [&](int ks, uint32_t acc) {
          umma_ts2(tmem + kColO, tmem + buf * 128 + (ks >> 2) * 64 + (ks & 3) * 8,
                   v_lo + ((ks * 2048) >> 4), desc_hi, idesc_o, acc);
        };
        // Observe PV(j-1)'s phase of bar_o before arming the next one: keeps the barrier at most one
        // phase ahead of its (rare) softmax waiters.  Free in the tensor-bound regime: S(j+1) is still
        // queued behind PV(j-1) when this returns.
        if (j > 0) mbar_wait(bar_o, (j - 1) & 1, 35);
        mbar_wait(bar_p_early(buf), par, 31);
        tc_fence_after();
        pv_step(0, j > 0);
        pv_step(1, 1);
        pv_step(4, 1);
        pv_step(5, 1);
        if (kPvParts == 3) {
          mbar_wait(bar_p_mid(buf), par, 32);
          tc_fence_after();
          pv_step(2, 1);
          pv_step(6, 1);
          mbar_wait(bar_p_late(buf), par, 33);
          tc_fence_after();
          pv_step(3, 1);
          pv_step(7, 1);
        } else {
          mbar_wait(bar_p_late(buf), par, 33);
          tc_fence_after();
          pv_step(2, 1);
          pv_step(3, 1);
          pv_step(6, 1);
          pv_step(7, 1);
        }
        tc_commit(bar_o);
        release_kv(idx);
        if (j == n - 1) tc_commit(bar_o_final);
      };

      mbar_wait(bar_q_full, 0, 34);
      tc_fence_after();
      issue_s(0);
      if (n > 1) issue_s(1);
#pragma unroll 1
      for (int j = 0; j < n; ++j) {
        issue_pv(j);
        if (j + 2 < n) issue_s(j + 2);
      }
    }
    __syncwarp();
  } else {
    // ========================================================================= softmax warps 0-7
    const int half = (warp >> 2) & 1;
    const int r = (warp & 3) * 32 + lane;  // query row inside the tile = TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tO = tmem + lane_base + kColO + half * kOHalf;
    const int pair_bar = 1 + (warp & 3);  // named barrier of the two warps sharing my rows
    float* my_max = sMax + half * 128 + r;
    const float* other_max = sMax + (half ^ 1) * 128 + r;

    float m_run = -INFINITY;
    float l_run = 0.f;  // partial row sum over my key half

    auto kv_step = [&](int j, auto first_tag, auto nomask_tag) {
      constexpr bool kFirstStep = decltype(first_tag)::value;
      constexpr bool kNoMaskStep = decltype(nomask_tag)::value;
      const int buf = j & 1;
      const uint32_t tS = tmem + lane_base + buf * 128 + half * 64;  // my 64 S columns; P over [0,32)
      mbar_wait_warp(bar_s_full(buf), (j >> 1) & 1, 40);
      tc_fence_after();
      float s[64];
      tmem_ld_x32(tS, reinterpret_cast<uint32_t*>(s));
      tmem_ld_x32(tS + 32, reinterpret_cast<uint32_t*>(s) + 32);
      tmem_wait_ld();
      ws_softmax_step<kDP, kBF16, false, kFirstStep, kNoMaskStep>(s, tS, tO, half, r, lane, j * kTileN + half * 64, p.Nkv,
                                  kCausal && (j == qtile), c, m_run, l_run, FA_PEEL_FIRST ? !kFirstStep : (j > 0),
                                  my_max + buf * 256, other_max + buf * 256, pair_bar,
                                  bar_p_early(buf), bar_p_late(buf), 0u, bar_p_mid(buf), bar_o,
                                  static_cast<uint32_t>((j - 1) & 1));
    };
    // (the causal diagonal and the ragged tail can only be the last tile: n = min(all tiles, qtile + 1))
#if FA_PEEL_FIRST && FA_PEEL_MASK
    kv_step(0, std::true_type{}, std::false_type{});  // (n >= 1)
#pragma unroll 1
    for (int j = 1; j < n - 1; ++j) kv_step(j, std::false_type{}, std::true_type{});
    if (n > 1) kv_step(n - 1, std::false_type{}, std::false_type{});
#elif FA_PEEL_FIRST
    kv_step(0, std::true_type{}, std::false_type{});  // (n >= 1)
#pragma unroll 1
    for (int j = 1; j < n; ++j) kv_step(j, std::false_type{}, std::false_type{});
#else
#pragma unroll 1
    for (int j = 0; j < n; ++j) kv_step(j, std::false_type{}, std::false_type{});
#endif

    // ---- epilogue: O / l -> 16 bit -> swizzled smem (the Q buffer) -> TMA store
    sFinal[half * 128 + r] = l_run;
    named_bar_sync(pair_bar, 64);
    const float l_tot = l_run + sFinal[(half ^ 1) * 128 + r];
    const int row = row0 + r;
    if (half == 0 && p.lse != nullptr && row < p.Nq)
      p.lse[(static_cast<int64_t>(b) * p.H + h) * p.Nq + row] = m_run * c + log2f(l_tot);
    const float inv_l = 1.f / l_tot;
    // The last PV - and so every MMA - is done: Q is free too.  (Not bar_o: softmax(n-1) may finish
    // before PV(n-2) has, and a parity wait cannot tell phase n-2 from phase n.)
    mbar_wait(bar_o_final, 0, 54);
    tc_fence_after();
    uint8_t* stage = smem + C::kQ;
    o_row_half_to_stage<kOHalf, kBF16, false>(tO, stage, r, half, inv_l);
    fence_proxy_async_smem();
    named_bar_sync(5, 256);
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
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}

}  // namespace fa
