// The two-tile warp-specialised forward kernel of fa_fwd_ws.cuh on CTA PAIRS ("ws2"): a thread-block
// cluster of two CTAs (two SMs) shares every K/V tile through tcgen05 cta_group::2.  Non-causal,
// head dims <= 128.
//
// Each CTA keeps the arrangement of fa_fwd_ws.cuh - two 128-row Q tiles that ping-pong on the tensor
// cores, 16 softmax warps (two threads per row), one MMA warp, one TMA warp, the same TMEM map - but one
// tcgen05.mma now computes tile t of BOTH CTAs (M = 256): each CTA supplies its own 128 rows of A (Q from its
// shared memory, P from its tensor memory) and HALF of B:
//   S_t = Q_t K^T : CTA r holds keys [64r, 64r+64) of the K tile
//   O_t += P_t V  : CTA r holds head-dim columns [D/2 r, D/2 r + D/2) of the V tile
// so an SM fetches half of every K/V tile (the ring holds 8 half tiles in the space of 4 whole ones) and
// reads half of the B operands.  The leader CTA (cluster rank 0) issues every MMA; the K/V "full", Q "full"
// and P hand-off barriers live in the leader (both CTAs' TMA loads and softmax warps signal them), and
// "S ready" / "slot free" / "O final" reach both CTAs through multicast tcgen05.commit.  The operand split is
// pinned on the hardware by umma2_probe.cuh; fa_fwd_wide2.cuh is the one-tile sibling of this kernel.
//
// Motivation: with the K/V TMA loads of fa_fwd_ws.cuh switched off the step gets 5 %
// shorter, and the one-tile kernel gains 5 % from pairing at D = 128 even where its softmax is the limit.
#pragma once
#include "fa_fwd_ws.cuh"

namespace fa {

template <int kDP>
struct Ws2Cfg {
  static_assert(kDP == 64 || kDP == 128, "ws2 kernel: padded head dim 64 or 128");
  static constexpr int kTileBytes = kTileM * kDP * 2;          // one Q tile
  static constexpr int kKHalfBytes = (kTileN / 2) * kDP * 2;   // 64 keys x kDP: kDP/64 blocks of 8 KB
  static constexpr int kVHalfBytes = kTileN * 64 * 2;          // 128 keys x kDP/2 columns in one 64-column block
                                                               // (half used at kDP = 64)
  static constexpr int kSlotBytes = 16384;
  static constexpr int kStages = 8;
  static constexpr int kQ = 0;                                 // 2 Q tiles (re-used as O staging)
  static constexpr int kKV = kQ + 2 * kTileBytes;
  static constexpr int kBars = kKV + kStages * kSlotBytes;
  static constexpr int kNumBars = 12 + 2 * kStages;
  static constexpr int kMax = kBars + 8 * kNumBars + 16;       // float [2 parity][2 tile][2 half][128]
  static constexpr int kFinal = kMax + 2 * 2 * 2 * 128 * 4;    // float [2 tile][2 half][128] row sums
  static constexpr int kTotal = kFinal + 2 * 2 * 128 * 4 + 1024;  // + alignment slack
  static_assert(kKHalfBytes <= kSlotBytes && kVHalfBytes <= kSlotBytes && kTotal <= 232448, "shared memory budget");
};

template <int kDP, bool kBF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWsThreads, 1)
fa_fwd_ws2_kernel(const __grid_constant__ CUtensorMap tmap_q,
                  const __grid_constant__ CUtensorMap tmap_k64,  // box {64 head-dim columns, 64 keys}
                  const __grid_constant__ CUtensorMap tmap_v,
                  const __grid_constant__ CUtensorMap tmap_o, const TcParams p) {
  using C = Ws2Cfg<kDP>;
  constexpr int kS = C::kStages;
  constexpr int kDBlocks = kDP / 64;
  constexpr int kKSteps = kDP / 16;
  constexpr int kOHalf = kDP / 2;
  auto col_s = [](int t) -> uint32_t { return static_cast<uint32_t>(t) * 128u; };
  auto col_o = [](int t) -> uint32_t { return 256u + static_cast<uint32_t>(t) * 128u; };

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t sQ = smem_u32(smem + C::kQ);
  const uint32_t sKV = smem_u32(smem + C::kKV);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::kBars + 8 * C::kNumBars);
  float* sMax = reinterpret_cast<float*>(smem + C::kMax);
  float* sFinal = reinterpret_cast<float*>(smem + C::kFinal);

  // "leader": only the copy in cluster rank 0 is used; "each": one per CTA, signalled by multicast commits
  auto bar_q_full = [&](int t) { return smem_u32(&bars[t]); };              // leader: tx of both CTAs' Q_t
  auto bar_s_full = [&](int t) { return smem_u32(&bars[2 + t]); };          // each
  auto bar_p_early = [&](int t) { return smem_u32(&bars[4 + t]); };         // leader: 16 softmax warps
  auto bar_p_mid = [&](int t) { return smem_u32(&bars[6 + t]); };
  auto bar_p_late = [&](int t) { return smem_u32(&bars[8 + t]); };
  auto bar_o_final = [&](int t) { return smem_u32(&bars[10 + t]); };        // each
  auto bar_kv_full = [&](int s) { return smem_u32(&bars[12 + s]); };        // leader: tx of both halves
  auto bar_kv_empty = [&](int s) { return smem_u32(&bars[12 + kS + s]); };  // each

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int blk = blockIdx.x;  // 256-row query block; the pair is blocks (2p, 2p+1), grid padded to even
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int row0 = blk * 2 * kTileM;
  const int n = (p.Nkv + kTileN - 1) / kTileN;  // KV tiles: the same for all four Q tiles of the pair

  if (warp == 16 && lane == 0) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_q_full(t), 1);
      mbar_init(bar_s_full(t), 1);
      mbar_init(bar_p_early(t), 16);
      mbar_init(bar_p_mid(t), 16);
      mbar_init(bar_p_late(t), 16);
      mbar_init(bar_o_final(t), 1);
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
    tma_prefetch_desc(&tmap_k64);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_o);
  }
  if (warp == 16) {
    tmem_alloc_2cta(smem_u32(tmem_slot), 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  // PDL: everything above overlapped the previous kernel's tail; global memory is touched only below
  pdl_wait();
  pdl_launch_dependents();
  if (*tmem_slot != 0u) __trap();
  constexpr uint32_t tmem = 0u;
  const float c = p.scale_log2;

  if (warp >= 16) {
    // =========================================================================================
    // warpgroup 4: MMA issuer (warp 16, leader CTA only), TMA producer (warp 17, each CTA)
    // =========================================================================================
    setmaxnreg_dec<56>();  // 512 x 104 + 128 x 56 <= 640 x 96: the issuing thread keeps its descriptors in registers
    if (warp == 17) {
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (leader) mbar_arrive_expect_tx(bar_q_full(t), 2 * C::kTileBytes);
          const uint32_t q_full_leader = mapa_shared(bar_q_full(t), 0);
#pragma unroll
          for (int db = 0; db < kDBlocks; ++db)
            tma_load_4d_2cta(sQ + t * C::kTileBytes + db * 16384, &tmap_q, q_full_leader, db * 64,
                             row0 + t * kTileM, h, b);
        }
#pragma unroll 1
        for (int idx = 0; idx < 2 * n; ++idx) {  // ring order K0 V0 K1 V1 ...
          const int slot = idx % kS;
          const int j = idx >> 1;
          mbar_wait(bar_kv_empty(slot), ((idx / kS) & 1) ^ 1, 20);
          const uint32_t full_leader = mapa_shared(bar_kv_full(slot), 0);
          const uint32_t dst = sKV + slot * C::kSlotBytes;
          if ((idx & 1) == 0) {  // my 64 keys of K_j: kDP/64 [64 keys x 64 columns] blocks, 8 KB apart
            if (leader) mbar_arrive_expect_tx(bar_kv_full(slot), 2 * C::kKHalfBytes);
#pragma unroll
            for (int db = 0; db < kDBlocks; ++db)
              tma_load_4d_2cta(dst + db * 8192, &tmap_k64, full_leader, db * 64, j * kTileN + rank * 64, h, b);
          } else {               // my kDP/2 head-dim columns of V_j: one [128 keys x 64 columns] block
            if (leader) mbar_arrive_expect_tx(bar_kv_full(slot), 2 * C::kVHalfBytes);
            tma_load_4d_2cta(dst, &tmap_v, full_leader, rank * (kDP / 2), j * kTileN, h, b);
          }
        }
      }
      __syncwarp();
    } else if (warp == 16) {
      if (leader && elect_one()) {
        constexpr uint32_t idesc_s = make_idesc_f16(2 * kTileM, kTileN, kBF16, false, false);
        constexpr uint32_t idesc_o = make_idesc_f16(2 * kTileM, kDP, kBF16, false, true);
        constexpr uint32_t desc_hi = smem_desc_hi_sw128(1024);
        auto wait_kv = [&](int idx) {
          mbar_wait(bar_kv_full(idx % kS), (idx / kS) & 1, 30);
          tc_fence_after();
        };
        auto release_kv = [&](int idx) { tc_commit_2cta(bar_kv_empty(idx % kS), 0b11); };
        auto issue_s = [&](int t, int j) {  // S_t = Q_t K_j^T for both CTAs
          const uint32_t k_lo = smem_desc_lo(sKV + ((2 * j) % kS) * C::kSlotBytes, 16);
          const uint32_t q_lo = smem_desc_lo(sQ + t * C::kTileBytes, 16);
#pragma unroll
          for (int k = 0; k < kKSteps; ++k) {
            const uint32_t q_off = ((k >> 2) * 16384 + (k & 3) * 32) >> 4;
            const uint32_t k_off = ((k >> 2) * 8192 + (k & 3) * 32) >> 4;
            umma_ss2_2cta(tmem + col_s(t), q_lo + q_off, desc_hi, k_lo + k_off, desc_hi, idesc_s, k > 0);
          }
          tc_commit_2cta(bar_s_full(t), 0b11);
        };
        auto pv_step = [&](int t, uint32_t v_lo, int ks, uint32_t acc) {
          umma_ts2_2cta(tmem + col_o(t), tmem + col_s(t) + (ks >> 2) * 64 + (ks & 3) * 8,
                        v_lo + ((ks * 2048) >> 4), desc_hi, idesc_o, acc);
        };
        auto issue_pv = [&](int t, int j) {  // O_t += P_t V_j for both CTAs
          const uint32_t v_lo = smem_desc_lo(sKV + ((2 * j + 1) % kS) * C::kSlotBytes, 16384);
          mbar_wait(bar_p_early(t), j & 1, 31 + t);
          tc_fence_after();
          pv_step(t, v_lo, 0, j > 0);
          pv_step(t, v_lo, 1, 1);
          pv_step(t, v_lo, 4, 1);
          pv_step(t, v_lo, 5, 1);
          mbar_wait(bar_p_mid(t), j & 1, 37 + t);
          tc_fence_after();
          pv_step(t, v_lo, 2, 1);
          pv_step(t, v_lo, 6, 1);
          mbar_wait(bar_p_late(t), j & 1, 35 + t);
          tc_fence_after();
          pv_step(t, v_lo, 3, 1);
          pv_step(t, v_lo, 7, 1);
          if (j == n - 1) tc_commit_2cta(bar_o_final(t), 0b11);
        };

        wait_kv(0);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(bar_q_full(t), 0, 33);
          tc_fence_after();
          issue_s(t, 0);
        }
        release_kv(0);
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
          const int nx = j + 1;
          wait_kv(2 * j + 1);
          issue_pv(0, j);
          if (nx < n) {
            wait_kv(2 * nx);
            issue_s(0, nx);
          }
          issue_pv(1, j);
          release_kv(2 * j + 1);
          if (nx < n) {
            issue_s(1, nx);
            release_kv(2 * nx);
          }
        }
      }
      __syncwarp();
    }
  } else {
    // =========================================================================================
    // softmax warps (0-7: tile 0, 8-15: tile 1), each CTA
    // =========================================================================================
    setmaxnreg_inc<104>();
    const int t = warp >> 3;
    const int half = (warp >> 2) & 1;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_base + col_s(t) + half * 64;
    const uint32_t tO = tmem + lane_base + col_o(t) + half * kOHalf;
    const int pair_bar = 1 + t * 4 + (warp & 3);
    const int tile_row0 = row0 + t * kTileM;
    float* my_max = sMax + (t * 2 + half) * 128 + r;
    const float* other_max = sMax + (t * 2 + (half ^ 1)) * 128 + r;
    const uint32_t p_early = mapa_shared(bar_p_early(t), 0);  // the pair's hand-off barriers: in the leader
    const uint32_t p_mid = mapa_shared(bar_p_mid(t), 0);
    const uint32_t p_late = mapa_shared(bar_p_late(t), 0);

    float m_run = -INFINITY;
    float l_run = 0.f;

#pragma unroll 1
    for (int j = 0; j < n; ++j) {
      mbar_wait_warp(bar_s_full(t), j & 1, 40 + t);
      tc_fence_after();
      float s[64];
      tmem_ld_x32(tS, reinterpret_cast<uint32_t*>(s));
      tmem_ld_x32(tS + 32, reinterpret_cast<uint32_t*>(s) + 32);
      tmem_wait_ld();
      ws_softmax_step<kDP, kBF16, true>(s, tS, tO, half, r, lane, j * kTileN + half * 64, p.Nkv, false, c, m_run,
                                        l_run, j > 0, my_max + (j & 1) * 512, other_max + (j & 1) * 512, pair_bar,
                                        p_early, p_late, 0u, p_mid);
    }

    // ---- epilogue: O / l -> 16 bit -> swizzled smem (the tile's Q buffer) -> TMA store
    sFinal[(t * 2 + half) * 128 + r] = l_run;
    named_bar_sync(pair_bar, 64);
    const float l_tot = l_run + sFinal[(t * 2 + (half ^ 1)) * 128 + r];
    const int row = tile_row0 + r;
    if (half == 0 && p.lse != nullptr && row < p.Nq)
      p.lse[(static_cast<int64_t>(b) * p.H + h) * p.Nq + row] = m_run * c + log2f(l_tot);
    const float inv_l = 1.f / l_tot;
    mbar_wait(bar_o_final(t), 0, 54 + t);  // every MMA that touches tile t (of both CTAs) is done
    tc_fence_after();
    uint8_t* stage = smem + C::kQ + t * C::kTileBytes;
    o_row_half_to_stage<kOHalf, kBF16, true>(tO, stage, r, half, inv_l);
    fence_proxy_async_smem();
    named_bar_sync(9 + t, 256);
    if ((warp & 7) == 0 && lane == 0) {
#pragma unroll
      for (int db = 0; db < kDBlocks; ++db)
        tma_store_4d(&tmap_o, sQ + t * C::kTileBytes + db * 16384, db * 64, row0 + t * kTileM, h, b);
      tma_store_commit();
      tma_store_wait_read();
    }
    __syncwarp();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 16) tmem_dealloc_2cta(tmem, 512);
}

}  // namespace fa
