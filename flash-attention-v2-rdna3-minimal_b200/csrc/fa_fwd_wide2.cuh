// CTA-pair wide-head forward kernel ("wide2"): head dims 129..256, on a thread-block
// cluster of TWO CTAs (two SMs) that share every K/V tile through tcgen05 cta_group::2.
//
// fa_fwd_wide.cuh at D = 256 is bound by shared-memory bandwidth and by its two-slot K/V ring: per KV
// tile an SM reads 128 KB of MMA operands and receives 128 KB of K/V by TMA (~2500 cycles of a 128 B/clk
// port against 2048 tensor cycles).  Here the pair computes one M = 256 product per MMA - each CTA owns
// its 128 query rows (A operand, accumulators and P in its own tensor memory) and HALF of the B operand:
//   S = Q K^T   : CTA r holds keys [64r, 64r+64) of the K tile   (half the rows of a K-major B)
//   O += P V    : CTA r holds head-dim columns [D/2 r, D/2 r + D/2) of the V tile   (half the columns)
// so each SM fetches and stores half of every K/V tile (64 KB instead of 128 KB per KV tile), reads
// half of the B operand, and the ring holds four half-tiles instead of two whole ones.
// umma2_probe.cuh pins the operand split on the hardware (tests: test_umma_cta_pair_selftest).
//
// Per CTA the roles are those of fa_fwd_wide.cuh (8 softmax warps, one MMA warp, one TMA warp), with:
//   - the LEADER CTA (cluster rank 0) issuing every MMA for the pair; the peer's MMA warp idles
//   - K/V "full" barriers and the P hand-off barriers living in the leader: both CTAs' TMA loads count
//     their bytes there (cp.async.bulk.tensor ... cta_group::2) and both CTAs' softmax warps arrive there
//     (mbarrier.arrive.shared::cluster), 16 warps per phase
//   - "S ready", "K/V slot free", "PV done" signalled to both CTAs by one multicast tcgen05.commit
// The two CTAs advance in lock step: under a causal mask both visit the KV tiles of the later Q tile and
// the earlier one sees its last tile fully masked.
//
// Replaces /root/reference/rocwmma_fattn/kernel_fp16.cu:306-544 for padded head dims 192 and 256.
#pragma once
#include "fa_fwd_wide.cuh"

namespace fa {

template <int kDP_>
struct Wide2Cfg {
  static_assert(kDP_ == 64 || kDP_ == 128 || kDP_ == 192 || kDP_ == 256, "pair kernel: padded head dim");
  static constexpr int kDP = kDP_;
  static constexpr int kQBytes = kTileM * kDP * 2;             // my 128 query rows (64 KB at 256)
  static constexpr int kKHalfBytes = (kTileN / 2) * kDP * 2;   // 64 keys x kDP: kDP/64 blocks of 8 KB
  static constexpr int kVBlocks = (kDP / 2 + 63) / 64;         // 64-column blocks holding my kDP/2 columns of V
  static constexpr int kVHalfBytes = kVBlocks * kTileN * 64 * 2;   // (the last block is half used at 192 and 64)
  static constexpr int kSlotBytes = kKHalfBytes > kVHalfBytes ? kKHalfBytes : kVHalfBytes;  // one ring slot
  static constexpr int kStages = (kDP == 256) ? 4 : (kDP == 192) ? 5 : 8;
  static constexpr int kQ = 0;
  static constexpr int kKV = kQ + kQBytes;
  static constexpr int kBars = kKV + kStages * kSlotBytes;
  static constexpr int kNumBars = 11 + 2 * kStages;
  static constexpr int kMax = kBars + 8 * kNumBars + 16;   // float [2 parity][2 half][128]
  static constexpr int kFinal = kMax + 2 * 2 * 128 * 4;    // float [2 half][128] row sums
  static constexpr int kTotal = kFinal + 2 * 128 * 4 + 1024;  // + alignment slack
  static_assert(kTotal <= 232448, "shared memory budget");
};

template <int kDP, bool kBF16, bool kCausal>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWideThreads, 1)
fa_fwd_wide2_kernel(const __grid_constant__ CUtensorMap tmap_q,
                    const __grid_constant__ CUtensorMap tmap_k64,  // box {64 head-dim columns, 64 keys}
                    const __grid_constant__ CUtensorMap tmap_v,
                    const __grid_constant__ CUtensorMap tmap_o, const TcParams p) {
  using C = Wide2Cfg<kDP>;
  constexpr int kS = C::kStages;
  constexpr int kDBlocks = kDP / 64;
  constexpr int kKSteps = kDP / 16;
  constexpr int kOHalf = kDP / 2;
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

  // Barriers.  "leader" = only the copy in cluster rank 0 is used; "each" = one per CTA, the leader's
  // multicast commit arrives on both.
  const uint32_t bar_q_full = smem_u32(&bars[0]);                           // leader: tx of both Q tiles
  const uint32_t bar_o = smem_u32(&bars[1]);                                // each: commit after PV(j)
  auto bar_s_full = [&](int buf) { return smem_u32(&bars[2 + buf]); };      // each: commit
  auto bar_p_early = [&](int buf) { return smem_u32(&bars[4 + buf]); };     // leader: 16 softmax warps
  auto bar_p_mid = [&](int buf) { return smem_u32(&bars[6 + buf]); };
  auto bar_p_late = [&](int buf) { return smem_u32(&bars[8 + buf]); };
  const uint32_t bar_o_final = smem_u32(&bars[10]);                         // each: commit after the last PV
  auto bar_kv_full = [&](int s) { return smem_u32(&bars[11 + s]); };        // leader: tx of both halves
  auto bar_kv_empty = [&](int s) { return smem_u32(&bars[11 + kS + s]); };  // each: commit

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  // the pair is Q tiles (2p, 2p+1); the grid is padded to an even number of tiles; causal: longest pairs first
  int pair, h, b;
  work_coords<kCausal>(((p.Nq + kTileM - 1) / kTileM + 1) / 2, p.H, 2, pair, h, b);
  const int qtile = 2 * pair + static_cast<int>(rank);
  const int row0 = qtile * kTileM;
  // KV tiles: the two CTAs advance in lock step, so under a causal mask both visit the tiles the LATER Q
  // tile needs (2p + 2 of them); the extra tile is fully masked for the earlier one (P = 0, nothing added)
  int n = (p.Nkv + kTileN - 1) / kTileN;
  if (kCausal) n = min(n, 2 * pair + 2);

  auto idx_k = [](int j) { return j == 0 ? 0 : 2 * j - 1; };
  auto idx_v = [n](int j) { return (j + 1 < n) ? 2 * j + 2 : 2 * j + 1; };

  if (warp == 8 && lane == 0) {
    mbar_init(bar_q_full, 1);
    mbar_init(bar_o, 1);
    mbar_init(bar_o_final, 1);
#pragma unroll
    for (int buf = 0; buf < 2; ++buf) {
      mbar_init(bar_s_full(buf), 1);
      mbar_init(bar_p_early(buf), 16);
      mbar_init(bar_p_mid(buf), 16);
      mbar_init(bar_p_late(buf), 16);
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
    tma_prefetch_desc(&tmap_k64);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_o);
  }
  if (warp == 8) {
    tmem_alloc_2cta(smem_u32(tmem_slot), 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers are initialised before anything arrives on them remotely
  tc_fence_after();
  // PDL: everything above overlapped the previous kernel's tail; global memory is touched only below
  pdl_wait();
  pdl_launch_dependents();
  if (*tmem_slot != 0u) __trap();  // each CTA of the pair owns all of its SM's tensor memory
  constexpr uint32_t tmem = 0u;
  const float c = p.scale_log2;

  if (warp == 9) {
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
        if (!is_v) {  // my 64 keys of K_j: kDP/64 [64 keys x 64 columns] blocks, 8 KB apart
#pragma unroll
          for (int db = 0; db < kDBlocks; ++db)
            tma_load_4d_2cta(dst + db * 8192, &tmap_k64, full_leader, db * 64, j * kTileN + rank * 64, h, b);
        } else {      // my kDP/2 head-dim columns of V_j: [128 keys x 64 columns] blocks, 16 KB apart
#pragma unroll
          for (int db = 0; db < C::kVBlocks; ++db)
            tma_load_4d_2cta(dst + db * 16384, &tmap_v, full_leader, rank * (kDP / 2) + db * 64, j * kTileN, h, b);
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
  } else if (warp == 8) {
    // ========================================================================= MMA issuer (leader only)
    if (leader && elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(2 * kTileM, kTileN, kBF16, false, false);
      constexpr uint32_t idesc_o = make_idesc_f16(2 * kTileM, kDP, kBF16, false, true);
      auto wait_kv = [&](int idx) {
        mbar_wait(bar_kv_full(idx % kS), (idx / kS) & 1, 30);
        tc_fence_after();
      };
      auto release_kv = [&](int idx) { tc_commit_2cta(bar_kv_empty(idx % kS), 0b11); };
      auto issue_s = [&](int j) {  // S(j) = Q K_j^T for both CTAs into buffer j % 2
        const int idx = idx_k(j);
        wait_kv(idx);
        const uint32_t kb = sKV + (idx % kS) * C::kSlotBytes;
#pragma unroll
        for (int k = 0; k < kKSteps; ++k) {
          umma_ss_2cta(tmem + (j & 1) * 128,
                       make_smem_desc_sw128(sQ + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                       make_smem_desc_sw128(kb + (k >> 2) * 8192 + (k & 3) * 32, 16, 1024), idesc_s, k > 0);
        }
        tc_commit_2cta(bar_s_full(j & 1), 0b11);
        release_kv(idx);
      };
      auto issue_pv = [&](int j) {  // O += P(j) V_j for both CTAs
        const int idx = idx_v(j);
        const int buf = j & 1;
        const uint32_t par = (j >> 1) & 1;
        wait_kv(idx);
        const uint32_t vb = sKV + (idx % kS) * C::kSlotBytes;
        auto pv_step = [&](int ks, uint32_t acc) {
          umma_ts_2cta(tmem + kColO, tmem + buf * 128 + (ks >> 2) * 64 + (ks & 3) * 8,
                       make_smem_desc_sw128(vb + ks * 2048, 16384, 1024), idesc_o, acc);
        };
        if (j > 0) mbar_wait(bar_o, (j - 1) & 1, 35);  // see fa_fwd_wide.cuh
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
        tc_commit_2cta(bar_o, 0b11);
        release_kv(idx);
        if (j == n - 1) tc_commit_2cta(bar_o_final, 0b11);
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
    // ========================================================================= softmax warps 0-7 (each CTA)
    const int half = (warp >> 2) & 1;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tO = tmem + lane_base + kColO + half * kOHalf;
    const int pair_bar = 1 + (warp & 3);
    float* my_max = sMax + half * 128 + r;
    const float* other_max = sMax + (half ^ 1) * 128 + r;
    // the P hand-off barriers of the pair live in the leader
    // (buffer 1's barrier sits 8 bytes after buffer 0's, in the cluster window as in the CTA's)
    const uint32_t p_early0 = mapa_shared(bar_p_early(0), 0);
    const uint32_t p_mid0 = mapa_shared(bar_p_mid(0), 0);
    const uint32_t p_late0 = mapa_shared(bar_p_late(0), 0);

    float m_run = -INFINITY;
    float l_run = 0.f;

#pragma unroll 1
    for (int j = 0; j < n; ++j) {
      const int buf = j & 1;
      const uint32_t tS = tmem + lane_base + buf * 128 + half * 64;
      mbar_wait_warp(bar_s_full(buf), (j >> 1) & 1, 40);
      tc_fence_after();
      float s[64];
      tmem_ld_x32(tS, reinterpret_cast<uint32_t*>(s));
      tmem_ld_x32(tS + 32, reinterpret_cast<uint32_t*>(s) + 32);
      tmem_wait_ld();
      // causal: tile j >= qtile needs the mask; for j > qtile the row limit r + 1 - 128 (j - qtile) is <= 0,
      // i.e. every key of the tile is hidden
      ws_softmax_step<kDP, kBF16, true>(s, tS, tO, half, kCausal ? r - (j - qtile) * kTileN : r, lane,
                                        j * kTileN + half * 64, p.Nkv, kCausal && j >= qtile, c, m_run,
                                        l_run, j > 0, my_max + buf * 256, other_max + buf * 256, pair_bar,
                                        p_early0 + buf * 8, p_late0 + buf * 8, 0u, p_mid0 + buf * 8, bar_o,
                                        static_cast<uint32_t>((j - 1) & 1));
    }

    // ---- epilogue: O / l -> 16 bit -> swizzled smem (my Q buffer) -> TMA store
    sFinal[half * 128 + r] = l_run;
    named_bar_sync(pair_bar, 64);
    const float l_tot = l_run + sFinal[(half ^ 1) * 128 + r];
    const int row = row0 + r;
    if (half == 0 && p.lse != nullptr && row < p.Nq)
      p.lse[(static_cast<int64_t>(b) * p.H + h) * p.Nq + row] = m_run * c + log2f(l_tot);
    const float inv_l = 1.f / l_tot;
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
  cluster_sync_all();  // neither CTA leaves while the pair's barriers / tensor memory may still be in use
  if (warp == 8) tmem_dealloc_2cta(tmem, 512);
}

}  // namespace fa
