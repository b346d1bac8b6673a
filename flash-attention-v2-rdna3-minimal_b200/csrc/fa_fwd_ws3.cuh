// "ws3": the two-tile kernel on CTA pairs (fa_fwd_ws2.cuh) with P routed through SHARED memory so that
// S_t(j+1) no longer has to wait for O_t += P_t(j) V.  Head dim 128 (padded) or 64.
//
// In fa_fwd_ws.cuh / fa_fwd_ws2.cuh P (16 bit) overwrites the S columns in tensor memory, so the next score
// tile of a Q tile can only be issued after the PV product that reads P - the tail of the S -> softmax -> P ->
// PV -> S chain that bounds those kernels (each softmax group waits ~970 of ~2900 cycles per step for its S).
// Tensor memory has no room for a separate P (S0 S1 O0 O1 = 512 columns), but on CTA pairs shared memory has:
// every K/V tile costs an SM half the bytes, the ring shrinks to 4 half tiles (64 KB), and a 32 KB swizzled P
// tile per Q tile fits (Q 64 + ring 64 + P 64 = 192 KB).  Per-SM shared-memory traffic per step becomes
// 96 KB (S) + 96 KB (PV, now SS form) + 32 KB (TMA) + 64 KB (P stores) = 288 KB ~ 2250 cycles - the same
// scheme on single CTAs would need 384 KB ~ 3000 cycles, more than the step it is meant to shorten.
//
// RESULT (B200, fp16, H=16, D=128): correct (the forced-kernel parity matrix passes) but SLOWER than ws2:
// 1389 vs 1458 TFLOPS at N=16384, 1139 vs 1174 at N=4096.  The step gets longer, not shorter: three
// fence.proxy.async per tile on the softmax path, eight STS.128 instead of three tcgen05.st, an SS instead of a
// TS product (a third more operand reads), and the P tile must wait for PV_t(j-1) before it can be rewritten
// (freeing its "early" chunks with an extra commit in the middle of the PV product made it worse: 1305).
// Without the tensor-core wait nothing staggers the two softmax groups either.  Kept as FA_KERNEL_WS3
// (selectable, never chosen automatically) with these numbers.
//
// HEAD DIM 64 is different: S0 S1 O0 O1 take 384 columns there, which leaves room for a P region per tile that
// does not alias S.  The same protocol then needs neither shared memory nor proxy fences (tcgen05.st + TS product
// as in ws / ws2), and the early S issue is free - see the result line in DESIGN.md 3.6.
//
// Protocol changes against ws2:
//   - softmax warps signal "S_t(j) has been read" (after the lazy-rescale decision, which may re-read S) on a
//     leader barrier; the leader then issues S_t(j+1) at once, BEFORE PV_t(j)
//   - P parts are written with st.shared (128-byte-swizzled K-major tile, the layout TMA gives Q) +
//     fence.proxy.async; PV_t(j) is an SS MMA; a "PV_t(j) done" commit to both CTAs frees the P tile and, since
//     S_t(j+1) now precedes PV_t(j), is also what the rare O rescale waits for
//   - K/V ring in consumption order K0 K1 V0 K2 V1 ... (K(j+1) and V(j) are live together)
#pragma once
#include "fa_fwd_ws2.cuh"

namespace fa {

struct Ws3StepArgs {
  uint32_t bar_early, bar_mid, bar_late;  // P hand-off barriers (shared::cluster addresses in the leader)
  uint32_t bar_s_read;                    // "S_t(j) has been read" (shared::cluster address in the leader)
  uint32_t bar_pv_done;                   // this CTA's "PV_t(j) done" barrier
  uint8_t* p_row;                         // my 128-byte row inside the P block of my half (shared-memory P)
  uint32_t tP;                            // kDP = 64: TMEM address of my 32 P columns (P in spare tensor memory)
  int swz;                                // row & 7: the 128-byte swizzle XOR of my row
};

// One softmax step of one thread, P through shared memory (see the header; the arithmetic is that of
// ws_softmax_step in fa_fwd_ws.cuh, non-causal).
//   causal_tile / lim_c   this tile needs the causal mask: columns i >= lim_c of my half are hidden (lim_c <= 0:
//                         the whole half - the lock-step extra tiles of the earlier Q tiles of a pair)
//   kFirst                compile-time "first KV tile of a pass", as in ws_softmax_step (the persistent kernel peels it)
//   kNoMask               compile-time "neither the ragged last KV tile nor a causal tile": no mask code (see ws_softmax_step)
template <int kDP, bool kBF16, bool kFirst = false, bool kNoMask = false>
__device__ __forceinline__ void ws3_softmax_step(float (&s)[64], uint32_t tS, uint32_t tO, int lane, int col0,
                                                 int Nkv, float c, float& m_run, float& l_run, int j,
                                                 float* my_max, const float* other_max, int pair_bar,
                                                 const Ws3StepArgs& a, bool causal_tile = false, int lim_c = 64,
                                                 int have_o_flag = -1) {
  constexpr int kOHalf = kDP / 2;
  // j is the parity source of the per-step barriers; O_t holds a partial sum when j > 0 - unless the caller runs several
  // passes over one barrier sequence (the persistent kernel: j counts across units) and says so itself
  const bool have_o = have_o_flag < 0 ? (j > 0) : (have_o_flag != 0);
  const bool tail = !kNoMask && (col0 + 64 > Nkv);
  const bool masked = !kNoMask && (tail || causal_tile);
  int lim = 64;
  if (masked) {
    const int valid = tail ? (Nkv - col0) : 64;
    lim = causal_tile ? min(valid, lim_c) : valid;
#pragma unroll
    for (int i = 0; i < 64; ++i)
      if (i >= lim) s[i] = -INFINITY;
  }
  // Share of the exponentials on the FMA pipes.  At head dim 64 the softmax groups no longer wait for the tensor
  // cores, so the MUFU is the binding pipe and a larger share pays: 2 of 8 pairs 870 TFLOPS, 1 843, 3 833, 4 770.
  constexpr int kEmu = (kDP == 64) ? 2 : kEmuPairs;
  auto exp4 = [&](int i, float nmc_) {
    ffma2(s[i], s[i + 1], s[i], s[i + 1], c, c, nmc_, nmc_);
    ffma2(s[i + 2], s[i + 3], s[i + 2], s[i + 3], c, c, nmc_, nmc_);
    if ((((i >> 1) * kEmu) & 7) < kEmu) {
      ex2_fma2(s[i], s[i + 1]);
    } else {
      s[i] = ex2_approx(s[i]);
      s[i + 1] = ex2_approx(s[i + 1]);
    }
    if (((((i >> 1) + 1) * kEmu) & 7) < kEmu) {
      ex2_fma2(s[i + 2], s[i + 3]);
    } else {
      s[i + 2] = ex2_approx(s[i + 2]);
      s[i + 3] = ex2_approx(s[i + 3]);
    }
  };
  // 16-byte chunk ch (8 keys) of my row of the P block, packed from s[8 ch .. 8 ch + 8)
  auto store_chunk = [&](int ch) {
    uint4 v;
    v.x = pack2<kBF16>(s[8 * ch + 0], s[8 * ch + 1]);
    v.y = pack2<kBF16>(s[8 * ch + 2], s[8 * ch + 3]);
    v.z = pack2<kBF16>(s[8 * ch + 4], s[8 * ch + 5]);
    v.w = pack2<kBF16>(s[8 * ch + 6], s[8 * ch + 7]);
    *reinterpret_cast<uint4*>(a.p_row + ((ch ^ a.swz) << 4)) = v;
  };

  // columns [0,32) against the stale max while this tile's max is reduced and exchanged
  float nmc = -m_run * c;
  float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    mx0 = fmaxf(mx0, fmaxf(s[i], s[i + 32]));
    mx1 = fmaxf(mx1, fmaxf(s[i + 1], s[i + 33]));
    mx2 = fmaxf(mx2, fmaxf(s[i + 2], s[i + 34]));
    mx3 = fmaxf(mx3, fmaxf(s[i + 3], s[i + 35]));
    if constexpr (!kFirst) exp4(i, nmc);
  }
  const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
#if FA_MAX_XCHG_SHARED  // 32-bit shared-space accesses instead of generic ones (see ws_softmax_step)
  st_shared_f32(smem_u32(my_max), mx);
  named_bar_sync(pair_bar, 64);
  const float m_cand = fmaxf(fmaxf(mx, ld_shared_f32(smem_u32(other_max))), m_run);
#else
  *my_max = mx;
  named_bar_sync(pair_bar, 64);
  const float m_cand = fmaxf(fmaxf(mx, *other_max), m_run);
#endif
  const bool grow = (m_cand - m_run) * c > kRescaleThreshold;  // always true on the first tile
  float alpha = 1.f;
  if constexpr (kFirst) {
    if (grow) m_run = m_cand;  // (l_run is 0 and O_t empty: alpha is never used; s[] still holds the raw, masked scores)
    nmc = -m_run * c;
#pragma unroll
    for (int i = 0; i < 32; i += 4) exp4(i, nmc);
  } else if (__any_sync(0xffffffffu, grow)) {
    if (grow) {
      alpha = ex2_approx((m_run - m_cand) * c);
      m_run = m_cand;
    }
    if (have_o) {
      mbar_wait(a.bar_pv_done, (j - 1) & 1, 44);  // S_t(j) was issued before PV_t(j-1): wait for PV_t(j-1) itself
      tc_fence_after();
#pragma unroll 1
      for (int c8 = 0; c8 < kOHalf; c8 += 8) {
        uint32_t o[8];
        tmem_ld_x8(tO + c8, o);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st_x8(tO + c8, o);
      }
      tmem_wait_st();
    }
    nmc = -m_run * c;
    tmem_ld_x32(tS, reinterpret_cast<uint32_t*>(s));  // S is still intact in tensor memory
    tmem_wait_ld();
    if (masked) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i >= lim) s[i] = -INFINITY;
    }
#pragma unroll
    for (int i = 0; i < 32; i += 4) exp4(i, nmc);
  }
  // ---- S_t(j) is in registers for good: the leader may overwrite it with S_t(j+1)
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster(a.bar_s_read);
  // ---- the P tile is free once PV_t(j-1) has read it
  if (j > 0) mbar_wait(a.bar_pv_done, (j - 1) & 1, 45);

  // At head dim 64 tensor memory has 128 spare columns (S0 S1 O0 O1 take 384), enough for a P region per tile
  // that does not alias S: P goes there with tcgen05.st and feeds a TS product - no fences, no shared memory.
  constexpr bool kPTmem = (kDP == 64);
  // columns [16 lo, 16 lo + 16 n) of my half, i.e. keys [32 lo ..): n = 2 (32 keys) or 1 (16 keys)
  auto hand_off = [&](int key0, int nkeys, uint32_t bar) {
    if constexpr (kPTmem) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = (2 * i < nkeys) ? pack2<kBF16>(s[key0 + 2 * i], s[key0 + 2 * i + 1]) : 0u;
      if (nkeys == 32) {
        tmem_st_x16(a.tP + key0 / 2, pk);
      } else {
        uint32_t lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) lo[e] = pk[e];
        tmem_st_x8(a.tP + key0 / 2, lo);
      }
      tmem_wait_st();
      tc_fence_before();
    } else {
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        if (ch < nkeys / 8) store_chunk(key0 / 8 + ch);
      fence_proxy_async_smem();
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(bar);
  };

  // ---- first 32 keys of my half -> "early" hand-off
  hand_off(0, 32, a.bar_early);

  // ---- second half, with the row sum of the first half in the MUFU shadow; "mid" and "late" hand-offs
  float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll
  for (int i = 32; i < 64; i += 4) {
    exp4(i, nmc);
    fadd2(sum0, sum1, sum0, sum1, s[i - 32], s[i - 31]);
    fadd2(sum2, sum3, sum2, sum3, s[i - 30], s[i - 29]);
    if (i == 44) hand_off(32, 16, a.bar_mid);  // keys [32,48) of my half
  }
  hand_off(48, 16, a.bar_late);
#pragma unroll
  for (int i = 32; i < 64; i += 4) {
    fadd2(sum0, sum1, sum0, sum1, s[i], s[i + 1]);
    fadd2(sum2, sum3, sum2, sum3, s[i + 2], s[i + 3]);
  }
  l_run = l_run * alpha + ((sum0 + sum1) + (sum2 + sum3));
}

template <int kDP>
struct Ws3Cfg {
  static_assert(kDP == 64 || kDP == 128, "ws3 kernel: padded head dim 64 or 128");
  static constexpr int kTileBytes = kTileM * kDP * 2;          // one Q tile
  static constexpr int kKHalfBytes = (kTileN / 2) * kDP * 2;   // 64 keys x kDP: kDP/64 blocks of 8 KB
  static constexpr int kVHalfBytes = kTileN * 64 * 2;          // 128 keys x kDP/2 columns in one 64-column block
                                                               // (half used at kDP = 64)
  static constexpr int kSlotBytes = 16384;
  static constexpr int kStages = 4;
  static constexpr int kPBytes = kTileM * kTileN * 2;          // one P tile: [128 rows][128 keys] 16 bit
  static constexpr int kQ = 0;                                 // 2 Q tiles (re-used as O staging)
  static constexpr int kKV = kQ + 2 * kTileBytes;
  static constexpr int kP = kKV + kStages * kSlotBytes;        // 2 P tiles
  static constexpr int kBars = kP + 2 * kPBytes;
  static constexpr int kNumBars = 16 + 2 * kStages;
  static constexpr int kMax = kBars + 8 * kNumBars + 16;       // float [2 parity][2 tile][2 half][128]
  static constexpr int kFinal = kMax + 2 * 2 * 2 * 128 * 4;    // float [2 tile][2 half][128] row sums
  static constexpr int kTotal = kFinal + 2 * 2 * 128 * 4 + 1024;  // + alignment slack
  static_assert(kKHalfBytes <= kSlotBytes && kVHalfBytes <= kSlotBytes && kTotal <= 232448, "shared memory budget");
};

template <int kDP, bool kBF16, bool kCausal>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWsThreads, 1)
fa_fwd_ws3_kernel(const __grid_constant__ CUtensorMap tmap_q,
                  const __grid_constant__ CUtensorMap tmap_k64,  // box {64 head-dim columns, 64 keys}
                  const __grid_constant__ CUtensorMap tmap_v,
                  const __grid_constant__ CUtensorMap tmap_o, const TcParams p) {
  using C = Ws3Cfg<kDP>;
  constexpr int kS = C::kStages;
  constexpr int kDBlocks = kDP / 64;
  constexpr int kKSteps = kDP / 16;
  constexpr int kOHalf = kDP / 2;
  auto col_s = [](int t) -> uint32_t { return static_cast<uint32_t>(t) * 128u; };
  auto col_o = [](int t) -> uint32_t { return 256u + static_cast<uint32_t>(t) * 128u; };
  // head dim 64 only: O_t uses columns [256 + 128 t, +64); the 64 columns after it hold P_t (2 halves x 32)
  auto col_p = [](int t) -> uint32_t { return 320u + static_cast<uint32_t>(t) * 128u; };

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t sQ = smem_u32(smem + C::kQ);
  const uint32_t sKV = smem_u32(smem + C::kKV);
  const uint32_t sP = smem_u32(smem + C::kP);
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
  auto bar_s_read = [&](int t) { return smem_u32(&bars[12 + t]); };         // leader: 16 softmax warps
  auto bar_pv_done = [&](int t) { return smem_u32(&bars[14 + t]); };        // each: commit after PV_t(j)
  auto bar_kv_full = [&](int s) { return smem_u32(&bars[16 + s]); };        // leader: tx of both halves
  auto bar_kv_empty = [&](int s) { return smem_u32(&bars[16 + kS + s]); };  // each

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  // 256-row query block; the pair is blocks (2p, 2p+1), grid padded to even; causal: longest pairs first
  int pair, h, b;
  work_coords<kCausal>(((p.Nq + 2 * kTileM - 1) / (2 * kTileM) + 1) / 2, p.H, 2, pair, h, b);
  const int blk = 2 * pair + static_cast<int>(rank);
  const int row0 = blk * 2 * kTileM;
  // KV tiles: the four Q tiles of the pair advance in lock step, so under a causal mask all visit the tiles the LAST
  // one needs (4 pair + 4); the extra tiles are fully masked for the earlier ones (P = 0, nothing is added)
  int n_ = (p.Nkv + kTileN - 1) / kTileN;
  if (kCausal) n_ = min(n_, 4 * pair + 4);
  const int n = n_;
  auto idx_k = [](int j) { return j == 0 ? 0 : 2 * j - 1; };  // ring (= consumption) order K0 K1 V0 K2 V1 ...
  auto idx_v = [n](int j) { return (j + 1 < n) ? 2 * j + 2 : 2 * j + 1; };

  if (warp == 16 && lane == 0) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_q_full(t), 1);
      mbar_init(bar_s_full(t), 1);
      mbar_init(bar_p_early(t), 16);
      mbar_init(bar_p_mid(t), 16);
      mbar_init(bar_p_late(t), 16);
      mbar_init(bar_o_final(t), 1);
      mbar_init(bar_s_read(t), 16);
      mbar_init(bar_pv_done(t), 1);
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
        auto load = [&](bool is_v, int j, int idx) {
          const int slot = idx % kS;
          mbar_wait(bar_kv_empty(slot), ((idx / kS) & 1) ^ 1, 20);
          const uint32_t full_leader = mapa_shared(bar_kv_full(slot), 0);
          const uint32_t dst = sKV + slot * C::kSlotBytes;
          if (!is_v) {  // my 64 keys of K_j: kDP/64 [64 keys x 64 columns] blocks, 8 KB apart
            if (leader) mbar_arrive_expect_tx(bar_kv_full(slot), 2 * C::kKHalfBytes);
#pragma unroll
            for (int db = 0; db < kDBlocks; ++db)
              tma_load_4d_2cta(dst + db * 8192, &tmap_k64, full_leader, db * 64, j * kTileN + rank * 64, h, b);
          } else {      // my kDP/2 head-dim columns of V_j: one [128 keys x 64 columns] block
            if (leader) mbar_arrive_expect_tx(bar_kv_full(slot), 2 * C::kVHalfBytes);
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
          const uint32_t k_lo = smem_desc_lo(sKV + (idx_k(j) % kS) * C::kSlotBytes, 16);
          const uint32_t q_lo = smem_desc_lo(sQ + t * C::kTileBytes, 16);
#pragma unroll
          for (int k = 0; k < kKSteps; ++k) {
            const uint32_t q_off = ((k >> 2) * 16384 + (k & 3) * 32) >> 4;
            const uint32_t k_off = ((k >> 2) * 8192 + (k & 3) * 32) >> 4;
            umma_ss2_2cta(tmem + col_s(t), q_lo + q_off, desc_hi, k_lo + k_off, desc_hi, idesc_s, k > 0);
          }
          tc_commit_2cta(bar_s_full(t), 0b11);
        };
        // k-step ks covers keys [16 ks, 16 ks + 16): P columns 16 ks.. of the K-major P tile (block ks / 4),
        // V rows 16 ks of the MN-major half tile
        auto pv_step = [&](int t, uint32_t p_lo, uint32_t v_lo, int ks, uint32_t acc) {
          if constexpr (kDP == 64) {  // P in spare tensor memory: half ks / 4 at columns col_p(t) + 32 (ks/4) + 8 (ks%4)
            umma_ts2_2cta(tmem + col_o(t), tmem + col_p(t) + (ks >> 2) * 32 + (ks & 3) * 8,
                          v_lo + ((ks * 2048) >> 4), desc_hi, idesc_o, acc);
          } else {
            umma_ss2_2cta(tmem + col_o(t), p_lo + (((ks >> 2) * 16384 + (ks & 3) * 32) >> 4), desc_hi,
                          v_lo + ((ks * 2048) >> 4), desc_hi, idesc_o, acc);
          }
        };
        auto issue_pv = [&](int t, int j) {  // O_t += P_t V_j for both CTAs
          const uint32_t v_lo = smem_desc_lo(sKV + (idx_v(j) % kS) * C::kSlotBytes, 16384);
          const uint32_t p_lo = smem_desc_lo(sP + t * C::kPBytes, 16);
          mbar_wait(bar_p_early(t), j & 1, 31 + t);
          tc_fence_after();
          pv_step(t, p_lo, v_lo, 0, j > 0);
          pv_step(t, p_lo, v_lo, 1, 1);
          pv_step(t, p_lo, v_lo, 4, 1);
          pv_step(t, p_lo, v_lo, 5, 1);
          mbar_wait(bar_p_mid(t), j & 1, 37 + t);
          tc_fence_after();
          pv_step(t, p_lo, v_lo, 2, 1);
          pv_step(t, p_lo, v_lo, 6, 1);
          mbar_wait(bar_p_late(t), j & 1, 35 + t);
          tc_fence_after();
          pv_step(t, p_lo, v_lo, 3, 1);
          pv_step(t, p_lo, v_lo, 7, 1);
          tc_commit_2cta(bar_pv_done(t), 0b11);
          if (j == n - 1) tc_commit_2cta(bar_o_final(t), 0b11);
        };

        wait_kv(idx_k(0));
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(bar_q_full(t), 0, 33);
          tc_fence_after();
          issue_s(t, 0);
        }
        release_kv(idx_k(0));
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
          const int nx = j + 1;
          if (nx < n) {
            wait_kv(idx_k(nx));
            mbar_wait(bar_s_read(0), j & 1, 38);  // both CTAs' tile-0 warps hold S_0(j) in registers
            tc_fence_after();
            issue_s(0, nx);
          }
          wait_kv(idx_v(j));
          issue_pv(0, j);
          if (nx < n) {
            mbar_wait(bar_s_read(1), j & 1, 39);
            tc_fence_after();
            issue_s(1, nx);
            release_kv(idx_k(nx));
          }
          issue_pv(1, j);
          release_kv(idx_v(j));
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
    Ws3StepArgs a;
    a.bar_early = mapa_shared(bar_p_early(t), 0);  // the pair's hand-off barriers: in the leader
    a.bar_mid = mapa_shared(bar_p_mid(t), 0);
    a.bar_late = mapa_shared(bar_p_late(t), 0);
    a.bar_s_read = mapa_shared(bar_s_read(t), 0);
    a.bar_pv_done = bar_pv_done(t);
    a.p_row = smem + C::kP + t * C::kPBytes + half * 16384 + r * 128;  // my 128-byte row of the P block of my half
    a.swz = r & 7;
    a.tP = tmem + lane_base + col_p(t) + half * 32;

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
      const int g = 2 * blk + t;  // my Q tile's index: KV tile j >= g needs the causal mask
      ws3_softmax_step<kDP, kBF16>(s, tS, tO, lane, j * kTileN + half * 64, p.Nkv, c, m_run, l_run, j,
                                   my_max + (j & 1) * 512, other_max + (j & 1) * 512, pair_bar, a,
                                   kCausal && j >= g, r - (j - g) * kTileN + 1 - half * 64);
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
