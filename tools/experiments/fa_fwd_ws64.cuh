// Two-tile warp-specialised forward kernel with 64-KEY score tiles ("ws64"): head dims 65..128.
//
// What bounds fa_fwd_ws.cuh at head dim 128 is not a pipe but a chain: P_t(j) overwrites S_t(j) in tensor
// memory, so S_t(j+1) cannot be issued before O_t += P_t(j) V has been, and each Q tile's softmax waits for the
// tensor cores once per KV tile (S -> softmax -> P -> PV -> S: ~2900 cycles per 128 keys for 2048 cycles of MMAs,
// tensor pipe 69 % busy).  At head dim 64 there are spare TMEM columns and giving P its own region
// was worth +20 % (fa_fwd_ws3.cuh).  At head dim 128 the two accumulators take 256 of the 512 columns; the
// only tiling that leaves room for a SECOND score buffer per Q tile is 64 keys per step:
//
//   TMEM   S_0 buf0 [0,64)  S_0 buf1 [64,128)  S_1 buf0 [128,192)  S_1 buf1 [192,256)  O_0 [256,384)  O_1 [384,512)
//
// so S_t(i+1) is computed into the other buffer while the softmax warps work on S_t(i), and a softmax group
// finds its next scores ready when it finishes a step.  The price: a 128x64x16 MMA takes 48 tensor cycles, not
// 32 (tools/microbench_umma.cu), i.e. 2560 instead of 2048 tensor cycles per 128 keys and two Q tiles - still
// below what the chain costs today.
//
// Roles and the softmax algorithm are those of fa_fwd_ws.cuh (16 softmax warps, two threads per query row - now
// 32 keys each per step -, stale-max speculation on the first half, lazy rescale, P handed over in two parts,
// one MMA warp, one TMA warp).  Tensor-core issue order per Q tile:  S(0) S(1) | PV(0) S(2) | PV(1) S(3) | ...,
// the two tiles interleaved.  K/V ring: 8 slots of one 64-key tile (16 KB) in consumption order
// K0 K1 V0 K2 V1 K3 ...  Because S_t(i) is issued before PV_t(i-1), the rare O rescale waits on the barrier
// PV_t(i-1) commits to (as in fa_fwd_wide.cuh).
//
// Replaces /root/reference/rocwmma_fattn/kernel_fp16.cu:306-544 / kernel_bf16.cu:329-576.
#pragma once
#include "fa_fwd_ws.cuh"

namespace fa {

constexpr int kTileN64 = 64;

struct Ws64Cfg {
  static constexpr int kDP = 128;
  static constexpr int kTileBytes = kTileM * kDP * 2;        // one Q tile (32 KB)
  static constexpr int kKVBytes = kTileN64 * kDP * 2;        // one 64-key K or V tile (16 KB)
  static constexpr int kStages = 8;
  static constexpr int kQ = 0;                               // 2 Q tiles (re-used as O staging)
  static constexpr int kKV = kQ + 2 * kTileBytes;
  static constexpr int kBars = kKV + kStages * kKVBytes;
  static constexpr int kNumBars = 18 + 2 * kStages;
  static constexpr int kMax = kBars + 8 * kNumBars + 16;     // float [2 parity][2 tile][2 half][128]
  static constexpr int kFinal = kMax + 2 * 2 * 2 * 128 * 4;  // float [2 tile][2 half][128] row sums
  static constexpr int kTotal = kFinal + 2 * 2 * 128 * 4 + 1024;  // + alignment slack
  static_assert(kTotal <= 232448, "shared memory budget");
};

// One softmax step of one thread on its 32 scores of a 64-key tile (see ws_softmax_step for the algorithm).
//   tS       TMEM address of my 32 S columns (P goes over the first 16)
//   lim      number of visible keys among my 32 (32 = no mask; <= 0 = all hidden)
//   bar_o    mbarrier (parity o_parity) that tells PV(i-1) of my tile has left the tensor cores
template <bool kBF16>
__device__ __forceinline__ void ws64_softmax_step(float (&s)[32], uint32_t tS, uint32_t tO, int lane, int lim, float c,
                                                  float& m_run, float& l_run, bool have_o, float* my_max,
                                                  const float* other_max, int pair_bar, uint32_t bar_early,
                                                  uint32_t bar_late, uint32_t bar_o, uint32_t o_parity) {
  const bool masked = lim < 32;
  if (masked) {
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

  // columns [0,16) against the running max of the previous tiles while this tile's max is being reduced
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
  *my_max = mx;
  named_bar_sync(pair_bar, 64);
  const float m_cand = fmaxf(fmaxf(mx, *other_max), m_run);
  const bool grow = (m_cand - m_run) * c > kRescaleThreshold;  // always true on the first tile
  float alpha = 1.f;
  if (__any_sync(0xffffffffu, grow)) {
    if (grow) {
      alpha = ex2_approx((m_run - m_cand) * c);
      m_run = m_cand;
    }
    if (have_o) {
      mbar_wait(bar_o, o_parity, 44);  // PV(i-1) has completed; PV(i) waits for my P
      tc_fence_after();
#pragma unroll 1
      for (int c8 = 0; c8 < 64; c8 += 8) {
        uint32_t o[8];
        tmem_ld_x8(tO + c8, o);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st_x8(tO + c8, o);
      }
    }
    nmc = -m_run * c;
    tmem_ld_x16(tS, reinterpret_cast<uint32_t*>(s));  // S is still intact in TMEM (no P stored yet)
    tmem_wait_ld();
    if (masked) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i >= lim) s[i] = -INFINITY;
    }
#pragma unroll
    for (int i = 0; i < 16; i += 4) exp4(i, nmc);
  }
  {
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pk[i] = pack2<kBF16>(s[2 * i], s[2 * i + 1]);
    tmem_st_x8(tS, pk);
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_early);
  }
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
    if (lane == 0) mbar_arrive(bar_late);
  }
#pragma unroll
  for (int i = 16; i < 32; i += 4) {
    fadd2(sum0, sum1, sum0, sum1, s[i], s[i + 1]);
    fadd2(sum2, sum3, sum2, sum3, s[i + 2], s[i + 3]);
  }
  l_run = l_run * alpha + ((sum0 + sum1) + (sum2 + sum3));
}

template <bool kBF16, bool kCausal>
__global__ void __launch_bounds__(kWsThreads, 1)
fa_fwd_ws64_kernel(const __grid_constant__ CUtensorMap tmap_q,
                   const __grid_constant__ CUtensorMap tmap_k64,  // box {64 head-dim columns, 64 keys}
                   const __grid_constant__ CUtensorMap tmap_v64,  // box {64 head-dim columns, 64 keys}
                   const __grid_constant__ CUtensorMap tmap_o, const TcParams p) {
  using C = Ws64Cfg;
  constexpr int kDP = C::kDP;
  constexpr int kS = C::kStages;
  constexpr int kDBlocks = kDP / 64;
  constexpr int kKSteps = kDP / 16;
  constexpr int kOHalf = kDP / 2;
  auto col_s = [](int t, int b) -> uint32_t { return static_cast<uint32_t>(t * 2 + b) * 64u; };
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

  auto bar_q_full = [&](int t) { return smem_u32(&bars[t]); };                   // tx, count 1
  auto bar_o = [&](int t) { return smem_u32(&bars[2 + t]); };                    // tcgen05.commit after PV_t(i)
  auto bar_o_final = [&](int t) { return smem_u32(&bars[4 + t]); };              // tcgen05.commit after the last PV_t
  auto bar_s_full = [&](int t, int b) { return smem_u32(&bars[6 + t * 2 + b]); };   // tcgen05.commit
  auto bar_p_early = [&](int t, int b) { return smem_u32(&bars[10 + t * 2 + b]); };  // 8 softmax warps
  auto bar_p_late = [&](int t, int b) { return smem_u32(&bars[14 + t * 2 + b]); };   // 8 softmax warps
  auto bar_kv_full = [&](int s) { return smem_u32(&bars[18 + s]); };             // tx, count 1
  auto bar_kv_empty = [&](int s) { return smem_u32(&bars[18 + kS + s]); };       // tcgen05.commit

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  int pair, h, b;  // causal: longest blocks first across the whole launch (work_coords)
  work_coords<kCausal>((p.Nq + 2 * kTileM - 1) / (2 * kTileM), p.H, 1, pair, h, b);
  const int row0 = pair * 2 * kTileM;
#ifdef FA_TRACE
  const bool tr_cta = p.trace != nullptr && pair == (p.Nq + 2 * kTileM - 1) / (2 * kTileM) / 2 && h == 0 && b == 0;
  const bool tr_on = tr_cta && (warp == 16 || (warp & 7) == 0);
#endif

  // per-tile trip counts in 64-key tiles; tile 1 never visits fewer than tile 0
  const int n_kv_total = (p.Nkv + kTileN64 - 1) / kTileN64;
  int n_t[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int r0 = row0 + t * kTileM;
    int n = (r0 < p.Nq) ? n_kv_total : 0;
    if (kCausal) n = min(n, (r0 + kTileM - 1) / kTileN64 + 1);
    n_t[t] = n;
  }
  const int n_max = max(n_t[0], n_t[1]);
  // position of a tile in the ring (= consumption) order K0 K1 V0 K2 V1 ... K(n-1) V(n-2) V(n-1)
  auto idx_k = [](int j) { return j == 0 ? 0 : 2 * j - 1; };
  auto idx_v = [n_max](int j) { return (j + 1 < n_max) ? 2 * j + 2 : 2 * j + 1; };

  if (warp == 16 && lane == 0) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_q_full(t), 1);
      mbar_init(bar_o(t), 1);
      mbar_init(bar_o_final(t), 1);
#pragma unroll
      for (int bf = 0; bf < 2; ++bf) {
        mbar_init(bar_s_full(t, bf), 1);
        mbar_init(bar_p_early(t, bf), 8);
        mbar_init(bar_p_late(t, bf), 8);
      }
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
    tma_prefetch_desc(&tmap_v64);
    tma_prefetch_desc(&tmap_o);
    // the Q tiles and the first K/V tiles -> L2, before pdl_wait() (see fa_fwd_ws.cuh)
#pragma unroll
    for (int db = 0; db < kDBlocks; ++db) {
#pragma unroll
      for (int t = 0; t < 2; ++t)
        if (n_t[t] > 0) tma_prefetch_l2_4d(&tmap_q, db * 64, row0 + t * kTileM, h, b);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < n_max) {
          tma_prefetch_l2_4d(&tmap_k64, db * 64, j * kTileN64, h, b);
          tma_prefetch_l2_4d(&tmap_v64, db * 64, j * kTileN64, h, b);
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
  // PDL: everything above overlapped the previous kernel's tail; global memory is touched only below
  pdl_wait();
  pdl_launch_dependents();
  if (*tmem_slot != 0u) __trap();  // constant TMEM addresses (see fa_fwd_ws.cuh)
  constexpr uint32_t tmem = 0u;
  const float c = p.scale_log2;

  if (warp >= 16) {
    setmaxnreg_dec<32>();
    if (warp == 17) {
      // ========================================================================= TMA producer
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (n_t[t] > 0) {
            mbar_arrive_expect_tx(bar_q_full(t), C::kTileBytes);
#pragma unroll
            for (int db = 0; db < kDBlocks; ++db)
              tma_load_4d(sQ + t * C::kTileBytes + db * 16384, &tmap_q, bar_q_full(t), db * 64,
                          row0 + t * kTileM, h, b);
          }
        }
        auto load = [&](const CUtensorMap* map, int j, int idx) {
          const int slot = idx % kS;
          mbar_wait(bar_kv_empty(slot), ((idx / kS) & 1) ^ 1, 20);
          mbar_arrive_expect_tx(bar_kv_full(slot), C::kKVBytes);
#pragma unroll
          for (int db = 0; db < kDBlocks; ++db)
            tma_load_4d(sKV + slot * C::kKVBytes + db * 8192, map, bar_kv_full(slot), db * 64, j * kTileN64, h, b);
        };
        if (n_max > 0) load(&tmap_k64, 0, 0);
#pragma unroll 1
        for (int j = 0; j < n_max; ++j) {
          if (j + 1 < n_max) load(&tmap_k64, j + 1, idx_k(j + 1));
          load(&tmap_v64, j, idx_v(j));
        }
      }
      __syncwarp();
    } else if (warp == 16) {
      // ========================================================================= MMA issuer
      if (elect_one()) {
        constexpr uint32_t idesc_s = make_idesc_f16(kTileM, kTileN64, kBF16, false, false);
        constexpr uint32_t idesc_o = make_idesc_f16(kTileM, kDP, kBF16, false, true);
        constexpr uint32_t desc_hi = smem_desc_hi_sw128(1024);
        auto wait_kv = [&](int idx) {
          mbar_wait(bar_kv_full(idx % kS), (idx / kS) & 1, 30);
          tc_fence_after();
        };
        auto release_kv = [&](int idx) { tc_commit(bar_kv_empty(idx % kS)); };
        auto issue_s = [&](int t, int j) {  // S_t(j) = Q_t K_j^T into buffer j % 2 (K_j has been waited for)
          const uint32_t k_lo = smem_desc_lo(sKV + (idx_k(j) % kS) * C::kKVBytes, 16);
          const uint32_t q_lo = smem_desc_lo(sQ + t * C::kTileBytes, 16);
#pragma unroll
          for (int k = 0; k < kKSteps; ++k) {
            const uint32_t q_off = ((k >> 2) * 16384 + (k & 3) * 32) >> 4;
            const uint32_t k_off = ((k >> 2) * 8192 + (k & 3) * 32) >> 4;
            umma_ss2(tmem + col_s(t, j & 1), q_lo + q_off, desc_hi, k_lo + k_off, desc_hi, idesc_s, k > 0);
          }
          tc_commit(bar_s_full(t, j & 1));
        };
        // k-step ks covers keys [16 ks, 16 ks + 16) of the 64-key tile: P of half ks / 2 at column
        // 32 (ks / 2) + 8 (ks % 2) of the S buffer; V rows 16 ks (2048 bytes apart in the MN-major tile)
        auto issue_pv = [&](int t, int j) {  // O_t += P_t(j) V_j (V_j has been waited for)
          const int bf = j & 1;
          const uint32_t par = (j >> 1) & 1;
          const uint32_t v_lo = smem_desc_lo(sKV + (idx_v(j) % kS) * C::kKVBytes, 8192);
          auto pv_step = [&](int ks, uint32_t acc) {
            umma_ts2(tmem + col_o(t), tmem + col_s(t, bf) + (ks >> 1) * 32 + (ks & 1) * 8,
                     v_lo + ((ks * 2048) >> 4), desc_hi, idesc_o, acc);
          };
          // observe PV_t(j-1)'s phase of bar_o before arming the next one (its waiters are at most one phase behind)
          if (j > 0) mbar_wait(bar_o(t), (j - 1) & 1, 35 + t);
          mbar_wait(bar_p_early(t, bf), par, 31 + t);
          tc_fence_after();
          FA_TR(2, j, 2 + 3 * t);
          pv_step(0, j > 0);
          pv_step(2, 1);
          mbar_wait(bar_p_late(t, bf), par, 33 + t);
          tc_fence_after();
          pv_step(1, 1);
          pv_step(3, 1);
          tc_commit(bar_o(t));
          if (j == n_t[t] - 1) tc_commit(bar_o_final(t));
        };

        if (n_max > 0) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (n_t[t] > 0) {
              mbar_wait(bar_q_full(t), 0, 37);
              tc_fence_after();
            }
          }
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (j < n_max) {
              wait_kv(idx_k(j));
              if (j < n_t[0]) issue_s(0, j);
              if (j < n_t[1]) issue_s(1, j);
              release_kv(idx_k(j));
            }
          }
        }
#pragma unroll 1
        for (int j = 0; j < n_max; ++j) {
          const int nx = j + 2;
          FA_TR(2, j, 0);
          wait_kv(idx_v(j));
          FA_TR(2, j, 1);
          if (j < n_t[0]) issue_pv(0, j);
          FA_TR(2, j, 3);
          if (nx < n_max) wait_kv(idx_k(nx));
          if (nx < n_t[0]) issue_s(0, nx);
          FA_TR(2, j, 4);
          if (j < n_t[1]) issue_pv(1, j);
          FA_TR(2, j, 6);
          release_kv(idx_v(j));
          if (nx < n_t[1]) issue_s(1, nx);
          if (nx < n_max) release_kv(idx_k(nx));
          FA_TR(2, j, 7);
        }
      }
      __syncwarp();
    }
  } else {
    // ========================================================================= softmax warps (0-7: tile 0, 8-15: tile 1)
    setmaxnreg_inc<112>();
    const int t = warp >> 3;
    const int half = (warp >> 2) & 1;
    const int r = (warp & 3) * 32 + lane;  // query row inside the tile = TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tO = tmem + lane_base + col_o(t) + half * kOHalf;
    const int pair_bar = 1 + t * 4 + (warp & 3);
    const int tile_row0 = row0 + t * kTileM;
    const int n = n_t[t];
    float* my_max = sMax + (t * 2 + half) * 128 + r;
    const float* other_max = sMax + (t * 2 + (half ^ 1)) * 128 + r;

    float m_run = -INFINITY;
    float l_run = 0.f;  // partial row sum over my key halves

#pragma unroll 1
    for (int j = 0; j < n; ++j) {
      const int bf = j & 1;
      const uint32_t tS = tmem + lane_base + col_s(t, bf) + half * 32;  // my 32 S columns; P over [0,16)
      FA_TR(t, j, 0);
      mbar_wait_warp(bar_s_full(t, bf), (j >> 1) & 1, 40 + t);
      tc_fence_after();
      FA_TR(t, j, 1);
      float s[32];
      tmem_ld_x32(tS, reinterpret_cast<uint32_t*>(s));
      tmem_wait_ld();
      FA_TR(t, j, 2);
      const int col0 = j * kTileN64 + half * 32;
      int lim = min(32, p.Nkv - col0);
      if (kCausal) lim = min(lim, tile_row0 + r - col0 + 1);
      ws64_softmax_step<kBF16>(s, tS, tO, lane, lim, c, m_run, l_run, j > 0, my_max + bf * 512, other_max + bf * 512,
                               pair_bar, bar_p_early(t, bf), bar_p_late(t, bf), bar_o(t),
                               static_cast<uint32_t>((j - 1) & 1));
      FA_TR(t, j, 6);
    }

    // ---- epilogue: O / l -> 16 bit -> swizzled smem (the tile's Q buffer) -> TMA store
    if (n > 0) {
      sFinal[(t * 2 + half) * 128 + r] = l_run;
      named_bar_sync(pair_bar, 64);
      const float l_tot = l_run + sFinal[(t * 2 + (half ^ 1)) * 128 + r];
      const int row = tile_row0 + r;
      if (half == 0 && p.lse != nullptr && row < p.Nq)
        p.lse[(static_cast<int64_t>(b) * p.H + h) * p.Nq + row] = m_run * c + log2f(l_tot);
      const float inv_l = 1.f / l_tot;
      mbar_wait(bar_o_final(t), 0, 54 + t);  // every MMA of this tile is done: Q_t is free too
      tc_fence_after();
      uint8_t* stage = smem + C::kQ + t * C::kTileBytes;
#pragma unroll
      for (int cidx = 0; cidx < kOHalf / 32; ++cidx) {
        uint32_t o[32];
        tmem_ld_x32(tO + cidx * 32, o);
        tmem_wait_ld();
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint4 val;
          val.x = pack2<kBF16>(__uint_as_float(o[ch * 8 + 0]) * inv_l, __uint_as_float(o[ch * 8 + 1]) * inv_l);
          val.y = pack2<kBF16>(__uint_as_float(o[ch * 8 + 2]) * inv_l, __uint_as_float(o[ch * 8 + 3]) * inv_l);
          val.z = pack2<kBF16>(__uint_as_float(o[ch * 8 + 4]) * inv_l, __uint_as_float(o[ch * 8 + 5]) * inv_l);
          val.w = pack2<kBF16>(__uint_as_float(o[ch * 8 + 6]) * inv_l, __uint_as_float(o[ch * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(stage + sw128_offset_16bit(r, half * kOHalf + cidx * 32 + ch * 8)) = val;
        }
      }
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 512);
}

}  // namespace fa
