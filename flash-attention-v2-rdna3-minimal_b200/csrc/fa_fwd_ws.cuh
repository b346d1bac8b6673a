// Warp-specialised tcgen05 forward kernel ("ws"): the fast path.
//
// One CTA = one (batch, head, 256-row Q block) = two 128-row Q tiles that ping-pong on the tensor
// cores: while the softmax warps of tile 0 work on S0(j) in registers, the tensor cores compute
// S1(j) / O1 += P1 V, and vice versa.  20 warps:
//
//   warps  0- 7  softmax for Q tile 0: warp w owns TMEM lanes 32*(w%4).. (query rows) and the
//                64-key half (w/4)%2 of every S tile, so one query row is shared by two threads
//                on the same SM sub-partition; their partial row maxima meet through shared memory
//   warps  8-15  softmax for Q tile 1
//   warp  16     MMA issuer (one elected thread issues every tcgen05.mma)
//   warp  17     TMA producer (Q once, then the K/V ring)
//   warps 18-19  idle (they only donate their registers)
//
// Why two threads per row: the exp / max / pack stream of one 128-key row is ~650 dependent-issue
// instructions; a single warp issues them at ~0.35 IPC, which made S -> P -> next S the critical
// chain.  Two warps per row-quarter halve the chain and let the scheduler overlap MUFU, FMA and ALU
// work of the two halves.  The softmax warps also rescale O (rarely: lazy rescale) and run the
// epilogue, so there is no separate correction warpgroup polling barriers.
//
// TMEM (512 columns x 128 lanes x 32 bit):  S0 [0,128)  S1 [128,256)  O0 [256,256+D)  O1 [384,384+D).
// P (16-bit) of key half h overwrites S columns [64h, 64h+32) - inside the S columns the same
// thread loaded - and feeds the PV MMA from TMEM (TS form).
//
// MMA issue order (K/V ring order is K0 V0 K1 V1 ...):
//   S0(0) S1(0) | PV0(0) S0(1) PV1(0) S1(1) | PV0(1) S0(2) PV1(1) S1(2) | ...
// PV is issued in two parts: the k-steps whose P columns were stored first ("early" barrier), then
// the rest ("late"), so the tensor cores start on O += P V while the second half of the
// exponentials is still in flight.
//
// The softmax step (ws_softmax_step) is compiled three times and a pass over the KV tiles runs
//   first tile (kFirst: row max first, every exponential once) | interior tiles (kNoMask: no mask code) | last tile (generic)
// - the causal diagonal and the ragged tail can only be the last tile.  The MMA thread likewise runs a loop without
// per-tile conditions while both Q tiles have a further KV tile to visit.  What such splits do to ptxas' schedule of the
// hot loop has to be measured per kernel: DESIGN.md 3.1 / 3.6 list where they won (here, wide, the persistent kernel's
// head-dim-64 path) and where they lost (the persistent kernel's head-dim-128 path, the early-S kernel on pairs).
//
// Replaces /root/reference/rocwmma_fattn/kernel_fp16.cu:306-544 / kernel_bf16.cu:329-576 and the
// device GEMM helpers (:115-302); see fa_fwd_tc.cuh for the serial version of the same algorithm.
#pragma once
#include <type_traits>

#include "fa_fwd_tc.cuh"

namespace fa {

constexpr int kWsThreads = 640;

template <int kDP>
struct WsCfg {
  static constexpr int kTileBytes = kTileM * kDP * 2;
  static constexpr int kStages = (kDP == 128) ? 4 : 8;  // K/V ring slots (one K or one V tile each)
  static constexpr int kQ = 0;                           // 2 Q tiles (re-used as O staging)
  static constexpr int kKV = kQ + 2 * kTileBytes;
  static constexpr int kBars = kKV + kStages * kTileBytes;
  static constexpr int kNumBars = 14 + 2 * kStages;
  static constexpr int kMax = kBars + 8 * kNumBars + 16;      // float [2 parity][2 tile][2 half][128]
  static constexpr int kFinal = kMax + 2 * 2 * 2 * 128 * 4;   // float [2 tile][2 half][128] row sums
  static constexpr int kTotal = kFinal + 2 * 2 * 128 * 4 + 1024;  // + alignment slack
};

// Debug-only timeline (-DFA_TRACE, never in the shipped library): lane 0 of each role's first warp
// stamps clock64() into p.trace[role][j][event] for one CTA; see tools/trace_ws.py.
#ifdef FA_TRACE
#define FA_TR(role, j, ev)                                                                  \
  do {                                                                                      \
    if (tr_on && lane == 0) p.trace[((role) * 128 + ((j) & 127)) * 8 + (ev)] = clock64();   \
  } while (0)
// (the same inside ws_softmax_step: stamps of the step's phases, tools/trace_ws_softmax.py; note that the stamps change
// what ptxas does with the step - the traced kernel issues its first-half exponentials before the pair barrier)
#define FA_TRS_PARAM , unsigned long long* trp = nullptr
#define FA_TRS(ev) do { if (trp != nullptr && lane == 0) trp[ev] = clock64(); } while (0)
#else
#define FA_TR(role, j, ev) do { } while (0)
#define FA_TRS_PARAM
#define FA_TRS(ev) do { } while (0)
#endif

constexpr float kRescaleThreshold = 8.0f;  // log2 units: rescale O only when the max grew by > 2^8

// Of every 8 pairs of P elements, how many compute 2^x on the FMA pipes instead of the MUFU.
#ifndef FA_EMU_PAIRS
#define FA_EMU_PAIRS 1
#endif
constexpr int kEmuPairs = FA_EMU_PAIRS;
// Turn-taking between the softmax groups of the two Q tiles (experiment): tile 1 starts its step on KV
// tile j only after tile 0 has issued its last exponentials of step j, and tile 0 starts step j+1
// only after tile 1's step j, so the two groups never compete for the MUFU.
// Peel the first KV tile of every pass into its own instantiation of the softmax step (kFirst, see ws_softmax_step).
#ifndef FA_PEEL_FIRST
#define FA_PEEL_FIRST 1
#endif
// Interior KV tiles (not the first, not the last of a pass) run a softmax-step instantiation without mask code.
#ifndef FA_PEEL_MASK
#define FA_PEEL_MASK 1
#endif
// MMA thread: condition-free steady-state loop when both Q tiles visit every KV tile (experiment switch).
#ifndef FA_MMA_FULL_LOOP
#define FA_MMA_FULL_LOOP 1
#endif
#ifndef FA_MAX_XCHG_SHARED
#define FA_MAX_XCHG_SHARED 1
#endif
#ifndef FA_SEQ
#define FA_SEQ 0
#endif
constexpr bool kSeq = FA_SEQ != 0;
// P hand-off in 3 parts (32 + 16 + 16 keys per half; default) or 2 (32 + 32).  Measured: 3 parts are
// +5.5 % at N=16384 and +7.7 % at N=4096 (the exposed tail of O += P V shrinks to two k-steps)
#ifndef FA_EXP_SKIP_LOADS
#define FA_EXP_SKIP_LOADS 0
#endif
#ifndef FA_PV_PARTS
#define FA_PV_PARTS 3
#endif
constexpr int kPvParts = FA_PV_PARTS;

// One softmax step of one thread: `s` holds its 64 raw scores of the current S tile (already loaded
// from TMEM).  Masks them, exponentiates against the running max (first half speculatively against
// the max of the previous tiles, see below), hands P to the MMA warp in two parts, rescales O when
// the lazy-rescale rule fires and updates (m_run, l_run).  Shared by the one-shot and the
// persistent kernel.
//   tS / tO     TMEM addresses of my 64 S columns (P goes over the first 32) / my O columns
//   col0        index of the first key of my half;  diag: this is the causal diagonal tile
//   have_o      O_t already holds a partial sum (not the first KV tile of this pass)
//   bar_o       0, or the mbarrier (with parity o_parity) that tells PV(j-1) has left the tensor cores
//   kPairArrive the P barriers are shared::cluster addresses in the leader CTA of a CTA pair (wide2 kernel)
//   kFirst      compile-time "first KV tile of a pass" (the kernels peel that step, FA_PEEL_FIRST): there is no running max to
//               speculate against, so the row max is reduced first and every exponential is computed once, against the real
//               max - no second tensor-memory load, no O to rescale.  have_o must be false.
//   kNoMask     compile-time "this tile is neither the last KV tile (ragged tail) nor the causal diagonal": the kernels run
//               the interior tiles of a pass through an instantiation without any mask code (FA_PEEL_MASK)
template <int kDP, bool kBF16, bool kPairArrive = false, bool kFirst = false, bool kNoMask = false>
__device__ __forceinline__ void ws_softmax_step(float (&s)[64], uint32_t tS, uint32_t tO, int half,
                                                int r, int lane, int col0, int Nkv, bool diag,
                                                float c, float& m_run, float& l_run, bool have_o,
                                                float* my_max, const float* other_max, int pair_bar,
                                                uint32_t bar_early, uint32_t bar_late,
                                                uint32_t bar_turn = 0u, uint32_t bar_mid = 0u,
                                                uint32_t bar_o = 0u, uint32_t o_parity = 0u FA_TRS_PARAM) {
  constexpr int kOHalf = kDP / 2;
  const bool tail = !kNoMask && (col0 + 64 > Nkv);
  const bool masked = !kNoMask && (tail || diag);
  int lim = 64;  // columns [0, lim) of my half are visible
  if (masked) {
    const int valid = tail ? (Nkv - col0) : 64;
    lim = diag ? min(valid, r + 1 - half * 64) : valid;
#pragma unroll
    for (int i = 0; i < 64; ++i)
      if (i >= lim) s[i] = -INFINITY;
  }

  auto arrive = [](uint32_t bar) {
    if constexpr (kPairArrive) mbar_arrive_cluster(bar); else mbar_arrive(bar);
  };
  // p = 2^(s*c - m*c) for one group of 4 columns: kEmuPairs of every 8 element pairs go through
  // the FMA pipes (ex2_fma2), the rest through the MUFU
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

  // Columns [0,32) are exponentiated against the running max of the PREVIOUS tiles while the
  // max of this tile is still being reduced (the MUFU is the busiest pipe: the max, the pair
  // exchange and the packing run in its shadow).  That is exact whenever the lazy-rescale rule
  // keeps m_run anyway (max grew by < 2^8); otherwise the slow path below redoes the columns.
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
  FA_TRS(0);
  // (ptxas schedules the MUFU instructions of the loop above BEHIND this barrier - SASS-checked in round 2 - so the
  // chain is max -> exchange -> exponentials; forcing them in front of it measured 4.5-8 % slower)
#if FA_MAX_XCHG_SHARED
  st_shared_f32(smem_u32(my_max), mx);
  named_bar_sync(pair_bar, 64);
  FA_TRS(1);
  const float m_cand = fmaxf(fmaxf(mx, ld_shared_f32(smem_u32(other_max))), m_run);
#else
  *my_max = mx;
  named_bar_sync(pair_bar, 64);
  FA_TRS(1);
  const float m_cand = fmaxf(fmaxf(mx, *other_max), m_run);
#endif
  // both threads of the row see the same three numbers, so they take the same decision
  const bool grow = (m_cand - m_run) * c > kRescaleThreshold;  // always true on the first tile
  float alpha = 1.f;
  if constexpr (kFirst) {
    // (l_run is 0 and O_t empty: alpha is never used; s[] still holds the raw, masked scores)
    if (grow) m_run = m_cand;
    nmc = -m_run * c;
#pragma unroll
    for (int i = 0; i < 32; i += 4) exp4(i, nmc);
  } else if (__any_sync(0xffffffffu, grow)) {
    // slow path (first tile; afterwards only when some row max grew by more than 2^8)
    if (grow) {
      alpha = ex2_approx((m_run - m_cand) * c);
      m_run = m_cand;
    }
    if (have_o) {
      // O_t *= alpha on my half of the row.  PV_t(j-1) has completed - it was issued before
      // S_t(j), whose commit we waited for - and PV_t(j) waits for bar_p_early.  (The wide kernel
      // issues S(j) ahead of PV(j-1) and passes the barrier PV(j-1) commits to instead.)
      if (bar_o != 0u) {
        mbar_wait(bar_o, o_parity, 44);
        tc_fence_after();
      }
#pragma unroll 1
      for (int c8 = 0; c8 < kOHalf; c8 += 8) {  // 8 columns at a time: keeps s[] in registers
        uint32_t o[8];
        tmem_ld_x8(tO + c8, o);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st_x8(tO + c8, o);
      }
    }
    // redo columns [0,32) against the new max: S is still intact in TMEM (no P stored yet)
    nmc = -m_run * c;
    tmem_ld_x32(tS, reinterpret_cast<uint32_t*>(s));
    tmem_wait_ld();
    if (masked) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i >= lim) s[i] = -INFINITY;
    }
#pragma unroll
    for (int i = 0; i < 32; i += 4) exp4(i, nmc);
  }

  // ---- first half of my P columns -> TMEM -> "early" hand-off
  {
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = pack2<kBF16>(s[2 * i], s[2 * i + 1]);
    tmem_st_x16(tS, pk);
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) arrive(bar_early);
    FA_TRS(2);
  }
  // ---- second half, with the row sum of the first half in the MUFU shadow
  float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
  {
    uint32_t pk[16];
#pragma unroll
    for (int i = 32; i < 64; i += 4) {
      exp4(i, nmc);
      fadd2(sum0, sum1, sum0, sum1, s[i - 32], s[i - 31]);
      fadd2(sum2, sum3, sum2, sum3, s[i - 30], s[i - 29]);
      pk[(i - 32) >> 1] = pack2<kBF16>(s[i], s[i + 1]);
      pk[((i - 32) >> 1) + 1] = pack2<kBF16>(s[i + 2], s[i + 3]);
      if (kPvParts == 3 && i == 44) {  // FA_PV_PARTS=3: columns [32,48) leave as a "mid" part
        uint32_t lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) lo[e] = pk[e];
        tmem_st_x8(tS + 16, lo);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive(bar_mid);
        FA_TRS(3);
      }
    }
    if (kPvParts == 3) {
      uint32_t hi[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) hi[e] = pk[8 + e];
      tmem_st_x8(tS + 24, hi);
    } else {
      tmem_st_x16(tS + 16, pk);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      arrive(bar_late);
      if (bar_turn != 0u) mbar_arrive(bar_turn);  // FA_SEQ: the other tile's softmax may start
    }
    FA_TRS(4);
  }
#pragma unroll
  for (int i = 32; i < 64; i += 4) {
    fadd2(sum0, sum1, sum0, sum1, s[i], s[i + 1]);
    fadd2(sum2, sum3, sum2, sum3, s[i + 2], s[i + 3]);
  }
  l_run = l_run * alpha + ((sum0 + sum1) + (sum2 + sum3));
}

// Epilogue building block shared by every two-threads-per-row kernel: my half of the normalised output row,
// O / l, from tensor memory -> 16 bit -> the 128-byte-swizzled staging tile a TMA store reads.
template <int kOHalf, bool kBF16, bool kUnroll = true>
__device__ __forceinline__ void o_row_half_to_stage(uint32_t tO, uint8_t* stage, int r, int half, float inv_l) {
  auto chunk = [&](int cidx) {
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
  };
  if constexpr (kUnroll) {
#pragma unroll
    for (int cidx = 0; cidx < kOHalf / 32; ++cidx) chunk(cidx);
  } else {  // head dims 192 / 256: keep the code small
#pragma unroll 1
    for (int cidx = 0; cidx < kOHalf / 32; ++cidx) chunk(cidx);
  }
}

template <int kDP, bool kBF16, bool kCausal>
__global__ void __launch_bounds__(kWsThreads, 1)
fa_fwd_ws_kernel(const __grid_constant__ CUtensorMap tmap_q,
                 const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v,
                 const __grid_constant__ CUtensorMap tmap_o, const TcParams p) {
  using C = WsCfg<kDP>;
  constexpr int kS = C::kStages;
  constexpr int kDBlocks = kDP / 64;
  constexpr int kKSteps = kDP / 16;
  constexpr int kOHalf = kDP / 2;  // O columns each of the two threads of a row owns
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

  // barrier map
  auto bar_q_full = [&](int t) { return smem_u32(&bars[t]); };          // tx, count 1
  auto bar_s_full = [&](int t) { return smem_u32(&bars[2 + t]); };      // tcgen05.commit
  // P goes to the MMA warp in two parts (one arrival per softmax warp, 8 warps per tile):
  // "early" = first 32 keys of each half stored AND O rescaled, "late" = the other 32 of each half.
  auto bar_p_early = [&](int t) { return smem_u32(&bars[4 + t]); };
  auto bar_p_late = [&](int t) { return smem_u32(&bars[6 + t]); };
  auto bar_o_final = [&](int t) { return smem_u32(&bars[8 + t]); };     // tcgen05.commit
  auto bar_turn = [&](int t) { return smem_u32(&bars[10 + t]); };       // kSeq: 8 warps of the other tile
  auto bar_p_mid = [&](int t) { return smem_u32(&bars[12 + t]); };      // kPvParts == 3: 8 warps
  auto bar_kv_full = [&](int s) { return smem_u32(&bars[14 + s]); };    // tx, count 1
  auto bar_kv_empty = [&](int s) { return smem_u32(&bars[14 + kS + s]); };  // tcgen05.commit

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  int pair, h, b;  // causal: longest blocks first across the whole launch (work_coords)
  work_coords<kCausal>((p.Nq + 2 * kTileM - 1) / (2 * kTileM), p.H, 1, pair, h, b);
  const int row0 = pair * 2 * kTileM;
#ifdef FA_TRACE
  const bool tr_cta = p.trace != nullptr && pair == (p.Nq + 2 * kTileM - 1) / (2 * kTileM) / 2 && h == 0 && b == 0;
  const bool tr_on = tr_cta && (warp >= 16 || (warp & 7) == 0);  // (row 3 of the trace: fine stamps of the MMA thread)
#endif

  // per-tile KV trip counts
  const int n_kv_total = (p.Nkv + kTileN - 1) / kTileN;
  int n_t[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int r0 = row0 + t * kTileM;
    int n = (r0 < p.Nq) ? n_kv_total : 0;
    if (kCausal) n = min(n, r0 / kTileN + 1);
    n_t[t] = n;
  }
  const int n_max = max(n_t[0], n_t[1]);

  if (warp == 16 && lane == 0) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_q_full(t), 1);
      mbar_init(bar_s_full(t), 1);
      mbar_init(bar_p_early(t), 8);
      mbar_init(bar_p_late(t), 8);
      mbar_init(bar_o_final(t), 1);
      mbar_init(bar_turn(t), 8);
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
    // the Q tiles and the first ring-full of K/V -> L2, before pdl_wait(): hides the HBM latency of the first
    // loads under the previous kernel's tail (ptx.cuh: L2 is coherent, the loads themselves come after the wait)
#pragma unroll
    for (int db = 0; db < kDBlocks; ++db) {
#pragma unroll
      for (int t = 0; t < 2; ++t)
        if (n_t[t] > 0) tma_prefetch_l2_4d(&tmap_q, db * 64, row0 + t * kTileM, h, b);
#pragma unroll
      for (int j = 0; j < kS / 2; ++j) {
        if (j < n_max) {
          tma_prefetch_l2_4d(&tmap_k, db * 64, j * kTileN, h, b);
          tma_prefetch_l2_4d(&tmap_v, db * 64, j * kTileN, h, b);
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
  // The CTA owns all 512 TMEM columns (one CTA per SM), so the allocation starts at lane 0,
  // column 0.  Using the literal 0 keeps every TMEM address a compile-time constant; otherwise
  // ptxas cannot prove the value read back from shared memory warp-uniform and wraps each
  // tcgen05.mma in an ELECT / R2UR.BROADCAST loop that made the issuing thread the bottleneck.
  if (*tmem_slot != 0u) __trap();
  constexpr uint32_t tmem = 0u;
  const float c = p.scale_log2;
#ifdef FA_TRACE
  if (tr_cta && tid == 0) {  // SM clock during the kernel = d(clock64) / d(globaltimer)
    p.trace[(4 * 128) * 8 + 4] = clock64();
    p.trace[(4 * 128) * 8 + 5] = globaltimer_ns();
  }
#endif

  if (warp >= 16) {
    // =========================================================================================
    // warpgroup 4: MMA issuer (warp 16), TMA producer (warp 17)
    // =========================================================================================
    setmaxnreg_dec<32>();
    if (warp == 17) {
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
#pragma unroll 1
        for (int idx = 0; idx < 2 * n_max; ++idx) {  // ring order K0 V0 K1 V1 ...
          const int slot = idx % kS;
          const uint32_t use = idx / kS;
          mbar_wait(bar_kv_empty(slot), (use & 1) ^ 1, 20);
          FA_TR(4, idx >> 1, idx & 1);
#if FA_EXP_SKIP_LOADS
          // timing experiment only (wrong results): after the ring has been filled once, FA_EXP_SKIP_LOADS=1
          // stops loading V tiles, =2 stops loading K and V tiles - how sensitive is the step to TMA /
          // shared-memory-port traffic?
          if (idx >= kS && (FA_EXP_SKIP_LOADS == 2 || (idx & 1))) {
            mbar_arrive(bar_kv_full(slot));
            continue;
          }
#endif
          mbar_arrive_expect_tx(bar_kv_full(slot), C::kTileBytes);
          const CUtensorMap* map = (idx & 1) ? &tmap_v : &tmap_k;
#pragma unroll
          for (int db = 0; db < kDBlocks; ++db)
            tma_load_4d(sKV + slot * C::kTileBytes + db * 16384, map, bar_kv_full(slot), db * 64,
                        (idx >> 1) * kTileN, h, b);
        }
      }
      __syncwarp();
    } else if (warp == 16) {
      if (elect_one()) {
        constexpr uint32_t idesc_s = make_idesc_f16(kTileM, kTileN, kBF16, false, false);
        constexpr uint32_t idesc_o = make_idesc_f16(kTileM, kDP, kBF16, false, true);

        auto wait_kv = [&](int idx) {
          mbar_wait(bar_kv_full(idx % kS), (idx / kS) & 1, 30);
          tc_fence_after();
        };
        auto release_kv = [&](int idx) { tc_commit(bar_kv_empty(idx % kS)); };
        constexpr uint32_t desc_hi = smem_desc_hi_sw128(1024);
        auto issue_s = [&](int t, int j) {  // S_t = Q_t K_j^T
          const uint32_t k_lo = smem_desc_lo(sKV + ((2 * j) % kS) * C::kTileBytes, 16);
          const uint32_t q_lo = smem_desc_lo(sQ + t * C::kTileBytes, 16);
#pragma unroll
          for (int k = 0; k < kKSteps; ++k) {
            const uint32_t off = ((k >> 2) * 16384 + (k & 3) * 32) >> 4;
            umma_ss2(tmem + col_s(t), q_lo + off, desc_hi, k_lo + off, desc_hi, idesc_s, k > 0);
          }
          if (t == 1) FA_TR(3, j - 1, 0);
          tc_commit(bar_s_full(t));
          if (t == 1) FA_TR(3, j - 1, 1);
        };
        // k-step ks covers keys [16 ks, 16 ks + 16): P columns of half ks/4 at S column
        // 64 (ks/4) + 8 (ks%4); V rows 16 ks of the tile (2048 bytes apart in the MN-major tile)
        auto pv_step = [&](int t, uint32_t v_lo, int ks, uint32_t acc) {
          umma_ts2(tmem + col_o(t), tmem + col_s(t) + (ks >> 2) * 64 + (ks & 3) * 8,
                   v_lo + ((ks * 2048) >> 4), desc_hi, idesc_o, acc);
        };
        auto issue_pv = [&](int t, int j) {  // O_t += P_t V_j
          const uint32_t v_lo = smem_desc_lo(sKV + ((2 * j + 1) % kS) * C::kTileBytes, 16384);
          mbar_wait(bar_p_early(t), j & 1, 31 + t);
          tc_fence_after();
          FA_TR(2, j, 2 + 3 * t);
          pv_step(t, v_lo, 0, j > 0);
          pv_step(t, v_lo, 1, 1);
          pv_step(t, v_lo, 4, 1);
          pv_step(t, v_lo, 5, 1);
          if (kPvParts == 3) {
            mbar_wait(bar_p_mid(t), j & 1, 37 + t);
            tc_fence_after();
            pv_step(t, v_lo, 2, 1);
            pv_step(t, v_lo, 6, 1);
            mbar_wait(bar_p_late(t), j & 1, 35 + t);
            tc_fence_after();
            pv_step(t, v_lo, 3, 1);
            pv_step(t, v_lo, 7, 1);
          } else {
            mbar_wait(bar_p_late(t), j & 1, 35 + t);
            tc_fence_after();
            pv_step(t, v_lo, 2, 1);
            pv_step(t, v_lo, 3, 1);
            pv_step(t, v_lo, 6, 1);
            pv_step(t, v_lo, 7, 1);
          }
          if (j == n_t[t] - 1) tc_commit(bar_o_final(t));
        };

        if (n_max > 0) {
          wait_kv(0);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (n_t[t] > 0) {
              mbar_wait(bar_q_full(t), 0, 33);
              tc_fence_after();
              issue_s(t, 0);
            }
          }
          release_kv(0);
        }
#if FA_MMA_FULL_LOOP
        // While both Q tiles have a further KV tile to visit (all iterations but the last when non-causal; under a causal
        // mask up to the first tile's diagonal): no per-tile conditions - the generic loop below finishes the rest (the
        // issuing thread sits on the P -> PV -> S chain of both tiles).
        // (non-causal: only when both tiles visit every KV tile - the bound written as n_max behind a `full` test; the general
        // form costs that instantiation 1 %, ptxas again.  Causal: up to the first tile's diagonal, +0.3...0.9 %.)
        const bool full = !kCausal && n_t[0] == n_max && n_t[1] == n_max;
        const int n_full = kCausal ? min(n_t[0], n_t[1]) : n_max;
        int j0 = 0;
        if (kCausal || full) {
#pragma unroll 1
          for (; j0 < n_full - 1; ++j0) {
            const int j = j0, nx = j0 + 1;
            FA_TR(2, j, 0);
            wait_kv(2 * j + 1);
            FA_TR(2, j, 1);
            issue_pv(0, j);
            FA_TR(2, j, 3);
            wait_kv(2 * nx);
            issue_s(0, nx);
            FA_TR(2, j, 4);
            issue_pv(1, j);
            FA_TR(2, j, 6);
            FA_TR(3, j, 3);
            release_kv(2 * j + 1);
            FA_TR(3, j, 4);
            issue_s(1, nx);
            FA_TR(3, j, 2);
            release_kv(2 * nx);
            FA_TR(2, j, 7);
          }
        }
#else
        const int j0 = 0;
#endif
#pragma unroll 1
        for (int j = j0; j < n_max; ++j) {
          const int nx = j + 1;
          FA_TR(2, j, 0);
          wait_kv(2 * j + 1);
          FA_TR(2, j, 1);
          if (j < n_t[0]) issue_pv(0, j);
          FA_TR(2, j, 3);
          if (nx < n_max) wait_kv(2 * nx);
          if (nx < n_t[0]) issue_s(0, nx);
          FA_TR(2, j, 4);
          if (j < n_t[1]) issue_pv(1, j);
          FA_TR(2, j, 6);
          FA_TR(3, j, 3);
          release_kv(2 * j + 1);
          FA_TR(3, j, 4);
          if (nx < n_t[1]) issue_s(1, nx);
          FA_TR(3, j, 2);
          if (nx < n_max) release_kv(2 * nx);
          FA_TR(2, j, 7);
        }
      }
      __syncwarp();
    }
  } else {
    // =========================================================================================
    // softmax warps (0-7: tile 0, 8-15: tile 1)
    // =========================================================================================
    setmaxnreg_inc<112>();  // 512 x 112 + 128 x 32 = 640 x 96: setmaxnreg only redistributes the CTA's launch allocation
    const int t = warp >> 3;
    const int half = (warp >> 2) & 1;
#ifdef FA_TRACE
    const int tr_role = (warp == 4) ? 3 : t;
#endif
    const int r = (warp & 3) * 32 + lane;  // query row inside the tile = TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_base + col_s(t) + half * 64;  // my 64 S columns; P over [0,32)
    const uint32_t tO = tmem + lane_base + col_o(t) + half * kOHalf;
    const int pair_bar = 1 + t * 4 + (warp & 3);  // named barrier of the two warps sharing my rows
    const int tile_row0 = row0 + t * kTileM;
    const int diag_j = tile_row0 / kTileN;
    const int n = n_t[t];
    float* my_max = sMax + (t * 2 + half) * 128 + r;
    const float* other_max = sMax + (t * 2 + (half ^ 1)) * 128 + r;

    float m_run = -INFINITY;
    float l_run = 0.f;  // partial row sum over my key half

    auto kv_step = [&](int j, auto first_tag, auto nomask_tag) {
      constexpr bool kFirstStep = decltype(first_tag)::value;
      constexpr bool kNoMaskStep = decltype(nomask_tag)::value;
      FA_TR(tr_role, j, 0);
      mbar_wait_warp(bar_s_full(t), j & 1, 40 + t);
      tc_fence_after();
      FA_TR(tr_role, j, 1);
      float s[64];
      tmem_ld_x32(tS, reinterpret_cast<uint32_t*>(s));
      tmem_ld_x32(tS + 32, reinterpret_cast<uint32_t*>(s) + 32);
      tmem_wait_ld();
      if (kSeq) {  // my turn?  tile 0 goes first; step j of tile 1 follows step j of tile 0
        if (t == 0) {
          if (j > 0 && j - 1 < n_t[1]) mbar_wait(bar_turn(0), (j - 1) & 1, 42);
        } else {
          if (j < n_t[0]) mbar_wait(bar_turn(1), j & 1, 43);
        }
      }
      FA_TR(tr_role, j, 2);

      ws_softmax_step<kDP, kBF16, false, kFirstStep, kNoMaskStep>(s, tS, tO, half, r, lane, j * kTileN + half * 64, p.Nkv,
                                  kCausal && (j == diag_j), c, m_run, l_run, FA_PEEL_FIRST ? !kFirstStep : (j > 0),
                                  my_max + (j & 1) * 512, other_max + (j & 1) * 512, pair_bar,
                                  bar_p_early(t), bar_p_late(t), kSeq ? bar_turn(t ^ 1) : 0u,
                                  bar_p_mid(t)
#ifdef FA_TRACE
                                  , 0u, 0u, (tr_on && j < 128) ? p.trace + ((5 + t) * 128 + j) * 8 : nullptr
#endif
                                  );
      FA_TR(tr_role, j, 6);
    };
    // (the causal diagonal and the ragged tail can only be a pass's last tile: n = min(all tiles, diagonal + 1))
#if FA_PEEL_FIRST && FA_PEEL_MASK
    if (n > 0) kv_step(0, std::true_type{}, std::false_type{});
#pragma unroll 1
    for (int j = 1; j < n - 1; ++j) kv_step(j, std::false_type{}, std::true_type{});
    if (n > 1) kv_step(n - 1, std::false_type{}, std::false_type{});
#elif FA_PEEL_FIRST
    if (n > 0) kv_step(0, std::true_type{}, std::false_type{});
#pragma unroll 1
    for (int j = 1; j < n; ++j) kv_step(j, std::false_type{}, std::false_type{});
#else
#pragma unroll 1
    for (int j = 0; j < n; ++j) kv_step(j, std::false_type{}, std::false_type{});
#endif

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
  }

  tc_fence_before();
  __syncthreads();
#ifdef FA_TRACE
  if (tr_cta && tid == 0) {
    p.trace[(4 * 128) * 8 + 6] = clock64();
    p.trace[(4 * 128) * 8 + 7] = globaltimer_ns();
  }
#endif
  if (warp == 16) tmem_dealloc(tmem, 512);
}

}  // namespace fa
