// Single-tile tcgen05 forward kernel ("tc1"): one CTA = one (batch, head, 128-row Q tile),
// one warpgroup doing every role in sequence.  K/V tiles are double-buffered by TMA, S = Q K^T and
// O += P V run on the tensor cores with accumulators in TMEM, the online softmax runs in registers
// with one thread per query row (no shuffles: a TMEM lane is a row).
//
// This is the simple, serial version: tensor cores idle during the softmax.  It is used for short
// query lengths (Nq <= 128, where the two-tile kernel would waste half its rows) and is the
// stepping stone the warp-specialised kernel (fa_fwd_ws.cuh) was validated against.
//
// Replaces /root/reference/rocwmma_fattn/kernel_fp16.cu:306-544 (fwd_kernel) and its device GEMM
// helpers mul_A_BT (:115-175) / mul_add_A_B (:178-232) / mul_add_A_B_mask_k (:235-302).
#pragma once
#include "ptx.cuh"
#include "umma_probe.cuh"  // sw128_offset_16bit

namespace fa {

struct TcParams {
  float* lse;            // [B,H,Nq] fp32, base-2 log-sum-exp of the scaled scores; may be null
  int Nq, Nkv, H;
  float scale_log2;      // scale * log2(e)
  // persistent stream-K kernel only (fa_fwd_sk.cuh); zero / null otherwise
  float* sk_ws;          // per-CTA partial results of split units
  int* sk_flags;         // [gridDim.x][2] "partial of tile t is ready" flags, all zero between launches
  long long sk_W;        // work items of the stream-K phase = (units - sk_dp * grid) * sk_T
  int sk_dp;             // leading data-parallel rounds (one whole unit per CTA and round)
  int sk_T;              // KV tiles per unit
  int sk_P;              // 256-row query blocks per (batch, head)
#ifdef FA_TRACE
  unsigned long long* trace;  // debug builds only: clock64() stamps of one CTA (tools/trace_ws.py)
#endif
};

// Work item of a CTA (or CTA pair): index `blk` of its row block inside a (batch, head), and (h, b).
//   non-causal  3-D grid (blocks per head, H, B), as launched
//   causal      1-D grid ordered longest block first ACROSS heads: linear index L -> block rank L / (B H),
//               head L % (B H).  A CTA's work grows with its block index (it visits the KV tiles up to its
//               diagonal), and the hardware dispatches CTAs in index order: with the 3-D grid every head's long
//               blocks queue up behind the short blocks of the heads before it and the launch ends on a few long
//               stragglers (fp16 H=16 D=128 N=4096: 90 us for 50 us of work per SM); longest-first over the whole
//               launch is the classic LPT list schedule.  `unit` = 1 for one CTA per block, 2 for CTA pairs
//               (gridDim.x counts CTAs).
template <bool kCausal>
__device__ __forceinline__ void work_coords(int blocks_per_head, int H, int unit, int& blk, int& h, int& b) {
  if constexpr (kCausal) {
    const int L = static_cast<int>(blockIdx.x) / unit;
    const int n_bh = (static_cast<int>(gridDim.x) / unit) / blocks_per_head;
    const int bh = L % n_bh;
    blk = blocks_per_head - 1 - L / n_bh;
    h = bh % H;
    b = bh / H;
  } else {
    blk = static_cast<int>(blockIdx.x) / unit;
    h = blockIdx.y;
    b = blockIdx.z;
  }
}

constexpr int kTileM = 128;  // query rows per tile
constexpr int kTileN = 128;  // keys per tile

template <int kDP>
struct Tc1Smem {
  static constexpr int kTileBytes = kTileM * kDP * 2;  // one Q / K / V tile
  static constexpr int kQ = 0;
  static constexpr int kK = kQ + kTileBytes;           // 2 stages
  static constexpr int kV = kK + 2 * kTileBytes;       // 2 stages
  static constexpr int kBars = kV + 2 * kTileBytes;
  static constexpr int kTotal = kBars + 128 + 1024;    // + alignment slack
};

template <int kDP, bool kBF16, bool kCausal>
__global__ void __launch_bounds__(128, 1)
fa_fwd_tc1_kernel(const __grid_constant__ CUtensorMap tmap_q,
                  const __grid_constant__ CUtensorMap tmap_k,
                  const __grid_constant__ CUtensorMap tmap_v,
                  const __grid_constant__ CUtensorMap tmap_o, const TcParams p) {
  using L = Tc1Smem<kDP>;
  constexpr int kDBlocks = kDP / 64;          // 64-element (128-byte) swizzle blocks per row
  constexpr int kKSteps = kDP / 16;           // UMMA K steps for S = Q K^T
  constexpr uint32_t kColS = 0, kColP = 128, kColO = 256;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t sQ = smem_u32(smem + L::kQ);
  const uint32_t sK = smem_u32(smem + L::kK);
  const uint32_t sV = smem_u32(smem + L::kV);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kBars + 64);
  const uint32_t bar_q = smem_u32(&bars[0]);
  const uint32_t bar_kv[2] = {smem_u32(&bars[1]), smem_u32(&bars[2])};
  const uint32_t bar_mma = smem_u32(&bars[3]);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  int qtile, h, b;  // causal: longest tiles first across the whole launch (work_coords)
  work_coords<kCausal>((p.Nq + kTileM - 1) / kTileM, p.H, 1, qtile, h, b);
  const int row0 = qtile * kTileM;

  if (tid == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_o);
#pragma unroll
    for (int db = 0; db < kDBlocks; ++db) {  // first tiles -> L2 before pdl_wait() (see fa_fwd_ws.cuh)
      tma_prefetch_l2_4d(&tmap_q, db * 64, row0, h, b);
      tma_prefetch_l2_4d(&tmap_k, db * 64, 0, h, b);
      tma_prefetch_l2_4d(&tmap_v, db * 64, 0, h, b);
    }
    mbar_init(bar_q, 1);
    mbar_init(bar_kv[0], 1);
    mbar_init(bar_kv[1], 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // PDL: everything above overlapped the previous kernel's tail; global memory is touched only below
  pdl_wait();
  pdl_launch_dependents();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;

  int n_kv = (p.Nkv + kTileN - 1) / kTileN;
  if (kCausal) n_kv = min(n_kv, qtile + 1);

  auto load_kv = [&](int j, int stage) {
    mbar_arrive_expect_tx(bar_kv[stage], 2 * L::kTileBytes);
#pragma unroll
    for (int db = 0; db < kDBlocks; ++db) {
      tma_load_4d(sK + stage * L::kTileBytes + db * 16384, &tmap_k, bar_kv[stage], db * 64,
                  j * kTileN, h, b);
      tma_load_4d(sV + stage * L::kTileBytes + db * 16384, &tmap_v, bar_kv[stage], db * 64,
                  j * kTileN, h, b);
    }
  };

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, L::kTileBytes);
#pragma unroll
    for (int db = 0; db < kDBlocks; ++db)
      tma_load_4d(sQ + db * 16384, &tmap_q, bar_q, db * 64, row0, h, b);
    load_kv(0, 0);
  }

  constexpr uint32_t idesc_s = make_idesc_f16(kTileM, kTileN, kBF16, false, false);
  constexpr uint32_t idesc_o = make_idesc_f16(kTileM, kDP, kBF16, false, true);

  float m_run = -INFINITY;  // running max of the raw (unscaled) scores
  float l_run = 0.f;
  uint32_t mma_phase = 0;
  const float c = p.scale_log2;

#pragma unroll 1
  for (int j = 0; j < n_kv; ++j) {
    const int stage = j & 1;
    if (tid == 0) {
      if (j + 1 < n_kv) load_kv(j + 1, stage ^ 1);
      if (j == 0) mbar_wait(bar_q, 0, 10);
      mbar_wait(bar_kv[stage], (j >> 1) & 1, 11);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < kKSteps; ++k) {
        const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
        umma_ss(tmem + kColS, make_smem_desc_sw128(sQ + off, 16, 1024),
                make_smem_desc_sw128(sK + stage * L::kTileBytes + off, 16, 1024), idesc_s, k > 0);
      }
      tc_commit(bar_mma);
    }
    mbar_wait(bar_mma, mma_phase, 12);
    mma_phase ^= 1;
    tc_fence_after();

    // ---- S row -> registers
    float s[kTileN];
#pragma unroll
    for (int cidx = 0; cidx < 4; ++cidx)
      tmem_ld_x32(tmem + lane_base + kColS + cidx * 32, reinterpret_cast<uint32_t*>(s) + cidx * 32);
    tmem_wait_ld();

    // ---- masks (KV tail, causal diagonal)
    const int col0 = j * kTileN;
    if (col0 + kTileN > p.Nkv) {
      const int valid = p.Nkv - col0;
#pragma unroll
      for (int i = 0; i < kTileN; ++i)
        if (i >= valid) s[i] = -INFINITY;
    }
    if (kCausal && j == qtile) {
#pragma unroll
      for (int i = 0; i < kTileN; ++i)
        if (i > tid) s[i] = -INFINITY;
    }

    // ---- online softmax (base 2)
    float mx = s[0];
#pragma unroll
    for (int i = 1; i < kTileN; ++i) mx = fmaxf(mx, s[i]);
    float m_new = fmaxf(m_run, mx);
    if (m_new == -INFINITY) m_new = 0.f;
    const float alpha = ex2_approx((m_run - m_new) * c);
    const float mc = m_new * c;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kTileN; ++i) {
      s[i] = ex2_approx(fmaf(s[i], c, -mc));
      sum += s[i];
    }
    l_run = l_run * alpha + sum;
    m_run = m_new;

    // ---- O *= alpha  (previous PV has completed: we waited on its commit)
    if (j > 0) {
#pragma unroll
      for (int cidx = 0; cidx < kDP / 32; ++cidx) {
        uint32_t r[32];
        tmem_ld_x32(tmem + lane_base + kColO + cidx * 32, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
        tmem_st_x32(tmem + lane_base + kColO + cidx * 32, r);
      }
    }

    // ---- P (16-bit) -> TMEM
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      uint32_t r[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = pack2<kBF16>(s[hlf * 64 + 2 * i], s[hlf * 64 + 2 * i + 1]);
      tmem_st_x32(tmem + lane_base + kColP + hlf * 32, r);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();

    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < kTileN / 16; ++k) {
        const uint64_t b_desc =
            make_smem_desc_sw128(sV + stage * L::kTileBytes + k * 2048, 16384, 1024);
        umma_ts(tmem + kColO, tmem + kColP + k * 8, b_desc, idesc_o, (j > 0) || (k > 0));
      }
      tc_commit(bar_mma);
    }
    mbar_wait(bar_mma, mma_phase, 13);
    mma_phase ^= 1;
    tc_fence_after();
  }

  // ---- epilogue: O / l -> 16-bit -> smem (Q buffer, swizzled) -> TMA store
  const float inv_l = 1.f / l_run;
#pragma unroll
  for (int cidx = 0; cidx < kDP / 32; ++cidx) {
    uint32_t r[32];
    tmem_ld_x32(tmem + lane_base + kColO + cidx * 32, r);
    tmem_wait_ld();
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      uint4 val;
      val.x = pack2<kBF16>(__uint_as_float(r[ch * 8 + 0]) * inv_l, __uint_as_float(r[ch * 8 + 1]) * inv_l);
      val.y = pack2<kBF16>(__uint_as_float(r[ch * 8 + 2]) * inv_l, __uint_as_float(r[ch * 8 + 3]) * inv_l);
      val.z = pack2<kBF16>(__uint_as_float(r[ch * 8 + 4]) * inv_l, __uint_as_float(r[ch * 8 + 5]) * inv_l);
      val.w = pack2<kBF16>(__uint_as_float(r[ch * 8 + 6]) * inv_l, __uint_as_float(r[ch * 8 + 7]) * inv_l);
      *reinterpret_cast<uint4*>(smem + L::kQ + sw128_offset_16bit(tid, cidx * 32 + ch * 8)) = val;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int db = 0; db < kDBlocks; ++db)
      tma_store_4d(&tmap_o, sQ + db * 16384, db * 64, row0, h, b);
    tma_store_commit();
  }
  const int row = row0 + tid;
  if (p.lse != nullptr && row < p.Nq) {
    p.lse[(static_cast<int64_t>(b) * p.H + h) * p.Nq + row] = m_run * c + log2f(l_run);
  }
  if (warp == 0) tmem_dealloc(tmem, 512);
  if (tid == 0) tma_store_wait_read();
}

}  // namespace fa
