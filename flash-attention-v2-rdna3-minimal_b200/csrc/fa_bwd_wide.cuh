// tcgen05 backward for head dims 129..256 ("bwd wide", round 2).
//
// fa_bwd_ws.cuh keeps S^T, dP^T, dV and dK in tensor memory at once (256 + 2 D columns) and K, V, 2 x Q, dO in shared
// memory (5 x 128 x D x 2 bytes): neither fits above head dim 128, and until this file those head dims (the 160 of
// SD 1.5 the reference serves by padding, kernel_fp16.cu:900) ran the CUDA-core kernels of fa_bwd_simt.cuh at under
// 1 TFLOP/s.  Here the three gradients are computed by THREE launches of one kernel, each of which holds a single
// 128 x D accumulator in tensor memory and recomputes the scores it needs (8 GEMMs instead of 5; no fp32 dQ buffer, no
// atomics, no conversion pass):
//
//   mode dV : CTA = 128 keys.   per 64-query tile:  S^T = K Q^T                 P^T        ->  dV += P^T dO
//   mode dK : CTA = 128 keys.   per 64-query tile:  S^T = K Q^T, dP^T = V dO^T  dS^T       ->  dK += dS^T Q
//   mode dQ : CTA = 128 queries per 64-key tile:    S = Q K^T,   dP = dO V^T    dS         ->  dQ += dS K
//   mode dKV: modes dV and dK in ONE launch where two 128 x D accumulators fit next to the scores (D <= 192: 128 + 2 x 192 = 512
//             TMEM columns): S^T, P^T and the Q / dO tiles are computed / fetched once (4 GEMMs per tile instead of 5)
//
// with P = 2^(S c - L), dS = P o (dP - D) (kernel_fp16.cu:698-737).  In every mode the CTA's own 128 rows ("resident":
// K,V or Q,dO) are TMEM lanes, the other operand streams through shared memory in 64-row tiles, the 16-bit P / dS goes
// back to tensor memory and feeds the output product as its A operand (tcgen05.mma TS form), and the streamed tile is
// the MN-major B operand of that product.  The score products are M=128, N=64, K=D; the output product M=128, N=D, K=64.
// Mode dQ is mode dK with the roles of (Q,dO) and (K,V) exchanged; its L_i, D_i are per-lane scalars, in the other two
// modes they are vectors over the 64 streamed queries (staged in shared memory, read as broadcasts).
//
// This is the serial arrangement (one thread issues TMA and MMA, 256 threads do the P / dS pass between two CTA
// barriers; the streamed tiles are double-buffered where shared memory allows): a first tensor-core version, ~500x
// the CUDA-core kernels it replaces, not yet pipelined like fa_bwd_ws.cuh.
//
// TMEM: scores [0,64), dP [64,128), accumulator [128,128+D) (mode dKV: dV there, dK at [128+D,128+2D)).  The 16-bit P / dS tile
// lives INSIDE the score columns: thread (row, half) packs its 32 columns [32 half, 32 half + 32) into the first 16 of them.
#pragma once
#include "fa_bwd_tc.cuh"

namespace fa {

constexpr int kBwdWideDV = 0, kBwdWideDK = 1, kBwdWideDQ = 2, kBwdWideDKV = 3;
constexpr int kBwdWideThreads = 256;
constexpr int kWideT = 64;  // rows of a streamed tile

template <int kDP, int kMode>
struct BwdWideSmem {
  static constexpr int kResBytes = kTileM * kDP * 2;  // one resident tile [128][kDP]
  static constexpr int kStrBytes = kWideT * kDP * 2;  // one streamed tile [64][kDP]
  static_assert(kMode != kBwdWideDKV || kDP <= 192, "two accumulators need 128 + 2 kDP <= 512 TMEM columns");
  static constexpr int kNumRes = (kMode == kBwdWideDV) ? 1 : 2;
  static constexpr int kStages = (kNumRes * kResBytes + 2 * 2 * kStrBytes <= 222 * 1024) ? 2 : 1;
  static constexpr int kRes = 0;
  static constexpr int kStr = kNumRes * kResBytes;            // [stage][2][kStrBytes]
  static constexpr int kLD = kStr + kStages * 2 * kStrBytes;  // float [2 (L, D)][64] (vector modes)
  static constexpr int kBars = kLD + 2 * 64 * 4;
  static constexpr int kTotal = kBars + 128 + 1024;           // + alignment slack
};

struct BwdWideParams {
  const float* lse;    // [B,H,Nq] base-2 log-sum-exp of the scaled scores (forward output)
  const float* delta;  // [B,H,Nq] rowsum(dO o O), fp32 (fa_bwd_delta_kernel)
  int Nq, Nkv, H;
  float scale_log2;    // scale * log2(e)
  float out_scale;     // 1 (dV) or scale (dK, dQ)
};

template <int kDP, bool kBF16, bool kCausal, int kMode>
__global__ void __launch_bounds__(kBwdWideThreads, 1)
fa_bwd_wide_kernel(const __grid_constant__ CUtensorMap tmap_res1,  // K (dV, dK) | Q (dQ), box {64, 128}
                   const __grid_constant__ CUtensorMap tmap_res2,  // V (dK) | dO (dQ); unused in mode dV
                   const __grid_constant__ CUtensorMap tmap_str1,  // Q (dV, dK) | K (dQ), box {64, 64}
                   const __grid_constant__ CUtensorMap tmap_str2,  // dO (dV, dK) | V (dQ)
                   const __grid_constant__ CUtensorMap tmap_out,   // dV | dK | dQ, box {64, 128}
                   const __grid_constant__ CUtensorMap tmap_out2,  // mode dKV: dK (tmap_out is dV); unused otherwise
                   const BwdWideParams p) {
  using L = BwdWideSmem<kDP, kMode>;
  constexpr int kS = L::kStages;
  constexpr int kDBlocks = kDP / 64;
  constexpr int kKSteps = kDP / 16;  // contraction over the head dim (scores)
  constexpr bool kVecLD = (kMode != kBwdWideDQ);  // L, D are vectors over the streamed queries
  constexpr bool kTwoOut = (kMode == kBwdWideDKV);
  constexpr bool kNeedDP = (kMode != kBwdWideDV);
  constexpr uint32_t kColS = 0, kColP = 64, kColAcc = 128, kColAcc2 = 128 + kDP;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t sRes1 = smem_u32(smem + L::kRes);
  const uint32_t sRes2 = sRes1 + L::kResBytes;  // (not present in mode dV)
  const uint32_t sStr = smem_u32(smem + L::kStr);
  float* sLD = reinterpret_cast<float*>(smem + L::kLD);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kBars + 64);
  const uint32_t bar_res = smem_u32(&bars[0]);                          // tx: resident tiles
  auto bar_str = [&](int s) { return smem_u32(&bars[1 + s]); };         // tx: streamed tile pair of a stage
  const uint32_t bar_mma1 = smem_u32(&bars[3]);                         // commit: scores ready
  const uint32_t bar_mma2 = smem_u32(&bars[4]);                         // commit: output product done

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int half = warp >> 2;             // which 32 of the 64 streamed columns this thread owns
  const int r = (warp & 3) * 32 + lane;   // row inside the resident tile = TMEM lane
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const int rt = blockIdx.x;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int row0 = rt * kTileM;
  const int64_t bh = static_cast<int64_t>(b) * p.H + h;

  // streamed-tile range of this CTA
  const int n_cols = kVecLD ? p.Nq : p.Nkv;
  const int n_ct = (n_cols + kWideT - 1) / kWideT;
  int ct0 = 0, ct1 = n_ct;
  if (kCausal) {
    if (kVecLD) ct0 = min(n_ct, row0 / kWideT);                         // queries >= the first key of the tile
    else ct1 = min(n_ct, (row0 + kTileM - 1) / kWideT + 1);             // keys <= the last query of the tile
  }
  const int n_iter = ct1 - ct0;

  if (tid == 0) {
    tma_prefetch_desc(&tmap_res1);
    tma_prefetch_desc(&tmap_res2);
    tma_prefetch_desc(&tmap_str1);
    tma_prefetch_desc(&tmap_str2);
    tma_prefetch_desc(&tmap_out);
    if (kTwoOut) tma_prefetch_desc(&tmap_out2);
#pragma unroll
    for (int i = 0; i < 5; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  auto load_str = [&](int it) {  // streamed tile pair of iteration `it` -> stage it % kS
    const int s = it % kS;
    const uint32_t dst = sStr + s * 2 * L::kStrBytes;
    mbar_arrive_expect_tx(bar_str(s), 2 * L::kStrBytes);
#pragma unroll
    for (int db = 0; db < kDBlocks; ++db) {
      tma_load_4d(dst + db * 8192, &tmap_str1, bar_str(s), db * 64, (ct0 + it) * kWideT, h, b);
      tma_load_4d(dst + L::kStrBytes + db * 8192, &tmap_str2, bar_str(s), db * 64, (ct0 + it) * kWideT, h, b);
    }
  };

  if (tid == 0 && n_iter > 0) {
    mbar_arrive_expect_tx(bar_res, L::kNumRes * L::kResBytes);
#pragma unroll
    for (int db = 0; db < kDBlocks; ++db) {
      tma_load_4d(sRes1 + db * 16384, &tmap_res1, bar_res, db * 64, row0, h, b);
      if (L::kNumRes == 2) tma_load_4d(sRes2 + db * 16384, &tmap_res2, bar_res, db * 64, row0, h, b);
    }
    load_str(0);
    if (kS == 2 && n_iter > 1) load_str(1);
  }

  constexpr uint32_t idesc_s = make_idesc_f16(kTileM, kWideT, kBF16, false, false);  // scores: M=128, N=64
  constexpr uint32_t idesc_o = make_idesc_f16(kTileM, kDP, kBF16, false, true);      // output: N=kDP, B MN-major
  const float c = p.scale_log2;
  const int row_abs = row0 + r;

  // -L / -D: per-lane scalars (mode dQ) or, one streamed tile ahead, the value this thread will stage (vector modes)
  float l_lane = 0.f, d_lane = 0.f, ld_next = 0.f;
  const float* ld_src = (tid < 64) ? p.lse : p.delta;
  if (!kVecLD) {
    if (row_abs < p.Nq) {
      l_lane = -p.lse[bh * p.Nq + row_abs];
      d_lane = -p.delta[bh * p.Nq + row_abs];
    }
  } else if (tid < 128 && n_iter > 0) {
    const int q = ct0 * kWideT + (tid & 63);
    ld_next = (q < p.Nq) ? -ld_src[bh * p.Nq + q] : 0.f;
  }

  int waited2 = 0;  // bar_mma2 phases thread 0 has observed (it must observe every phase in order)
#pragma unroll 1
  for (int it = 0; it < n_iter; ++it) {
    const int ct = ct0 + it;
    const int s = it % kS;
    if (kVecLD && tid < 128) {
      sLD[tid] = ld_next;
      if (it + 1 < n_iter) {
        const int q = (ct + 1) * kWideT + (tid & 63);
        ld_next = (q < p.Nq) ? -ld_src[bh * p.Nq + q] : 0.f;
      }
    }
    __syncthreads();  // (A) L / D of this tile visible; every thread is done with the previous tile's scores

    if (tid == 0) {
      if (it == 0) mbar_wait(bar_res, 0, 80);
      mbar_wait(bar_str(s), (it / kS) & 1, 81);
      tc_fence_after();
      const uint32_t t1 = sStr + s * 2 * L::kStrBytes, t2 = t1 + L::kStrBytes;
      // split descriptors (ptx.cuh): the high word is a constant, stepping an operand is one add on the low word
      constexpr uint32_t desc_hi = smem_desc_hi_sw128(1024);
      const uint32_t r1_lo = smem_desc_lo(sRes1, 16), t1_lo = smem_desc_lo(t1, 16);
#pragma unroll
      for (int k = 0; k < kKSteps; ++k) {
        umma_ss2(tmem + kColS, r1_lo + (((k >> 2) * 16384 + (k & 3) * 32) >> 4), desc_hi,
                 t1_lo + (((k >> 2) * 8192 + (k & 3) * 32) >> 4), desc_hi, idesc_s, k > 0);
      }
      if (kNeedDP) {
        const uint32_t r2_lo = smem_desc_lo(sRes2, 16), t2_lo = smem_desc_lo(t2, 16);
#pragma unroll
        for (int k = 0; k < kKSteps; ++k) {
          umma_ss2(tmem + kColP, r2_lo + (((k >> 2) * 16384 + (k & 3) * 32) >> 4), desc_hi,
                   t2_lo + (((k >> 2) * 8192 + (k & 3) * 32) >> 4), desc_hi, idesc_s, k > 0);
        }
      }
      tc_commit(bar_mma1);
      // two stages: the tile of iteration it+1 goes where iteration it-1's was, once its output product is done.
      // (Phase it-1 is observed here in EVERY iteration, before phase it can complete: a parity wait cannot tell
      // phase n from phase n+2 - compute-sanitizer's slow clock turned the missing wait into a deadlock.)
      if (kS == 2 && it >= 1) {
        mbar_wait(bar_mma2, waited2 & 1, 82);
        ++waited2;
        if (it + 1 < n_iter) load_str(it + 1);
      }
    }

    // ---- P (and dS) for my 32 columns of my row
    const bool need_mask = (kCausal && (kVecLD ? (ct * kWideT < row0 + kTileM) : ((ct + 1) * kWideT > row0))) ||
                           ((ct + 1) * kWideT > n_cols) || (row0 + kTileM > (kVecLD ? p.Nkv : p.Nq));
    mbar_wait(bar_mma1, it & 1, 83);
    tc_fence_after();
    {
      uint32_t sv[32], dv[32];
      tmem_ld_x32(tmem + lane_base + kColS + half * 32, sv);
      if (kNeedDP) tmem_ld_x32(tmem + lane_base + kColP + half * 32, dv);
      tmem_wait_ld();
      // P = 2^(S c + (-L)), dS = P (dP + (-D)): packed FFMA2 / FADD2 / FMUL2 (sLD, l_lane, d_lane hold the NEGATED values:
      // the packed adds take no negate modifier); the mask is a block of its own that only diagonal / edge tiles execute
      uint32_t pk[16], pk2[16];
      float pf[32];
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        const int col = half * 32 + e;
        float x0, x1;
        ffma2(x0, x1, __uint_as_float(sv[e]), __uint_as_float(sv[e + 1]), c, c, kVecLD ? sLD[col] : l_lane,
              kVecLD ? sLD[col + 1] : l_lane);
        pf[e] = ex2_approx(x0);
        pf[e + 1] = ex2_approx(x1);
      }
      if (need_mask) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int col_abs = ct * kWideT + half * 32 + e;
          const int key = kVecLD ? row_abs : col_abs;
          const int qrow = kVecLD ? col_abs : row_abs;
          const bool ok = key < p.Nkv && qrow < p.Nq && (!kCausal || key <= qrow);
          pf[e] = ok ? pf[e] : 0.f;
        }
      }
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        if (kNeedDP) {
          const int col = half * 32 + e;
          float d0, d1;
          fadd2(d0, d1, __uint_as_float(dv[e]), __uint_as_float(dv[e + 1]), kVecLD ? sLD[64 + col] : d_lane,
                kVecLD ? sLD[64 + col + 1] : d_lane);
          fmul2(d0, d1, d0, d1, pf[e], pf[e + 1]);
          if (kTwoOut) {
            pk[e >> 1] = pack2<kBF16>(pf[e], pf[e + 1]);
            pk2[e >> 1] = pack2<kBF16>(d0, d1);
          } else {
            pk[e >> 1] = pack2<kBF16>(d0, d1);
          }
        } else {
          pk[e >> 1] = pack2<kBF16>(pf[e], pf[e + 1]);
        }
      }
      // the 16-bit tile over the first 16 of my own 32 score columns (P, or dS in the one-output dK / dQ modes); mode dKV
      // puts dS over my dP columns likewise
      tmem_st_x16(tmem + lane_base + kColS + half * 32, pk);
      if (kTwoOut) tmem_st_x16(tmem + lane_base + kColP + half * 32, pk2);
      tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();  // (B) the 16-bit tile is complete

    if (tid == 0) {
      tc_fence_after();
      // k-step ks covers streamed rows [16 ks, 16 ks + 16): 16-bit A columns 32 (ks / 2) + 8 (ks % 2) of the score tile
      const uint32_t t_first = sStr + s * 2 * L::kStrBytes;  // Q | K
      const uint32_t t_second = t_first + L::kStrBytes;       // dO | V
      const uint32_t out_lo = smem_desc_lo((kMode == kBwdWideDV || kTwoOut) ? t_second : t_first, 8192);
#pragma unroll
      for (int ks = 0; ks < kWideT / 16; ++ks) {  // contraction over the 64 streamed rows
        umma_ts2(tmem + kColAcc, tmem + kColS + (ks >> 1) * 32 + (ks & 1) * 8, out_lo + ((ks * 2048) >> 4),
                 smem_desc_hi_sw128(1024), idesc_o, (it > 0) || (ks > 0));
      }
      if (kTwoOut) {  // dK += dS^T Q
        const uint32_t out2_lo = smem_desc_lo(t_first, 8192);
#pragma unroll
        for (int ks = 0; ks < kWideT / 16; ++ks) {
          umma_ts2(tmem + kColAcc2, tmem + kColP + (ks >> 1) * 32 + (ks & 1) * 8, out2_lo + ((ks * 2048) >> 4),
                   smem_desc_hi_sw128(1024), idesc_o, (it > 0) || (ks > 0));
        }
      }
      tc_commit(bar_mma2);
      if (kS == 1 && it + 1 < n_iter) {  // one stage: the next tile pair can only be fetched now
        mbar_wait(bar_mma2, waited2 & 1, 84);
        ++waited2;
        load_str(it + 1);
      }
    }
  }

  // ---- epilogue: accumulator -> x out_scale -> 16 bit -> swizzled smem (the first resident tile) -> TMA store.
  // With no visible streamed tile (causal, keys beyond the last query) the gradients of this tile are zero.
  if (tid == 0) {
    while (waited2 < n_iter) {
      mbar_wait(bar_mma2, waited2 & 1, 85);
      ++waited2;
    }
  }
  __syncthreads();
  tc_fence_after();
  constexpr int kHalfD = kDP / 2;
  auto acc_to_stage = [&](uint32_t col_acc, uint8_t* stage, float out_scale) {
#pragma unroll 1
    for (int cidx = 0; cidx < kHalfD / 32; ++cidx) {
      uint32_t a[32];
      if (n_iter > 0) {
        tmem_ld_x32(tmem + lane_base + col_acc + half * kHalfD + cidx * 32, a);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) a[e] = 0u;
      }
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint4 vv;
        vv.x = pack2<kBF16>(__uint_as_float(a[ch * 8 + 0]) * out_scale, __uint_as_float(a[ch * 8 + 1]) * out_scale);
        vv.y = pack2<kBF16>(__uint_as_float(a[ch * 8 + 2]) * out_scale, __uint_as_float(a[ch * 8 + 3]) * out_scale);
        vv.z = pack2<kBF16>(__uint_as_float(a[ch * 8 + 4]) * out_scale, __uint_as_float(a[ch * 8 + 5]) * out_scale);
        vv.w = pack2<kBF16>(__uint_as_float(a[ch * 8 + 6]) * out_scale, __uint_as_float(a[ch * 8 + 7]) * out_scale);
        *reinterpret_cast<uint4*>(stage + sw128_offset_16bit(r, half * kHalfD + cidx * 32 + ch * 8)) = vv;
      }
    }
  };
  uint8_t* stage = smem + L::kRes;                  // the first resident tile
  uint8_t* stage2 = smem + L::kRes + L::kResBytes;  // mode dKV: the second one (dK)
  acc_to_stage(kColAcc, stage, kTwoOut ? 1.f : p.out_scale);
  if (kTwoOut) acc_to_stage(kColAcc2, stage2, p.out_scale);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int db = 0; db < kDBlocks; ++db) {
      tma_store_4d(&tmap_out, smem_u32(stage) + db * 16384, db * 64, row0, h, b);
      if (kTwoOut) tma_store_4d(&tmap_out2, smem_u32(stage2) + db * 16384, db * 64, row0, h, b);
    }
    tma_store_commit();
    tma_store_wait_read();
  }
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace fa
