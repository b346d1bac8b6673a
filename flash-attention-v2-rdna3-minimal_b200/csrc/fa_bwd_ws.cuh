// Warp-specialised tcgen05 backward kernel ("bwd ws", v2): same mathematics and the same CTA
// decomposition as fa_bwd_tc.cuh (one CTA per 128-key K/V tile walking the query tiles; reference:
// /root/reference/rocwmma_fattn/kernel_fp16.cu:547-740), re-arranged so that the tensor cores, the
// TMA loads and the P/dS pass overlap:
//
//   * scores are computed TRANSPOSED, S^T = K_j Q_i^T and dP^T = V_j dO_i^T (TMEM lane = key), so that
//     P^T and dS^T - rounded to 16 bit and written back over their own accumulators - feed
//     dV += P^T dO_i and dK += dS^T Q_i straight from TMEM (tcgen05.mma TS form): no P tile in
//     shared memory, half the operand traffic for two of the five products;
//   * the 32 KB that frees double-buffers Q_i / dO_i, so the TMA loads of tile i+1 (and i+2) run
//     under the work on tile i;
//   * a dedicated warp issues TMA and MMA; S^T(i+1) is issued right behind dQ(i), i.e. while the
//     256 P/dS threads are still draining dQ(i);
//   * dQ_i = dS K_j (A = dS from shared memory, written transposed by the P/dS threads) lands in the
//     dP columns and leaves through the fp32 staging tile + TMA reduce-add of v1.
//
// TMEM: S^T [0,128) (P^T of query half h over [64h,64h+32)), dP^T [128,256) (dS^T likewise; later
// dQ_i over [128,128+D)), dV [256,256+D), dK [384,384+D).
// Shared memory (D = 128): K 32 + V 32 + Q 2x32 + dO 2x32 + dS 32 (aliased by the dQ staging) = 224 KB.
#pragma once
#include "fa_bwd_tc.cuh"

namespace fa {

constexpr int kBwdWsThreads = 288;  // 8 P/dS warps + 1 TMA/MMA warp

template <int kDP>
struct BwdWsSmem {
  static constexpr int kTileBytes = kTileM * kDP * 2;
  static constexpr int kK = 0;
  static constexpr int kV = kK + kTileBytes;
  static constexpr int kQ = kV + kTileBytes;             // 2 buffers
  static constexpr int kdO = kQ + 2 * kTileBytes;        // 2 buffers
  static constexpr int kdS = kdO + 2 * kTileBytes;       // [128 q][128 keys] 16 bit; dQ staging
  static constexpr int kLD = kdS + kTileM * kTileN * 2;  // float [2 (L, D)][128] of the current Q tile
  static constexpr int kBars = kLD + 2 * 128 * 4;
  static constexpr int kTotal = kBars + 128 + 1024;      // + alignment slack
};

template <int kDP, bool kBF16, bool kCausal>
__global__ void __launch_bounds__(kBwdWsThreads, 1)
fa_bwd_ws_kernel(const __grid_constant__ CUtensorMap tmap_q,
                 const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v,
                 const __grid_constant__ CUtensorMap tmap_do,
                 const __grid_constant__ CUtensorMap tmap_dk,
                 const __grid_constant__ CUtensorMap tmap_dv,
                 const __grid_constant__ CUtensorMap tmap_dq, const BwdParams p) {
  using L = BwdWsSmem<kDP>;
  constexpr int kDBlocks = kDP / 64;
  constexpr int kKSteps = kDP / 16;  // contraction over the head dim (S^T, dP^T)
  constexpr int kHalfD = kDP / 2;
  constexpr uint32_t kColS = 0, kColdP = 128, kColdV = 256, kColdK = 384, kColdQ = 128;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t sK = smem_u32(smem + L::kK);
  const uint32_t sV = smem_u32(smem + L::kV);
  const uint32_t sQ = smem_u32(smem + L::kQ);
  const uint32_t sdO = smem_u32(smem + L::kdO);
  const uint32_t sdS = smem_u32(smem + L::kdS);
  float* sLD = reinterpret_cast<float*>(smem + L::kLD);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kBars + 96);
  const uint32_t bar_kv = smem_u32(&bars[0]);                                     // tx
  auto bar_qdo_full = [&](int b_) { return smem_u32(&bars[1 + b_]); };            // tx
  auto bar_qdo_free = [&](int b_) { return smem_u32(&bars[3 + b_]); };            // commit (dV, dK done)
  const uint32_t bar_s = smem_u32(&bars[5]);        // commit: S^T ready
  const uint32_t bar_dp = smem_u32(&bars[6]);       // commit: dP^T ready
  const uint32_t bar_p_ready = smem_u32(&bars[7]);  // 8 warps: P^T, dS^T in TMEM and dS in smem
  const uint32_t bar_dq = smem_u32(&bars[8]);       // commit: dQ_i ready (every earlier MMA done)
  const uint32_t bar_drained = smem_u32(&bars[9]);  // 8 warps: dQ_i has left TMEM

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int j = blockIdx.x;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int key0 = j * kTileN;
  const int i_end = (p.Nq + kTileM - 1) / kTileM;
  const int i_begin = kCausal ? min(j, i_end) : 0;
  const int n_iter = i_end - i_begin;
  const int64_t bh = static_cast<int64_t>(b) * p.H + h;

  if (tid == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_dk);
    tma_prefetch_desc(&tmap_dv);
    tma_prefetch_desc(&tmap_dq);
    mbar_init(bar_kv, 1);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_qdo_full(i), 1);
      mbar_init(bar_qdo_free(i), 1);
    }
    mbar_init(bar_s, 1);
    mbar_init(bar_dp, 1);
    mbar_init(bar_p_ready, 8);
    mbar_init(bar_dq, 1);
    mbar_init(bar_drained, 8);
    fence_mbar_init();
  }
  if (warp == 8) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  constexpr uint32_t idesc_s = make_idesc_f16(kTileM, kTileN, kBF16, false, false);  // S^T, dP^T
  constexpr uint32_t idesc_t = make_idesc_f16(kTileM, kDP, kBF16, false, true);      // dV, dK, dQ

  if (warp == 8) {
    // =========================================================================================
    // TMA + MMA warp (one elected thread)
    // =========================================================================================
    if (elect_one() && n_iter > 0) {
      auto load_qdo = [&](int it) {
        const int buf = it & 1;
        const int i = i_begin + it;
        mbar_arrive_expect_tx(bar_qdo_full(buf), 2 * L::kTileBytes);
#pragma unroll
        for (int db = 0; db < kDBlocks; ++db) {
          tma_load_4d(sQ + buf * L::kTileBytes + db * 16384, &tmap_q, bar_qdo_full(buf), db * 64,
                      i * kTileM, h, b);
          tma_load_4d(sdO + buf * L::kTileBytes + db * 16384, &tmap_do, bar_qdo_full(buf), db * 64,
                      i * kTileM, h, b);
        }
      };
      auto issue_s = [&](int it) {  // S^T = K Q_i^T
        const uint32_t q = sQ + (it & 1) * L::kTileBytes;
#pragma unroll
        for (int k = 0; k < kKSteps; ++k) {
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_ss(tmem + kColS, make_smem_desc_sw128(sK + off, 16, 1024),
                  make_smem_desc_sw128(q + off, 16, 1024), idesc_s, k > 0);
        }
        tc_commit(bar_s);
      };
      auto issue_dp = [&](int it) {  // dP^T = V dO_i^T
        const uint32_t d_o = sdO + (it & 1) * L::kTileBytes;
#pragma unroll
        for (int k = 0; k < kKSteps; ++k) {
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_ss(tmem + kColdP, make_smem_desc_sw128(sV + off, 16, 1024),
                  make_smem_desc_sw128(d_o + off, 16, 1024), idesc_s, k > 0);
        }
        tc_commit(bar_dp);
      };

      mbar_arrive_expect_tx(bar_kv, 2 * L::kTileBytes);
#pragma unroll
      for (int db = 0; db < kDBlocks; ++db) {
        tma_load_4d(sK + db * 16384, &tmap_k, bar_kv, db * 64, key0, h, b);
        tma_load_4d(sV + db * 16384, &tmap_v, bar_kv, db * 64, key0, h, b);
      }
      load_qdo(0);
      if (n_iter > 1) load_qdo(1);
      mbar_wait(bar_kv, 0, 60);
      mbar_wait(bar_qdo_full(0), 0, 61);
      tc_fence_after();
      issue_s(0);
      issue_dp(0);

#pragma unroll 1
      for (int it = 0; it < n_iter; ++it) {
        const int buf = it & 1;
        const uint32_t q = sQ + buf * L::kTileBytes;
        const uint32_t d_o = sdO + buf * L::kTileBytes;
        mbar_wait(bar_p_ready, it & 1, 62);
        tc_fence_after();
        // k-step ks covers queries [16 ks, 16 ks + 16): 16-bit A columns of half ks/4 at
        // 64 (ks/4) + 8 (ks%4) of the S^T / dP^T accumulator; B rows 16 ks of the dO / Q tile
#pragma unroll
        for (int ks = 0; ks < kTileM / 16; ++ks) {  // dV += P^T dO_i
          umma_ts(tmem + kColdV, tmem + kColS + (ks >> 2) * 64 + (ks & 3) * 8,
                  make_smem_desc_sw128(d_o + ks * 2048, 16384, 1024), idesc_t, (it > 0) || (ks > 0));
        }
#pragma unroll
        for (int ks = 0; ks < kTileM / 16; ++ks) {  // dK += dS^T Q_i
          umma_ts(tmem + kColdK, tmem + kColdP + (ks >> 2) * 64 + (ks & 3) * 8,
                  make_smem_desc_sw128(q + ks * 2048, 16384, 1024), idesc_t, (it > 0) || (ks > 0));
        }
        tc_commit(bar_qdo_free(buf));
#pragma unroll
        for (int k = 0; k < kTileN / 16; ++k) {  // dQ_i = dS K_j over the dP^T columns
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_ss(tmem + kColdQ, make_smem_desc_sw128(sdS + off, 16, 1024),
                  make_smem_desc_sw128(sK + k * 2048, 16384, 1024), idesc_t, k > 0);
        }
        tc_commit(bar_dq);
        if (it + 1 < n_iter) {
          mbar_wait(bar_qdo_full(buf ^ 1), ((it + 1) >> 1) & 1, 63);
          tc_fence_after();
          issue_s(it + 1);  // in order behind dV(it), which read P^T from these columns
          if (it + 2 < n_iter) {
            mbar_wait(bar_qdo_free(buf), (it >> 1) & 1, 64);  // dV(it), dK(it) done with buffer `buf`
            load_qdo(it + 2);
          }
          mbar_wait(bar_drained, it & 1, 65);  // dQ(it) has left the dP^T columns
          tc_fence_after();
          issue_dp(it + 1);
        }
      }
    }
    __syncwarp();
  } else {
    // =========================================================================================
    // P / dS warps: thread (r, half) owns key row r of the tile and the 64-query half `half`
    // =========================================================================================
    const int half = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int key = key0 + r;
    const bool key_ok = key < p.Nkv;
    // shared-memory address pieces of column `r` of the 128-byte-swizzled [q][key] dS tile
    const uint32_t ds_col = sdS + (r >> 6) * 16384 + (r & 7) * 2;
    const uint32_t key_chunk = (r & 63) >> 3;
    const float c = p.scale_log2;

    const float* ld_src = (tid < 128) ? p.lse : p.delta;
    float ld_next = 0.f;
    if (n_iter > 0) {
      const int qrow = i_begin * kTileM + (tid & 127);
      ld_next = (qrow < p.Nq) ? ld_src[bh * p.Nq + qrow] : 0.f;
    }
#pragma unroll 1
    for (int it = 0; it < n_iter; ++it) {
      const int i = i_begin + it;
      const int par = it & 1;
      // stage L_i and D_i of the 128 queries of this tile (zero for rows >= Nq); the values were
      // fetched from global memory one tile ahead (ld_next), so no load latency is exposed here
      // (every thread passed the drain barriers of the previous tile after its last read of sLD)
      sLD[tid] = ld_next;
      if (it + 1 < n_iter) {
        const int qrow = (i + 1) * kTileM + (tid & 127);
        ld_next = (qrow < p.Nq) ? ld_src[bh * p.Nq + qrow] : 0.f;
      }
      named_bar_sync(1, 256);
      const float* sL = sLD;
      const float* sD = sL + 128;
      const bool need_mask = (kCausal && i == j) || (key0 + kTileN > p.Nkv) || ((i + 1) * kTileM > p.Nq);

      // ---- phase A: P^T = 2^(S^T c - L) for my 64 queries; 16-bit copy over the S^T columns
      mbar_wait(bar_s, par, 66);
      tc_fence_after();
      float pf[64];
#pragma unroll
      for (int q2 = 0; q2 < 2; ++q2) {
        const int qb = half * 64 + q2 * 32;  // first query (inside the tile) of this chunk
        uint32_t sv[32];
        tmem_ld_x32(tmem + lane_base + kColS + qb, sv);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          float pe = ex2_approx(fmaf(__uint_as_float(sv[e]), c, -sL[qb + e]));
          if (need_mask) {
            const int qrow = i * kTileM + qb + e;
            const bool ok = key_ok && qrow < p.Nq && (!kCausal || key <= qrow);
            pe = ok ? pe : 0.f;
          }
          pf[q2 * 32 + e] = pe;
        }
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[e] = pack2<kBF16>(pf[q2 * 32 + 2 * e], pf[q2 * 32 + 2 * e + 1]);
        tmem_st_x16(tmem + lane_base + kColS + half * 64 + q2 * 16, pk);
      }

      // ---- phase B: dS^T = P^T o (dP^T - D); 16-bit copy over the dP^T columns and, transposed,
      // into the [q][key] shared-memory tile that is the A operand of dQ_i = dS K_j
      mbar_wait(bar_dp, par, 67);
      tc_fence_after();
#pragma unroll
      for (int q2 = 0; q2 < 2; ++q2) {
        const int qb = half * 64 + q2 * 32;
        uint32_t dv[32];
        tmem_ld_x32(tmem + lane_base + kColdP + qb, dv);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float d0 = pf[q2 * 32 + e] * (__uint_as_float(dv[e]) - sD[qb + e]);
          const float d1 = pf[q2 * 32 + e + 1] * (__uint_as_float(dv[e + 1]) - sD[qb + e + 1]);
          const uint32_t w = pack2<kBF16>(d0, d1);
          pk[e >> 1] = w;
          // element (query qb+e, key r) of the [q][key] tile; qb is a multiple of 32, so the swizzle
          // term (q & 7) is the compile-time e & 7
          st_shared_u16(ds_col + (qb + e) * 128 + ((key_chunk ^ (e & 7)) << 4), w & 0xffffu);
          st_shared_u16(ds_col + (qb + e + 1) * 128 + ((key_chunk ^ ((e + 1) & 7)) << 4), w >> 16);
        }
        tmem_st_x16(tmem + lane_base + kColdP + half * 64 + q2 * 16, pk);
      }
      tmem_wait_st();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p_ready);

      // ---- drain dQ_i (TMEM lane = query row here): 32 head-dim columns at a time through the
      // swizzled fp32 staging tile (aliases the dS tile, dead once dQ_i is complete) + TMA reduce-add
      mbar_wait(bar_dq, par, 68);
      tc_fence_after();
#pragma unroll 1
      for (int cidx = 0; cidx < kDP / 32; ++cidx) {
        uint8_t* stage = smem + L::kdS + (cidx & 1) * (kTileM * 128);
        uint32_t v[16];
        tmem_ld_x16(tmem + lane_base + kColdQ + cidx * 32 + half * 16, v);
        if (tid == 0 && cidx >= 2) tma_store_wait_read_1();  // the reduce that last read this tile
        named_bar_sync(2, 256);
        tmem_wait_ld();
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = half * 4 + ch;  // 16-byte chunk inside the 128-byte row
          *reinterpret_cast<uint4*>(stage + r * 128 + ((chunk ^ (r & 7)) << 4)) =
              make_uint4(v[ch * 4 + 0], v[ch * 4 + 1], v[ch * 4 + 2], v[ch * 4 + 3]);
        }
        fence_proxy_async_smem();
        named_bar_sync(3, 256);
        if (tid == 0) {
          tma_reduce_add_3d(&tmap_dq, smem_u32(stage), cidx * 32, i * kTileM, static_cast<int>(bh));
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_drained);
      // the staging tiles alias the dS tile that phase B of the next query tile overwrites: every
      // reduce must have read its source first (the named barrier at the loop top publishes this)
      if (tid == 0) tma_store_wait_read();
    }

    // ---- epilogue: dV and scale * dK (TMEM lane = key row) -> 16 bit -> swizzled smem (the two
    // Q buffers) -> TMA store.  bar_dq of the last tile covered every MMA.  With no visible query
    // tile (causal, keys beyond the last query) the gradients of this key tile are zero.
    tc_fence_after();
    uint8_t* st_dv = smem + L::kQ;
    uint8_t* st_dk = smem + L::kQ + L::kTileBytes;
#pragma unroll
    for (int cidx = 0; cidx < kHalfD / 32; ++cidx) {
      uint32_t a[32], k2[32];
      if (n_iter > 0) {
        tmem_ld_x32(tmem + lane_base + kColdV + half * kHalfD + cidx * 32, a);
        tmem_ld_x32(tmem + lane_base + kColdK + half * kHalfD + cidx * 32, k2);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) a[e] = k2[e] = 0u;
      }
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint4 vv, kk;
        vv.x = pack2<kBF16>(__uint_as_float(a[ch * 8 + 0]), __uint_as_float(a[ch * 8 + 1]));
        vv.y = pack2<kBF16>(__uint_as_float(a[ch * 8 + 2]), __uint_as_float(a[ch * 8 + 3]));
        vv.z = pack2<kBF16>(__uint_as_float(a[ch * 8 + 4]), __uint_as_float(a[ch * 8 + 5]));
        vv.w = pack2<kBF16>(__uint_as_float(a[ch * 8 + 6]), __uint_as_float(a[ch * 8 + 7]));
        kk.x = pack2<kBF16>(__uint_as_float(k2[ch * 8 + 0]) * p.scale, __uint_as_float(k2[ch * 8 + 1]) * p.scale);
        kk.y = pack2<kBF16>(__uint_as_float(k2[ch * 8 + 2]) * p.scale, __uint_as_float(k2[ch * 8 + 3]) * p.scale);
        kk.z = pack2<kBF16>(__uint_as_float(k2[ch * 8 + 4]) * p.scale, __uint_as_float(k2[ch * 8 + 5]) * p.scale);
        kk.w = pack2<kBF16>(__uint_as_float(k2[ch * 8 + 6]) * p.scale, __uint_as_float(k2[ch * 8 + 7]) * p.scale);
        const uint32_t off = sw128_offset_16bit(r, half * kHalfD + cidx * 32 + ch * 8);
        *reinterpret_cast<uint4*>(st_dv + off) = vv;
        *reinterpret_cast<uint4*>(st_dk + off) = kk;
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 256);
    if (tid == 0) {
#pragma unroll
      for (int db = 0; db < kDBlocks; ++db) {
        tma_store_4d(&tmap_dv, smem_u32(st_dv) + db * 16384, db * 64, key0, h, b);
        tma_store_4d(&tmap_dk, smem_u32(st_dk) + db * 16384, db * 64, key0, h, b);
      }
      tma_store_commit();
      tma_store_wait_read();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}

}  // namespace fa
