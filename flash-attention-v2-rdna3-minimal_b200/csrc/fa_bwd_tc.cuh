// tcgen05 backward kernels (dQ, dK, dV) - SURVEY.md section 8(f) rank 3.
//
// Replaces /root/reference/rocwmma_fattn/kernel_fp16.cu:547-740 / kernel_bf16.cu:580-798 (bwd_kernel)
// and the launchers backward_fp16 / backward_bf16 (:878-1028 / :943-1092).  Same decomposition as
// the reference - one CTA per 128-key K/V tile walking the query tiles, D_i = rowsum(dO o O)
// (kernel_fp16.cu:605-631), P = 2^(S - L) from the forward's base-2 LSE (:698-719), dV += P^T dO
// (:724), dP = dO V^T (:725), dS = P o (dP - D) (:732), dQ += dS K (:736), dK += dS^T Q (:737) -
// with three deliberate differences:
//   * the five products run on the tensor cores (tcgen05.mma, fp32 accumulators in TMEM);
//   * dQ is accumulated across K/V tiles in an fp32 buffer with TMA reduce-add (the reference adds
//     into an fp16 dQ from different CTAs without atomics, kernel_fp16.cu:736 - a data race);
//   * D_i is computed once by a small pre-pass instead of by every CTA.
//
// Per (K/V tile j, Q tile i), 256 threads (two per query row, key halves like the forward):
//   S  = Q_i K_j^T          SS MMA, both K-major                         -> TMEM [0,128)
//   dP = dO_i V_j^T         SS MMA, both K-major                         -> TMEM [128,256)
//   P = 2^(S c - L_i), dS = P (dP - D_i) in registers -> 16 bit -> smem (128-byte swizzle)
//   dV += P^T  dO_i         SS MMA, A = P  MN-major, B = dO MN-major     -> TMEM [256,256+D)
//   dK += dS^T Q_i          SS MMA, A = dS MN-major, B = Q  MN-major     -> TMEM [384,384+D)
//   dQ_i = dS K_j           SS MMA, A = dS K-major,  B = K  MN-major     -> TMEM [0,D) (over S)
//   dQ_i: TMEM -> registers -> swizzled fp32 staging -> TMA reduce-add (cp.reduce.async.bulk) into dq_acc
// One elected thread issues the TMA loads and every MMA; this first version is a serial pipeline
// (the tensor cores idle during the P/dS pass and the dQ drain).
#pragma once
#include "fa_fwd_tc.cuh"

namespace fa {

// Of every 4 pairs of P elements, how many compute 2^x on the FMA pipes instead of the MUFU (backward).
// Measured (fp16, D=128, N=16384 / 4096): 0 -> 685 / 557 TFLOPS, 1 -> 669 / 546, 2 -> 660 / 537: the MUFU is not
// what bounds the P / dS pass (its shared-memory stores and the TMEM round trips are), so the default is 0.
#ifndef FA_BWD_EMU_PAIRS
#define FA_BWD_EMU_PAIRS 0
#endif
constexpr int kBwdEmuPairs = FA_BWD_EMU_PAIRS;

struct BwdParams {
  const float* lse;    // [B,H,Nq] base-2 log-sum-exp of the scaled scores (forward output)
  const float* delta;  // [B,H,Nq] rowsum(dO o O), fp32 (fa_bwd_delta_kernel)
  float* dq_acc;       // [B,H,Nq,dq_ld] fp32, zeroed by fa_bwd_delta_kernel; receives sum_j dS K_j
  int Nq, Nkv, H;
  int dq_ld;           // row pitch of dq_acc in floats (the padded head dim)
  float scale_log2;    // scale * log2(e)
  float scale;
  int lpt_group;       // fa_bwd_ws.cuh, causal: (batch, head) pairs per longest-first group of the 1-D grid
  unsigned long long* trace;  // debug builds only (-DFA_TRACE): clock64() stamps of one CTA (tools/trace_bwd.py)
};

template <int kDP>
struct BwdSmem {
  static constexpr int kTileBytes = kTileM * kDP * 2;
  static constexpr int kK = 0;
  static constexpr int kV = kK + kTileBytes;
  static constexpr int kQ = kV + kTileBytes;
  static constexpr int kdO = kQ + kTileBytes;
  static constexpr int kP = kdO + kTileBytes;            // [128 q][128 keys] 16 bit; dV staging
  static constexpr int kdS = kP + kTileM * kTileN * 2;    // [128 q][128 keys] 16 bit; dK staging
  static constexpr int kStage = kdS + kTileM * kTileN * 2;  // 2 x [128 rows][32 fp32] dQ staging
  static constexpr int kBars = kStage + 2 * kTileM * 128;
  static constexpr int kTotal = kBars + 128 + 1024;       // + alignment slack
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// pre-pass: delta_i = sum_d dO_id O_id (fp32) and dq_acc row = 0.  One warp per query row.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
fa_bwd_delta_kernel(const T* __restrict__ o, const T* __restrict__ d_o, float* __restrict__ delta,
                    float* __restrict__ dq_acc, int B, int H, int Nq, int D, int dq_ld, int64_t os0,
                    int64_t os1, int64_t os2, int64_t ds0, int64_t ds1, int64_t ds2) {
  const int64_t row_id = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int64_t total = static_cast<int64_t>(B) * H * Nq;
  if (row_id >= total) return;
  const int n = static_cast<int>(row_id % Nq);
  const int h = static_cast<int>((row_id / Nq) % H);
  const int b = static_cast<int>(row_id / (static_cast<int64_t>(Nq) * H));
  const T* po = o + b * os0 + h * os1 + n * os2;
  const T* pd = d_o + b * ds0 + h * ds1 + n * ds2;
  float acc = 0.f;
  for (int d = lane; d < D; d += 32) acc += static_cast<float>(po[d]) * static_cast<float>(pd[d]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) delta[row_id] = acc;
  float* pq = dq_acc + row_id * dq_ld;
  for (int d = lane; d < dq_ld; d += 32) pq[d] = 0.f;
}

// post-pass: dQ = scale * dq_acc, rounded to the 16-bit output type.  One warp per query row.
template <typename T>
__global__ void __launch_bounds__(256)
fa_bwd_dq_convert_kernel(const float* __restrict__ dq_acc, T* __restrict__ dq, int B, int H, int Nq,
                         int D, int dq_ld, int64_t s0, int64_t s1, int64_t s2, float scale) {
  const int64_t row_id = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int64_t total = static_cast<int64_t>(B) * H * Nq;
  if (row_id >= total) return;
  const int n = static_cast<int>(row_id % Nq);
  const int h = static_cast<int>((row_id / Nq) % H);
  const int b = static_cast<int>(row_id / (static_cast<int64_t>(Nq) * H));
  const float* src = dq_acc + row_id * dq_ld;
  T* dst = dq + b * s0 + h * s1 + n * s2;
  for (int d = lane; d < D; d += 32) dst[d] = static_cast<T>(src[d] * scale);
}

// The same two passes for 16-byte-aligned rows and head dims that are a multiple of 8 (everything the tensor-core
// kernels take): 16-byte loads and stores, 16 lanes per row (two rows per warp) up to head dim 128.  At N <= 4096 the
// passes are 10-30 % of a backward call, so their bandwidth matters: the scalar versions above move 2 bytes per lane
// and instruction.
template <typename T, int kLanes>
__global__ void __launch_bounds__(256)
fa_bwd_delta_vec_kernel(const T* __restrict__ o, const T* __restrict__ d_o, float* __restrict__ delta,
                        float* __restrict__ dq_acc, int B, int H, int Nq, int D, int dq_ld, int64_t os0,
                        int64_t os1, int64_t os2, int64_t ds0, int64_t ds1, int64_t ds2) {
  constexpr int kRowsPerWarp = 32 / kLanes;
  const int lane = threadIdx.x & 31;
  const int sub = lane % kLanes;
  const int64_t row_id = (static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5)) * kRowsPerWarp + lane / kLanes;
  const int64_t total = static_cast<int64_t>(B) * H * Nq;
  const bool live = row_id < total;
  float acc = 0.f;
  if (live) {
    const int n = static_cast<int>(row_id % Nq);
    const int h = static_cast<int>((row_id / Nq) % H);
    const int b = static_cast<int>(row_id / (static_cast<int64_t>(Nq) * H));
    const uint4* po = reinterpret_cast<const uint4*>(o + b * os0 + h * os1 + n * os2);
    const uint4* pd = reinterpret_cast<const uint4*>(d_o + b * ds0 + h * ds1 + n * ds2);
    for (int c = sub; c < D / 8; c += kLanes) {
      const uint4 a = __ldg(po + c), g = __ldg(pd + c);
      const T* av = reinterpret_cast<const T*>(&a);
      const T* gv = reinterpret_cast<const T*>(&g);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc = fmaf(static_cast<float>(av[e]), static_cast<float>(gv[e]), acc);
    }
    float4* pq = reinterpret_cast<float4*>(dq_acc + row_id * dq_ld);
    for (int c = sub; c < dq_ld / 4; c += kLanes) pq[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int off = kLanes / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (live && sub == 0) delta[row_id] = acc;
}

template <typename T, int kLanes>
__global__ void __launch_bounds__(256)
fa_bwd_dq_convert_vec_kernel(const float* __restrict__ dq_acc, T* __restrict__ dq, int B, int H, int Nq, int D,
                             int dq_ld, int64_t s0, int64_t s1, int64_t s2, float scale) {
  pdl_wait();  // launched with programmatic stream serialization behind the main kernel
  constexpr int kRowsPerWarp = 32 / kLanes;
  const int lane = threadIdx.x & 31;
  const int sub = lane % kLanes;
  const int64_t row_id = (static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5)) * kRowsPerWarp + lane / kLanes;
  const int64_t total = static_cast<int64_t>(B) * H * Nq;
  if (row_id >= total) return;
  const int n = static_cast<int>(row_id % Nq);
  const int h = static_cast<int>((row_id / Nq) % H);
  const int b = static_cast<int>(row_id / (static_cast<int64_t>(Nq) * H));
  const float4* src = reinterpret_cast<const float4*>(dq_acc + row_id * dq_ld);
  uint4* dst = reinterpret_cast<uint4*>(dq + b * s0 + h * s1 + n * s2);
  for (int c = sub; c < D / 8; c += kLanes) {
    const float4 lo = src[2 * c], hi = src[2 * c + 1];
    uint4 out;
    T* ov = reinterpret_cast<T*>(&out);
    ov[0] = static_cast<T>(lo.x * scale); ov[1] = static_cast<T>(lo.y * scale);
    ov[2] = static_cast<T>(lo.z * scale); ov[3] = static_cast<T>(lo.w * scale);
    ov[4] = static_cast<T>(hi.x * scale); ov[5] = static_cast<T>(hi.y * scale);
    ov[6] = static_cast<T>(hi.z * scale); ov[7] = static_cast<T>(hi.w * scale);
    dst[c] = out;
  }
}

// ---------------------------------------------------------------------------------------------
// main kernel: one CTA per (batch, head, 128-key tile)
// ---------------------------------------------------------------------------------------------
template <int kDP, bool kBF16, bool kCausal>
__global__ void __launch_bounds__(256, 1)
fa_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q,
                 const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v,
                 const __grid_constant__ CUtensorMap tmap_do,
                 const __grid_constant__ CUtensorMap tmap_dk,
                 const __grid_constant__ CUtensorMap tmap_dv,
                 const __grid_constant__ CUtensorMap tmap_dq, const BwdParams p) {
  using L = BwdSmem<kDP>;
  constexpr int kDBlocks = kDP / 64;
  constexpr int kKSteps = kDP / 16;     // contraction over the head dim (S, dP)
  constexpr int kHalfD = kDP / 2;       // dQ / dK / dV columns each of a row's two threads owns
  constexpr uint32_t kColS = 0, kColdP = 128, kColdV = 256, kColdK = 384, kColdQ = 0;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t sK = smem_u32(smem + L::kK);
  const uint32_t sV = smem_u32(smem + L::kV);
  const uint32_t sQ = smem_u32(smem + L::kQ);
  const uint32_t sdO = smem_u32(smem + L::kdO);
  const uint32_t sP = smem_u32(smem + L::kP);
  const uint32_t sdS = smem_u32(smem + L::kdS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kBars + 64);
  const uint32_t bar_kv = smem_u32(&bars[0]);    // tx: K_j, V_j
  const uint32_t bar_qdo = smem_u32(&bars[1]);   // tx: Q_i, dO_i
  const uint32_t bar_mma1 = smem_u32(&bars[2]);  // commit: S, dP ready
  const uint32_t bar_free = smem_u32(&bars[3]);  // commit: dV, dK done -> Q_i, dO_i buffers free
  const uint32_t bar_dq = smem_u32(&bars[4]);    // commit: dQ_i ready (and every earlier MMA done)

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int half = warp >> 2;                     // which 64-key half of the row this thread owns
  const int r = (warp & 3) * 32 + lane;           // row inside a tile = TMEM lane
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const int j = blockIdx.x;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int key0 = j * kTileN;

  const int i_end = (p.Nq + kTileM - 1) / kTileM;
  const int i_begin = kCausal ? min(j, i_end) : 0;  // query tiles above the diagonal see no key of j
  const int n_iter = i_end - i_begin;

  if (tid == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_dk);
    tma_prefetch_desc(&tmap_dv);
    tma_prefetch_desc(&tmap_dq);
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

  auto load_qdo = [&](int i) {
    mbar_arrive_expect_tx(bar_qdo, 2 * L::kTileBytes);
#pragma unroll
    for (int db = 0; db < kDBlocks; ++db) {
      tma_load_4d(sQ + db * 16384, &tmap_q, bar_qdo, db * 64, i * kTileM, h, b);
      tma_load_4d(sdO + db * 16384, &tmap_do, bar_qdo, db * 64, i * kTileM, h, b);
    }
  };

  if (tid == 0 && n_iter > 0) {
    mbar_arrive_expect_tx(bar_kv, 2 * L::kTileBytes);
#pragma unroll
    for (int db = 0; db < kDBlocks; ++db) {
      tma_load_4d(sK + db * 16384, &tmap_k, bar_kv, db * 64, key0, h, b);
      tma_load_4d(sV + db * 16384, &tmap_v, bar_kv, db * 64, key0, h, b);
    }
    load_qdo(i_begin);
  }

  constexpr uint32_t idesc_s = make_idesc_f16(kTileM, kTileN, kBF16, false, false);  // S, dP
  constexpr uint32_t idesc_t = make_idesc_f16(kTileM, kDP, kBF16, true, true);       // dV, dK
  constexpr uint32_t idesc_q = make_idesc_f16(kTileM, kDP, kBF16, false, true);      // dQ
  const float c = p.scale_log2;
  const int64_t bh = static_cast<int64_t>(b) * p.H + h;
  // L_i / D_i of my query row, fetched one tile ahead so the global-load latency is not exposed
  float lse_next = 0.f, dl_next = 0.f;
  if (n_iter > 0 && i_begin * kTileM + r < p.Nq) {
    lse_next = p.lse[bh * p.Nq + i_begin * kTileM + r];
    dl_next = p.delta[bh * p.Nq + i_begin * kTileM + r];
  }

#pragma unroll 1
  for (int it = 0; it < n_iter; ++it) {
    const int i = i_begin + it;
    const uint32_t ph = it & 1;
    if (tid == 0) {
      if (it == 0) mbar_wait(bar_kv, 0, 60);
      mbar_wait(bar_qdo, ph, 61);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < kKSteps; ++k) {
        const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
        umma_ss(tmem + kColS, make_smem_desc_sw128(sQ + off, 16, 1024),
                make_smem_desc_sw128(sK + off, 16, 1024), idesc_s, k > 0);
      }
#pragma unroll
      for (int k = 0; k < kKSteps; ++k) {
        const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
        umma_ss(tmem + kColdP, make_smem_desc_sw128(sdO + off, 16, 1024),
                make_smem_desc_sw128(sV + off, 16, 1024), idesc_s, k > 0);
      }
      tc_commit(bar_mma1);
    }

    // ---- P and dS for my half of the row
    const int row = i * kTileM + r;
    const bool row_ok = row < p.Nq;
    const float lse = lse_next;
    const float dl = dl_next;
    if (it + 1 < n_iter) {
      const int nrow = row + kTileM;
      lse_next = (nrow < p.Nq) ? p.lse[bh * p.Nq + nrow] : 0.f;
      dl_next = (nrow < p.Nq) ? p.delta[bh * p.Nq + nrow] : 0.f;
    }
    const float nl = -lse;
    // masks are needed on the diagonal tile, on the last (partial) key tile and on partial Q tiles
    const bool need_mask = (kCausal && i == j) || (key0 + kTileN > p.Nkv) || ((i + 1) * kTileM > p.Nq);
    mbar_wait(bar_mma1, ph, 62);
    tc_fence_after();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int cb = half * 64 + q * 32;  // first column (key inside the tile) of this chunk
      uint32_t sv[32], dv[32];
      tmem_ld_x32(tmem + lane_base + kColS + cb, sv);
      tmem_ld_x32(tmem + lane_base + kColdP + cb, dv);
      tmem_wait_ld();
      float pf[32], df[32];
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        // P = 2^(S c - L): kBwdEmuPairs of every 4 element pairs on the FMA pipes (ex2_fma2), the rest on the
        // MUFU, which is the busiest pipe of this phase (64 exponentials per thread, 16 per clock per SM)
        float x0 = fmaf(__uint_as_float(sv[e]), c, nl), x1 = fmaf(__uint_as_float(sv[e + 1]), c, nl);
        if (((e >> 1) & 3) < kBwdEmuPairs) {
          ex2_fma2(x0, x1);
        } else {
          x0 = ex2_approx(x0);
          x1 = ex2_approx(x1);
        }
        pf[e] = x0;
        pf[e + 1] = x1;
      }
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        float pe = pf[e];
        if (need_mask) {
          const int key = key0 + cb + e;
          const bool ok = row_ok && key < p.Nkv && (!kCausal || key <= row);
          pe = ok ? pe : 0.f;
        }
        pf[e] = pe;
        df[e] = pe * (__uint_as_float(dv[e]) - dl);
      }
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint4 pv, dsv;
        pv.x = pack2<kBF16>(pf[ch * 8 + 0], pf[ch * 8 + 1]);
        pv.y = pack2<kBF16>(pf[ch * 8 + 2], pf[ch * 8 + 3]);
        pv.z = pack2<kBF16>(pf[ch * 8 + 4], pf[ch * 8 + 5]);
        pv.w = pack2<kBF16>(pf[ch * 8 + 6], pf[ch * 8 + 7]);
        dsv.x = pack2<kBF16>(df[ch * 8 + 0], df[ch * 8 + 1]);
        dsv.y = pack2<kBF16>(df[ch * 8 + 2], df[ch * 8 + 3]);
        dsv.z = pack2<kBF16>(df[ch * 8 + 4], df[ch * 8 + 5]);
        dsv.w = pack2<kBF16>(df[ch * 8 + 6], df[ch * 8 + 7]);
        const uint32_t off = sw128_offset_16bit(r, cb + ch * 8);
        *reinterpret_cast<uint4*>(smem + L::kP + off) = pv;
        *reinterpret_cast<uint4*>(smem + L::kdS + off) = dsv;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();

    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < kTileM / 16; ++k) {  // dV += P^T dO: contraction over the 128 query rows
        umma_ss(tmem + kColdV, make_smem_desc_sw128(sP + k * 2048, 16384, 1024),
                make_smem_desc_sw128(sdO + k * 2048, 16384, 1024), idesc_t, (it > 0) || (k > 0));
      }
#pragma unroll
      for (int k = 0; k < kTileM / 16; ++k) {  // dK += dS^T Q
        umma_ss(tmem + kColdK, make_smem_desc_sw128(sdS + k * 2048, 16384, 1024),
                make_smem_desc_sw128(sQ + k * 2048, 16384, 1024), idesc_t, (it > 0) || (k > 0));
      }
      tc_commit(bar_free);
#pragma unroll
      for (int k = 0; k < kTileN / 16; ++k) {  // dQ_i = dS K_j: contraction over the 128 keys
        const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
        umma_ss(tmem + kColdQ, make_smem_desc_sw128(sdS + off, 16, 1024),
                make_smem_desc_sw128(sK + k * 2048, 16384, 1024), idesc_q, k > 0);
      }
      tc_commit(bar_dq);
      if (it + 1 < n_iter) {
        mbar_wait(bar_free, ph, 63);
        load_qdo(i + 1);
      }
    }

    // ---- drain dQ_i: 32 head-dim columns at a time through a swizzled fp32 staging tile, then one
    // TMA reduce-add per chunk: dq_acc[b,h, 128 rows, 32 cols] += staging (rows >= Nq are clipped).
    // Thread (r, half) moves 16 of the 32 columns of row r.
    mbar_wait(bar_dq, ph, 64);
    tc_fence_after();
#pragma unroll 1
    for (int cidx = 0; cidx < kDP / 32; ++cidx) {
      uint8_t* stage = smem + L::kStage + (cidx & 1) * (kTileM * 128);
      uint32_t v[16];
      tmem_ld_x16(tmem + lane_base + kColdQ + cidx * 32 + half * 16, v);
      if (tid == 0) tma_store_wait_read_1();  // the reduce that last read this staging tile is done
      named_bar_sync(1, 256);
      tmem_wait_ld();
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const int chunk = half * 4 + ch;  // 16-byte chunk inside the 128-byte row
        *reinterpret_cast<uint4*>(stage + r * 128 + ((chunk ^ (r & 7)) << 4)) =
            make_uint4(v[ch * 4 + 0], v[ch * 4 + 1], v[ch * 4 + 2], v[ch * 4 + 3]);
      }
      fence_proxy_async_smem();
      named_bar_sync(2, 256);
      if (tid == 0) {
        tma_reduce_add_3d(&tmap_dq, smem_u32(stage), cidx * 32, i * kTileM, static_cast<int>(bh));
        tma_store_commit();
      }
    }
    tc_fence_before();
    __syncthreads();  // S / dQ columns and the P / dS tiles may be overwritten by the next tile
  }
  if (tid == 0) tma_store_wait_read();  // staging tiles are re-used by nothing below, but be tidy

  // ---- epilogue: dV and scale * dK -> 16 bit -> swizzled smem (P / dS tiles) -> TMA store.
  // TMEM lane = key row here.  With no visible query tile (causal, keys beyond the last query)
  // the gradients of this key tile are zero.
  tc_fence_after();
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
      *reinterpret_cast<uint4*>(smem + L::kP + off) = vv;
      *reinterpret_cast<uint4*>(smem + L::kdS + off) = kk;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int db = 0; db < kDBlocks; ++db) {
      tma_store_4d(&tmap_dv, sP + db * 16384, db * 64, key0, h, b);
      tma_store_4d(&tmap_dk, sdS + db * 16384, db * 64, key0, h, b);
    }
    tma_store_commit();
    tma_store_wait_read();
  }
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace fa
