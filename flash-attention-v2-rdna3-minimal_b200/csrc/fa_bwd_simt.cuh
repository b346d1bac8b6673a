// Generic CUDA-core backward kernels: any head dim up to 1024, any strides, fp32 math.
//
// The tcgen05 backward (fa_bwd_tc.cuh) holds S, dP, dV and dK in tensor memory at once, which caps it at
// head dim 128; the reference's launcher pads and serves any head dim (kernel_fp16.cu:878-1028, D-pad at
// :900), e.g. the 160 of SD 1.5.  These two kernels cover what is left - head dims 129..1024, head dims that
// are not a multiple of 8 without padding, unaligned pointers or strides - so that every tensor the forward
// accepts can also be differentiated.  They are the counterpart of fa_fwd_simt.cuh: CUDA kernels, not a CPU
// fallback, and never selected for a BASELINE configuration.
//
//   P_ij  = 2^(s_ij c - L_i)            s = q.k, c = scale log2(e), L = the forward's base-2 LSE
//   dP_ij = dO_i . v_j,  dS_ij = P_ij (dP_ij - delta_i),  delta_i = rowsum(dO o O) (fa_bwd_delta_kernel)
//   dQ_i  = scale sum_j dS_ij k_j       fa_bwd_simt_dq_kernel : one warp per query row
//   dK_j  = scale sum_i dS_ij q_i       fa_bwd_simt_dkv_kernel: one warp per key row
//   dV_j  =       sum_i  P_ij dO_i
// (kernel_fp16.cu:698-737).  Inside a warp the 32 lanes each own one key (or query) of the current group of
// 32 for the two dot products, then the 32 weights are broadcast with shuffles while every lane accumulates
// its strided slice of the output row - the arrangement of the generic forward kernel.
#pragma once
#include "fa_fwd_simt.cuh"

namespace fa {

struct SimtBwdParams {
  const void *q, *k, *v, *d_o;
  void *dq, *dk, *dv;
  const float* lse;    // [B,H,Nq]
  const float* delta;  // [B,H,Nq]
  int B, H, Nq, Nkv, D;
  int64_t qs[4], ks[4], vs[4], dos[4], dqs[4], dks[4], dvs[4];
  int causal;
  float scale, scale_log2;
};

template <typename T>
__global__ void __launch_bounds__(kSimtWarps * 32) fa_bwd_simt_dq_kernel(const SimtBwdParams p) {
  extern __shared__ float rows_smem[];  // [kSimtWarps][2][D]: q_i * c, dO_i
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kSimtWarps + warp;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  if (row >= p.Nq) return;

  const T* q = static_cast<const T*>(p.q) + b * p.qs[0] + h * p.qs[1] + row * p.qs[2];
  const T* d_o = static_cast<const T*>(p.d_o) + b * p.dos[0] + h * p.dos[1] + row * p.dos[2];
  const T* k = static_cast<const T*>(p.k) + b * p.ks[0] + h * p.ks[1];
  const T* v = static_cast<const T*>(p.v) + b * p.vs[0] + h * p.vs[1];
  T* dq = static_cast<T*>(p.dq) + b * p.dqs[0] + h * p.dqs[1] + row * p.dqs[2];
  const int64_t stat = (static_cast<int64_t>(b) * p.H + h) * p.Nq + row;
  const float L = p.lse[stat], delta = p.delta[stat];

  float* qrow = rows_smem + warp * 2 * p.D;
  float* dorow = qrow + p.D;
  for (int d = lane; d < p.D; d += 32) {
    qrow[d] = to_f32(q[d * p.qs[3]]) * p.scale_log2;
    dorow[d] = to_f32(d_o[d * p.dos[3]]);
  }
  __syncwarp();

  float acc[kSimtMaxD / 32];
#pragma unroll
  for (int i = 0; i < kSimtMaxD / 32; ++i) acc[i] = 0.f;
  int kv_end = p.Nkv;
  if (p.causal) kv_end = min(kv_end, row + 1);

  for (int j0 = 0; j0 < kv_end; j0 += 32) {
    const int j = j0 + lane;
    float ds = 0.f;
    if (j < kv_end) {
      const T* kr = k + j * p.ks[2];
      const T* vr = v + j * p.vs[2];
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < p.D; ++d) {
        s = fmaf(qrow[d], to_f32(kr[d * p.ks[3]]), s);
        dp = fmaf(dorow[d], to_f32(vr[d * p.vs[3]]), dp);
      }
      ds = exp2f(s - L) * (dp - delta);
    }
    const int cnt = min(32, kv_end - j0);
    for (int jj = 0; jj < cnt; ++jj) {
      const float w = __shfl_sync(0xffffffffu, ds, jj);
      const T* kr = k + (j0 + jj) * p.ks[2];
#pragma unroll
      for (int i = 0; i < kSimtMaxD / 32; ++i) {
        const int d = lane + 32 * i;
        if (d < p.D) acc[i] = fmaf(w, to_f32(kr[d * p.ks[3]]), acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kSimtMaxD / 32; ++i) {
    const int d = lane + 32 * i;
    if (d < p.D) dq[d * p.dqs[3]] = from_f32<T>(acc[i] * p.scale);
  }
}

template <typename T>
__global__ void __launch_bounds__(kSimtWarps * 32) fa_bwd_simt_dkv_kernel(const SimtBwdParams p) {
  extern __shared__ float rows_smem[];  // [kSimtWarps][2][D]: k_j * c, v_j
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int col = blockIdx.x * kSimtWarps + warp;  // key row
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  if (col >= p.Nkv) return;

  const T* q = static_cast<const T*>(p.q) + b * p.qs[0] + h * p.qs[1];
  const T* d_o = static_cast<const T*>(p.d_o) + b * p.dos[0] + h * p.dos[1];
  const T* k = static_cast<const T*>(p.k) + b * p.ks[0] + h * p.ks[1] + col * p.ks[2];
  const T* v = static_cast<const T*>(p.v) + b * p.vs[0] + h * p.vs[1] + col * p.vs[2];
  T* dk = static_cast<T*>(p.dk) + b * p.dks[0] + h * p.dks[1] + col * p.dks[2];
  T* dv = static_cast<T*>(p.dv) + b * p.dvs[0] + h * p.dvs[1] + col * p.dvs[2];
  const float* lse = p.lse + (static_cast<int64_t>(b) * p.H + h) * p.Nq;
  const float* delta = p.delta + (static_cast<int64_t>(b) * p.H + h) * p.Nq;

  float* krow = rows_smem + warp * 2 * p.D;
  float* vrow = krow + p.D;
  for (int d = lane; d < p.D; d += 32) {
    krow[d] = to_f32(k[d * p.ks[3]]) * p.scale_log2;
    vrow[d] = to_f32(v[d * p.vs[3]]);
  }
  __syncwarp();

  float acc_k[kSimtMaxD / 32], acc_v[kSimtMaxD / 32];
#pragma unroll
  for (int i = 0; i < kSimtMaxD / 32; ++i) acc_k[i] = acc_v[i] = 0.f;
  const int i_begin = p.causal ? col : 0;  // col > row is masked: query rows below the key index see nothing

  for (int i0 = i_begin; i0 < p.Nq; i0 += 32) {
    const int i = i0 + lane;
    float pw = 0.f, ds = 0.f;
    if (i < p.Nq) {
      const T* qr = q + i * p.qs[2];
      const T* dor = d_o + i * p.dos[2];
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < p.D; ++d) {
        s = fmaf(krow[d], to_f32(qr[d * p.qs[3]]), s);
        dp = fmaf(vrow[d], to_f32(dor[d * p.dos[3]]), dp);
      }
      pw = exp2f(s - lse[i]);
      ds = pw * (dp - delta[i]);
    }
    const int cnt = min(32, p.Nq - i0);
    for (int ii = 0; ii < cnt; ++ii) {
      const float wp = __shfl_sync(0xffffffffu, pw, ii);
      const float wd = __shfl_sync(0xffffffffu, ds, ii);
      const T* qr = q + (i0 + ii) * p.qs[2];
      const T* dor = d_o + (i0 + ii) * p.dos[2];
#pragma unroll
      for (int x = 0; x < kSimtMaxD / 32; ++x) {
        const int d = lane + 32 * x;
        if (d < p.D) {
          acc_v[x] = fmaf(wp, to_f32(dor[d * p.dos[3]]), acc_v[x]);
          acc_k[x] = fmaf(wd, to_f32(qr[d * p.qs[3]]), acc_k[x]);
        }
      }
    }
  }
#pragma unroll
  for (int x = 0; x < kSimtMaxD / 32; ++x) {
    const int d = lane + 32 * x;
    if (d < p.D) {
      dk[d * p.dks[3]] = from_f32<T>(acc_k[x] * p.scale);
      dv[d * p.dvs[3]] = from_f32<T>(acc_v[x]);
    }
  }
}

}  // namespace fa
