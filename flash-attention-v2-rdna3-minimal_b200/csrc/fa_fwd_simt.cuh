// Generic CUDA-core forward kernel: any head dim up to 1024, any strides, fp32 math.
// It exists for the shapes the tcgen05 kernels do not cover (head dim > 128, head dim not a
// multiple of 8, unaligned base pointers) and as an on-device cross-check of the tensor-core
// kernels.  It is a CUDA kernel, not a CPU fallback; it is never selected for the BASELINE configs.
//
// One warp owns one query row at a time.  Keys are visited 32 at a time: lane j computes the full
// dot product for key (j0 + j); the 32 probabilities are then broadcast with shuffles while every
// lane accumulates its strided slice of the output row.
// Semantics follow /root/reference/rocwmma_fattn/kernel_fp16.cu:381-543 (online softmax in base 2,
// top-left causal mask `col > row`, L = m + log2(l)).
#pragma once
#include "ptx.cuh"

namespace fa {

template <typename T>
__device__ __forceinline__ float to_f32(T x);
template <>
__device__ __forceinline__ float to_f32<__half>(__half x) {
  return __half2float(x);
}
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 x) {
  return __bfloat162float(x);
}
template <typename T>
__device__ __forceinline__ T from_f32(float x);
template <>
__device__ __forceinline__ __half from_f32<__half>(float x) {
  return __float2half_rn(x);
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float x) {
  return __float2bfloat16_rn(x);
}

struct SimtParams {
  const void* q;
  const void* k;
  const void* v;
  void* o;
  float* lse;  // [B,H,Nq] contiguous, may be null
  int B, H, Nq, Nkv, D;
  int64_t qs[4], ks[4], vs[4], os[4];  // element strides in logical order (b,h,n,d)
  int causal;
  float scale_log2;  // scale * log2(e)
};

constexpr int kSimtWarps = 4;
constexpr int kSimtMaxD = 1024;

template <typename T>
__global__ void __launch_bounds__(kSimtWarps * 32) fa_fwd_simt_kernel(const SimtParams p) {
  extern __shared__ float q_smem[];  // [kSimtWarps][D]
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kSimtWarps + warp;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  if (row >= p.Nq) return;

  const T* q = static_cast<const T*>(p.q) + b * p.qs[0] + h * p.qs[1] + row * p.qs[2];
  const T* k = static_cast<const T*>(p.k) + b * p.ks[0] + h * p.ks[1];
  const T* v = static_cast<const T*>(p.v) + b * p.vs[0] + h * p.vs[1];
  T* o = static_cast<T*>(p.o) + b * p.os[0] + h * p.os[1] + row * p.os[2];

  float* qrow = q_smem + warp * p.D;
  for (int d = lane; d < p.D; d += 32) qrow[d] = to_f32(q[d * p.qs[3]]) * p.scale_log2;
  __syncwarp();

  float acc[kSimtMaxD / 32];
#pragma unroll
  for (int i = 0; i < kSimtMaxD / 32; ++i) acc[i] = 0.f;
  float m = -INFINITY, l = 0.f;

  int kv_end = p.Nkv;
  if (p.causal) kv_end = min(kv_end, row + 1);

  for (int j0 = 0; j0 < kv_end; j0 += 32) {
    const int j = j0 + lane;
    float s = -INFINITY;
    if (j < kv_end) {
      const T* kr = k + j * p.ks[2];
      float dot = 0.f;
      for (int d = 0; d < p.D; ++d) dot = fmaf(qrow[d], to_f32(kr[d * p.ks[3]]), dot);
      s = dot;
    }
    float tile_max = s;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
      tile_max = fmaxf(tile_max, __shfl_xor_sync(0xffffffffu, tile_max, off));
    const float m_new = fmaxf(m, tile_max);
    const float alpha = exp2f(m - m_new);
    const float pj = (j < kv_end) ? exp2f(s - m_new) : 0.f;
    float psum = pj;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, off);
    l = l * alpha + psum;
    m = m_new;
#pragma unroll
    for (int i = 0; i < kSimtMaxD / 32; ++i) acc[i] *= alpha;
    const int cnt = min(32, kv_end - j0);
    for (int jj = 0; jj < cnt; ++jj) {
      const float pb = __shfl_sync(0xffffffffu, pj, jj);
      const T* vr = v + (j0 + jj) * p.vs[2];
#pragma unroll
      for (int i = 0; i < kSimtMaxD / 32; ++i) {
        const int d = lane + 32 * i;
        if (d < p.D) acc[i] = fmaf(pb, to_f32(vr[d * p.vs[3]]), acc[i]);
      }
    }
  }

  const float inv_l = 1.f / l;
#pragma unroll
  for (int i = 0; i < kSimtMaxD / 32; ++i) {
    const int d = lane + 32 * i;
    if (d < p.D) o[d * p.os[3]] = from_f32<T>(acc[i] * inv_l);
  }
  if (p.lse != nullptr && lane == 0) {
    p.lse[(static_cast<int64_t>(b) * p.H + h) * p.Nq + row] = m + log2f(l);
  }
}

}  // namespace fa
