// tcgen05.mma rate microbenchmark (sm_100a): cycles per UMMA (M=128, K=16) issued back to back by one
// thread, by operand source / layout.  Mode and N are template parameters and every descriptor is
// computed before the timed loop, so the loop is 8 UTCHMMA + a branch (issue cost far below the
// tensor-pipe time).  Operand values do not matter.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I flash-attention-v2-rdna3-minimal_b200/csrc \
//        -o tools/microbench_umma.bin tools/microbench_umma.cu
#include <cstdio>
#include <vector>

#include "ptx.cuh"

using namespace fa;

// MODE 0: SS  A K-major, B K-major        (S = Q K^T)
//      1: TS  A TMEM,    B MN-major       (O += P V as shipped)
//      2: SS  A K-major, B MN-major       (backward dQ = dS K)
//      3: TS  A TMEM,    B K-major        (O += P V with transposed V tiles)
//      4: SS  A MN-major, B MN-major      (backward dV / dK)
template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) umma_rate(long long* out, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = smem_u32(smem + 65536);
  for (int i = threadIdx.x; i < 196608 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    constexpr bool a_mn = (MODE == 4);
    constexpr bool b_mn = (MODE == 1 || MODE == 2 || MODE == 4);
    constexpr uint32_t idesc = make_idesc_f16(128, N, false, a_mn, b_mn);
    uint64_t da[8], db[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
      da[k] = a_mn ? make_smem_desc_sw128(sA + k * 2048, 16384, 1024) : make_smem_desc_sw128(sA + off, 16, 1024);
      db[k] = b_mn ? make_smem_desc_sw128(sB + k * 2048, 16384, 1024) : make_smem_desc_sw128(sB + off, 16, 1024);
    }
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tmem + 256 + (r & 1) * 128;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (MODE == 1 || MODE == 3) {
          umma_ts(d, tmem + k * 8, db[k], idesc, k > 0);
        } else {
          umma_ss(d, da[k], db[k], idesc, k > 0);
        }
      }
    }
    tc_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, 1);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// CTA pair (cluster of 2, cta_group::2): M = 256, each SM holds its 128 rows of A / D and half of B.
//   MODE 0: SS A K-major, B K-major (N/2 rows per CTA)     MODE 1: TS A TMEM, B MN-major (N/2 columns per CTA)
template <int MODE, int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) umma2_rate(long long* out, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = smem_u32(smem + 65536);
  for (int i = threadIdx.x; i < 196608 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc_2cta(smem_u32(&tmem_slot), 512);
    tmem_relinquish_2cta();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (cluster_ctarank() == 0 && threadIdx.x == 0) {
    constexpr bool b_mn = (MODE == 1);
    constexpr uint32_t idesc = make_idesc_f16(256, N, false, false, b_mn);
    uint64_t da[8], db[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      da[k] = make_smem_desc_sw128(sA + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
      // K-major half: N/2 rows, blocks (N/2)*128 bytes apart; MN-major half: 16 k-rows per step, LBO 16 KB
      db[k] = b_mn ? make_smem_desc_sw128(sB + k * 2048, 16384, 1024)
                   : make_smem_desc_sw128(sB + (k >> 2) * (N / 2) * 128 + (k & 3) * 32, 16, 1024);
    }
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tmem + 256 * (MODE == 1 && N == 256 ? 1 : 1) * 0 + (N == 256 ? 256 : 256 + (r & 1) * 128);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (MODE == 1) {
          umma_ts_2cta(d, tmem + k * 8, db[k], idesc, k > 0);
        } else {
          umma_ss_2cta(d, da[k], db[k], idesc, k > 0);
        }
      }
    }
    tc_commit_2cta(smem_u32(&bar), 0b11);
    mbar_wait(smem_u32(&bar), 0, 1);
    long long t1 = clock64();
    out[blockIdx.x / 2] = t1 - t0;
  } else if (threadIdx.x == 0) {
    mbar_wait(smem_u32(&bar), 0, 2);
  }
  tc_fence_before();
  cluster_sync_all();
  if (threadIdx.x < 32) tmem_dealloc_2cta(tmem, 512);
}

template <int MODE, int N>
void run2(const char* name, long long* d_out) {
  const int reps = 4000;
  const int smem = 196608 + 1024;
  cudaFuncSetAttribute(umma2_rate<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma2_rate<MODE, N><<<148, 128, smem>>>(d_out, reps);
  cudaDeviceSynchronize();
  umma2_rate<MODE, N><<<148, 128, smem>>>(d_out, reps);
  cudaDeviceSynchronize();
  std::vector<long long> h(74);
  cudaMemcpy(h.data(), d_out, 74 * 8, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (auto v : h) mx = v > mx ? v : mx;
  const double per = (double)mx / (reps * 8.0);
  printf("%-34s %7.2f cycles per UMMA (M=256 over 2 SMs,N=%3d,K=16) -> %5.0f MAC/clk/SM  [%s]\n", name, per, N,
         128.0 * N * 16 / per, cudaGetErrorString(cudaGetLastError()));
}

template <int MODE, int N>
void run(const char* name, long long* d_out) {
  const int reps = 4000;
  const int smem = 196608 + 1024;
  cudaFuncSetAttribute(umma_rate<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma_rate<MODE, N><<<148, 128, smem>>>(d_out, reps);
  cudaDeviceSynchronize();
  umma_rate<MODE, N><<<148, 128, smem>>>(d_out, reps);
  cudaDeviceSynchronize();
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), d_out, 148 * 8, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (auto v : h) mx = v > mx ? v : mx;
  const double per = (double)mx / (reps * 8.0);
  printf("%-34s %7.2f cycles per UMMA (M=128,N=%3d,K=16) -> %5.0f MAC/clk/SM  [%s]\n", name, per, N,
         128.0 * N * 16 / per, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 148 * 8);
  run<0, 128>("SS A K-major  B K-major", d_out);
  run<0, 64>("SS A K-major  B K-major", d_out);
  run<1, 128>("TS A TMEM     B MN-major", d_out);
  run<1, 64>("TS A TMEM     B MN-major", d_out);
  run<3, 128>("TS A TMEM     B K-major", d_out);
  run<3, 64>("TS A TMEM     B K-major", d_out);
  run<2, 128>("SS A K-major  B MN-major", d_out);
  run<4, 128>("SS A MN-major B MN-major", d_out);
  run2<0, 128>("2CTA SS A K-major B K-major", d_out);
  run2<0, 256>("2CTA SS A K-major B K-major", d_out);
  run2<1, 128>("2CTA TS A TMEM  B MN-major", d_out);
  run2<1, 256>("2CTA TS A TMEM  B MN-major", d_out);
  return 0;
}
