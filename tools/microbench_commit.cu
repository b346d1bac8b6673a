// Does tcgen05.commit block the issuing thread?  How deep is the tcgen05.mma queue?  (sm_100a)
// One thread issues G back-to-back UMMAs (M=128, N=128, K=16, SS form), then a commit, then polls the mbarrier;
// clock64 is read after the last MMA (t1), after the commit (t2), after one more trivial instruction (t3) and
// when the barrier flips (t4).  If the queue is deeper than G and the commit is asynchronous, t1, t2, t3 are
// small and only t4 grows with G (64 cycles per MMA).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I flash-attention-v2-rdna3-minimal_b200/csrc \
//        -o tools/microbench_commit.bin tools/microbench_commit.cu
#include <cstdio>
#include <vector>

#include "ptx.cuh"

using namespace fa;

template <int G, int MIDWAIT>
__global__ void __launch_bounds__(128, 1) commit_probe(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = smem_u32(smem + 65536);
  for (int i = threadIdx.x; i < 131072 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar[0]), 1);
    mbar_init(smem_u32(&bar[1]), 1);
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
    constexpr uint32_t idesc = make_idesc_f16(128, 128, false, false, false);
    uint64_t da[8], db[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
      da[k] = make_smem_desc_sw128(sA + off, 16, 1024);
      db[k] = make_smem_desc_sw128(sB + off, 16, 1024);
    }
    long long acc[5] = {0, 0, 0, 0, 0};
    uint32_t phase = 0;
    for (int rep = 0; rep < 64; ++rep) {
      long long t0 = clock64();
#pragma unroll
      for (int g = 0; g < G; ++g) umma_ss(tmem + 256 + ((g >> 3) & 1) * 128, da[g & 7], db[g & 7], idesc, (g & 7) > 0);
      long long t1 = clock64();
      tc_commit(smem_u32(&bar[0]));
      long long t2 = clock64();
      if (MIDWAIT) {  // a second group behind the commit: does IT block on the first group?
#pragma unroll
        for (int g = 0; g < 8; ++g) umma_ss(tmem + ((g >> 3) & 1) * 128, da[g & 7], db[g & 7], idesc, (g & 7) > 0);
      }
      long long t3 = clock64();
      mbar_wait(smem_u32(&bar[0]), phase, 1);
      long long t4 = clock64();
      if (MIDWAIT) {
        tc_commit(smem_u32(&bar[1]));
        mbar_wait(smem_u32(&bar[1]), phase, 2);
      }
      phase ^= 1;
      if (rep >= 8) {
        acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2; acc[3] += t4 - t3; acc[4] += t4 - t0;
      }
    }
    for (int i = 0; i < 5; ++i) out[i] = acc[i] / 56;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int G, int MIDWAIT>
void run(long long* d_out) {
  const int smem = 131072 + 1024;
  cudaFuncSetAttribute(commit_probe<G, MIDWAIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  commit_probe<G, MIDWAIT><<<1, 128, smem>>>(d_out);
  cudaDeviceSynchronize();
  long long h[5];
  cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
  printf("G=%2d second_group=%d : issue %5lld  commit %5lld  %s %5lld  wait-for-barrier %5lld  total %5lld cycles  [%s]\n", G, MIDWAIT,
         h[0], h[1], MIDWAIT ? "8 more MMAs" : "(nothing)  ", h[2], h[3], h[4], cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  run<1, 0>(d_out); run<2, 0>(d_out); run<4, 0>(d_out); run<8, 0>(d_out); run<16, 0>(d_out); run<32, 0>(d_out);
  run<8, 1>(d_out); run<16, 1>(d_out);
  return 0;
}
