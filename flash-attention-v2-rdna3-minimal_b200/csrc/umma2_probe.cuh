// CTA-pair tensor-core plumbing probe (cta_group::2): D[256 x 128] = A[256 x 128] * B over a cluster of
// two CTAs, each holding its own 128 rows of A / D and half of B.  It pins down, on the hardware, the
// facts a paired attention kernel depends on and that no document in this image states:
//   mode 0  SS: A K-major from shared memory; B K-major [n][k], CTA r holds rows n in [64r, 64r+64)
//           (the S = Q K^T shape: each CTA of the pair loads half of the keys)
//   mode 1  TS: A from each CTA's own tensor memory; B MN-major [k][n], CTA r holds columns n in
//           [64r, 64r+64) (the O += P V shape: each CTA loads half of the head dim of V)
// Operands are written to shared / tensor memory by the threads themselves (no TMA), the leader CTA's
// thread 0 issues every MMA and one multicast commit releases both CTAs.
#pragma once
#include "umma_probe.cuh"

namespace fa {

template <bool kBF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma2_probe_kernel(const uint16_t* __restrict__ a_gmem, const uint16_t* __restrict__ b_gmem,
                   float* __restrict__ out, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;           // [128 rows][128 k] as two swizzled 64-column blocks (32 KB)
  uint8_t* sB = smem + 32768;   // mode 0: [64 n][128 k] two 8 KB blocks; mode 1: [128 k][64 n] one 16 KB block
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 65536 + 64);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const uint32_t bar_mma = smem_u32(&bars[0]);

  if (tid == 0) {
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc_2cta(smem_u32(tmem_slot), 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  constexpr uint32_t kColP = 256;

  // ---- operands: my 128 rows of A, my half of B
  const uint16_t* a_row = a_gmem + (static_cast<size_t>(rank) * 128 + tid) * 128;
  if (mode == 0) {
    const uint4* row = reinterpret_cast<const uint4*>(a_row);
#pragma unroll
    for (int c = 0; c < 16; ++c) *reinterpret_cast<uint4*>(sA + sw128_offset_16bit(tid, c * 8)) = row[c];
    if (tid < 64) {  // B row n = 64 rank + tid: 128 k values -> two 64-row blocks of 128-byte rows
      const uint4* brow = reinterpret_cast<const uint4*>(b_gmem + (static_cast<size_t>(rank) * 64 + tid) * 128);
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int blk = c >> 3, chunk = c & 7;
        *reinterpret_cast<uint4*>(sB + blk * 8192 + tid * 128 + ((chunk ^ (tid & 7)) << 4)) = brow[c];
      }
    }
  } else {
    const uint32_t* row = reinterpret_cast<const uint32_t*>(a_row);
    uint32_t r[32];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = row[h * 32 + i];
      tmem_st_x32(tmem + lane_base + kColP + h * 32, r);
    }
    tmem_wait_st();
    // B row k = tid: my 64 columns n in [64 rank, 64 rank + 64) = one 128-byte swizzled row
    const uint4* brow = reinterpret_cast<const uint4*>(b_gmem + static_cast<size_t>(tid) * 128 + rank * 64);
#pragma unroll
    for (int chunk = 0; chunk < 8; ++chunk)
      *reinterpret_cast<uint4*>(sB + tid * 128 + ((chunk ^ (tid & 7)) << 4)) = brow[chunk];
  }
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();  // both CTAs' operands are in place (and both barriers initialised)
  tc_fence_after();

  if (rank == 0 && tid == 0) {
    const uint32_t a_base = smem_u32(sA);
    const uint32_t b_base = smem_u32(sB);
    const uint32_t idesc = make_idesc_f16(256, 128, kBF16, false, mode != 0);
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
      if (mode == 0) {
        const uint64_t a_desc = make_smem_desc_sw128(a_base + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
        const uint64_t b_desc = make_smem_desc_sw128(b_base + (k >> 2) * 8192 + (k & 3) * 32, 16, 1024);
        umma_ss_2cta(tmem, a_desc, b_desc, idesc, k > 0);
      } else {
        const uint64_t b_desc = make_smem_desc_sw128(b_base + k * 2048, 16384, 1024);
        umma_ts_2cta(tmem, tmem + kColP + k * 8, b_desc, idesc, k > 0);
      }
    }
    tc_commit_2cta(bar_mma, 0b11);
  }

  mbar_wait(bar_mma, 0, 3);
  tc_fence_after();
  float* my_out = out + (static_cast<size_t>(rank) * 128 + tid) * 128;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    tmem_ld_x32(tmem + lane_base + c * 32, r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) my_out[c * 32 + i] = __uint_as_float(r[i]);
  }

  tc_fence_before();
  cluster_sync_all();  // nobody leaves (or frees tensor memory) while the pair's MMA may still touch it
  if (warp == 0) tmem_dealloc_2cta(tmem, 512);
}

}  // namespace fa
