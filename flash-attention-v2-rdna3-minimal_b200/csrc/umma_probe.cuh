// UMMA / TMA / TMEM unit probes: one CTA computes one 128x128x128 product through exactly the
// operand paths the attention kernels use, so that a descriptor or layout mistake shows up as a
// failed GEMM against torch.matmul instead of as a wrong attention output.
// (Plays the role of the reference's gemm_test/ micro-kernels: /root/reference/gemm_test/kernel.cu.)
//
//   mode 0  D = A * B^T   A,B K-major via TMA (SWIZZLE_128B)            -> the S = Q K^T path
//   mode 1  D = A * B     A K-major via TMA, B MN-major via TMA         -> the O = P V path (SS)
//   mode 2  D = A * B     A packed into TMEM by tcgen05.st, B MN-major  -> the O = P V path (TS)
//   mode 3  D = A * B     A written to smem by threads (manual swizzle) -> P-through-smem variant
//   mode 4  D = A^T * B   A MN-major (stored [k][m]), B MN-major        -> dV = P^T dO, dK = dS^T Q
#pragma once
#include "ptx.cuh"

namespace fa {

// byte offset of element (row, col) of a [128 x 128] 16-bit tile stored as two [128 x 64]
// SWIZZLE_128B blocks (the layout TMA produces with a {64,128} box and UMMA consumes K-major)
__device__ __forceinline__ uint32_t sw128_offset_16bit(int row, int col) {
  const int blk = col >> 6;
  const int chunk = (col & 63) >> 3;  // 16-byte chunk inside the 128-byte row
  return blk * 16384 + row * 128 + (((chunk ^ (row & 7)) << 4) | ((col & 7) << 1));
}

template <bool kBF16>
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tmap_a,
                  const __grid_constant__ CUtensorMap tmap_b, const uint16_t* __restrict__ a_gmem,
                  float* __restrict__ out, int mode, uint32_t lbo_b, uint32_t sbo_b) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;           // 32 KB
  uint8_t* sB = smem + 32768;   // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 65536 + 64);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const uint32_t bar_load = smem_u32(&bars[0]);
  const uint32_t bar_mma = smem_u32(&bars[1]);

  if (tid == 0) {
    mbar_init(bar_load, 1);
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
  const uint32_t tmem = *tmem_slot;

  // mode 5 = mode 0 plus ONE more k-step whose operands are un-swizzled K-major [128 rows][16 columns] tiles written by
  // the threads: A_ext = (1, 2, 0, ...) in every row, B_ext row n = (bias0[n], bias1[n], 0, ...), the second 16-byte
  // k-chunk of both aliased onto a shared block of zeros through the descriptor's LBO: out = A B^T + bias0[n] + 2 bias1[n]
  uint8_t* sExt = smem + 65536 + 1024;  // A_ext 2 KB | B_ext 2 KB | zeros 2 KB
  if (mode == 5) {
    const bool bf = kBF16;
    auto cvt = [&](float x) -> uint32_t {
      return bf ? static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16(x)))
                : static_cast<uint32_t>(__half_as_ushort(__float2half(x)));
    };
    const uint32_t off = (tid >> 3) * 128 + (tid & 7) * 16;
    *reinterpret_cast<uint4*>(sExt + off) = make_uint4(cvt(1.f) | (cvt(2.f) << 16), 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(sExt + 2048 + off) =
        make_uint4(cvt((tid - 64) * 0.125f) | (cvt((tid % 7) * 0.25f) << 16), 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(sExt + 4096 + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  if (mode == 5) mode = 0;
  const bool ext_step = (sExt != nullptr) && (lbo_b == 0xE5u);  // (set by the host for mode 5)

  if (tid == 0) {
    const bool a_tma = (mode == 0 || mode == 1 || mode == 4);
    mbar_arrive_expect_tx(bar_load, a_tma ? 65536 : 32768);
    if (a_tma) {
      tma_load_4d(smem_u32(sA), &tmap_a, bar_load, 0, 0, 0, 0);
      tma_load_4d(smem_u32(sA + 16384), &tmap_a, bar_load, 64, 0, 0, 0);
    }
    tma_load_4d(smem_u32(sB), &tmap_b, bar_load, 0, 0, 0, 0);
    tma_load_4d(smem_u32(sB + 16384), &tmap_b, bar_load, 64, 0, 0, 0);
  }

  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  constexpr uint32_t kColP = 256;

  if (mode == 2) {
    // thread r packs row r of A into TMEM columns [kColP, kColP+64)
    const uint32_t* row = reinterpret_cast<const uint32_t*>(a_gmem + tid * 128);
    uint32_t r[32];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = row[h * 32 + i];
      tmem_st_x32(tmem + lane_base + kColP + h * 32, r);
    }
    tmem_wait_st();
  } else if (mode == 3) {
    const uint4* row = reinterpret_cast<const uint4*>(a_gmem + tid * 128);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      *reinterpret_cast<uint4*>(sA + sw128_offset_16bit(tid, c * 8)) = row[c];
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (tid == 0) {
    mbar_wait(bar_load, 0, 1);
    tc_fence_after();
    const uint32_t a_base = smem_u32(sA);
    const uint32_t b_base = smem_u32(sB);
    const uint32_t idesc = make_idesc_f16(128, 128, kBF16, mode == 4, mode != 0);
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
      const uint32_t a_addr = a_base + (k >> 2) * 16384 + (k & 3) * 32;
      const uint64_t a_desc = (mode == 4) ? make_smem_desc_sw128(a_base + k * 2048, lbo_b, sbo_b)
                                          : make_smem_desc_sw128(a_addr, 16, 1024);
      uint64_t b_desc;
      if (mode == 0) {
        b_desc = make_smem_desc_sw128(b_base + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
      } else {
        b_desc = make_smem_desc_sw128(b_base + k * 2048, lbo_b, sbo_b);
      }
      if (mode == 2) {
        umma_ts(tmem, tmem + kColP + k * 8, b_desc, idesc, k > 0);
      } else {
        umma_ss(tmem, a_desc, b_desc, idesc, k > 0);
      }
    }
    if (ext_step) {
      const uint32_t e = smem_u32(sExt);
      umma_ss(tmem, make_smem_desc_nosw(e, 4096, 128), make_smem_desc_nosw(e + 2048, 2048, 128), idesc, 1);
    }
    tc_commit(bar_mma);
  }

  mbar_wait(bar_mma, 0, 2);
  tc_fence_after();
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    tmem_ld_x32(tmem + lane_base + c * 32, r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) out[tid * 128 + c * 32 + i] = __uint_as_float(r[i]);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace fa
