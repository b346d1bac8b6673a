/*
 * fa_fwd_sm100_test.h - test and benchmarking hooks of libfa_fwd_sm100.so.
 *
 * NOT part of the drop-in boundary (include/fa_fwd_sm100.h): nothing a caller of the reference's
 * rocwmma_fattn extension needs lives here.  tests/ and tools/ use these to force a kernel, to switch
 * optional mechanisms off for A/B runs, and to run the tensor-core plumbing probes (the counterpart of the
 * reference's gemm_test/ micro-kernels).
 */
#ifndef FA_FWD_SM100_TEST_H_
#define FA_FWD_SM100_TEST_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* FA_KERNEL_WIDE at head dims 193..256 runs on CTA pairs (thread-block cluster of two,
 * tcgen05 cta_group::2: each SM fetches half of every K/V tile) unless disabled here (test / benchmarking
 * hook).  Returns the previous setting. */
int fa_set_wide_pairs(int enable);

/* Programmatic dependent launch of the forward kernels (on by default; FA_NO_PDL=1 in the environment starts
 * with it off): each launch may become resident under its predecessor's tail and does its prologue there;
 * the kernels order their global-memory accesses with griddepcontrol.wait.  Returns the previous setting. */
int fa_set_pdl(int enable);

/* Force a kernel (FA_KERNEL_*) for subsequent calls in this process; FA_KERNEL_AUTO restores the
 * heuristic.  Test / benchmarking hook.  Returns the previous setting. */
int fa_set_kernel(int kernel);

/* Backward kernel selection (test / benchmarking hook).  Head dims <= 128: 0 = automatic (the pipelined, warp-specialised
 * kernel fa_bwd_ws.cuh), 1 = the serial kernel fa_bwd_tc.cuh (round 1), 2 = fa_bwd_ws.cuh.  Head dims 129..256 run the
 * three-launch tcgen05 kernel fa_bwd_wide.cuh unless 3 is set, which forces the generic CUDA-core kernels (fa_bwd_simt.cuh)
 * there as well.  Returns the previous setting. */
int fa_set_bwd_kernel(int kernel);

/*
 * UMMA / TMA / TMEM self-test: computes one 128x128x128 product through the same operand paths the
 * attention kernels use (mode 0: A.B^T both K-major; 1: A.B with B MN-major; 2: A from TMEM;
 * 3: A written to smem by threads; 4: A^T.B with A and B MN-major, the backward's dV/dK products;
 * 5: mode 0 plus one more k-step from UN-swizzled K-major [128][16] tiles whose second k-chunk aliases a shared zeros block
 * through the descriptor's LBO - out = A.B^T + bias0[n] + 2 bias1[n], bias0[n] = (n-64)/8, bias1[n] = (n%7)/4).
 * a, b: device [128,128] 16-bit row-major; out: device
 * [128,128] fp32.  lbo/sbo: B-descriptor byte offsets for modes 1-3 (0,0 = the values the kernels
 * use).  Counterpart of the reference's gemm_test/ micro-kernels.
 */
int fa_umma_selftest(const void* a, const void* b, float* out, int dtype, int mode, uint32_t lbo,
                     uint32_t sbo, void* stream);

/*
 * CTA-pair (cluster of 2, tcgen05 cta_group::2) plumbing probe: out[256,128] (fp32) = A[256,128] * B with
 * every operand 16-bit, each CTA of the pair holding its own 128 rows of A / out and half of B.
 *   mode 0  A from shared memory (K-major); b = B as [n=128][k=128] row-major, CTA r takes rows 64r..64r+63
 *   mode 1  A from tensor memory;           b = B as [k=128][n=128] row-major, CTA r takes columns 64r..64r+63
 */
int fa_umma2_selftest(const void* a, const void* b, float* out, int dtype, int mode, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* FA_FWD_SM100_TEST_H_ */
