// Inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM).
// Everything here is a thin, single-instruction wrapper so the kernels read as a sequence of
// hardware operations.  No CUTLASS / CuTe dependency.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fa {

// ---------------------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// shared-space scalar accesses (a generic `float*` into shared memory costs a 64-bit address and an ST.E / LD.E)
__device__ __forceinline__ void st_shared_f32(uint32_t saddr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
  return v;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------------------------------------
// programmatic dependent launch (PDL).  A kernel launched with the programmatic-stream-serialization
// attribute may start while the previous kernel on the stream is still running; everything before
// pdl_wait() (barrier init, TMEM allocation, descriptor + L2 prefetch) then overlaps that kernel's tail.
// pdl_wait() returns once every prerequisite grid has completed and its memory operations are visible;
// it is a no-op for a normally launched kernel.  pdl_launch_dependents() lets the NEXT kernel's CTAs be
// scheduled as soon as every CTA of this grid has issued it (or exited).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// packed fp32 pairs (sm_100: FFMA2 / FADD2 / FMUL2 issue two fp32 operations per instruction)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1,
                                      float c0, float c1) {
  asm("{\n\t"
      ".reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %5};\n\t"
      "mov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t"
      "}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t"
      ".reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t"
      "}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fsub2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t"
      ".reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %5};\n\t"
      "sub.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t"
      "}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fmul2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t"
      ".reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t"
      "}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// 2^x for a pair of values on the FMA / ALU pipes instead of the MUFU (which sustains only 16
// ex2 per clock per SM and is the co-bottleneck of the forward at head dim 128).  Cody-Waite:
// n = round(x) through the 1.5 * 2^23 magic constant, f = x - n in [-0.5, 0.5], 2^f by a degree-3
// minimax polynomial (relative error 7.5e-5, below the half-ulp of fp16), then n is added to the
// exponent field.  Inputs are clamped to >= -120 (result ~1e-36, i.e. zero after the 16-bit
// rounding of P); inputs must be <= 127.
__device__ __forceinline__ void ex2_fma2(float& x0, float& x1) {
  constexpr float kMagic = 12582912.f;  // 1.5 * 2^23
  constexpr float k0 = 0.9999280571937561f, k1 = 0.6932609677314758f, k2 = 0.2426111251115799f,
                  k3 = 0.0551716685295105f;
  x0 = fmaxf(x0, -120.f);
  x1 = fmaxf(x1, -120.f);
  float t0, t1, n0, n1, f0, f1, p0, p1;
  fadd2(t0, t1, x0, x1, kMagic, kMagic);
  fadd2(n0, n1, t0, t1, -kMagic, -kMagic);
  fsub2(f0, f1, x0, x1, n0, n1);
  ffma2(p0, p1, f0, f1, k3, k3, k2, k2);
  ffma2(p0, p1, p0, p1, f0, f1, k1, k1);
  ffma2(p0, p1, p0, p1, f0, f1, k0, k0);
  x0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  x1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

// make mbarrier.init visible to the async proxy (TMA / tcgen05.commit)
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t tx_bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx_bytes)
               : "memory");
}

// FA_TRY_WAIT_HINT_NS: suspend-time hint of mbarrier.try_wait.  A waiting thread sleeps in hardware
// until the phase completes or the hint expires instead of returning to the polling loop every few
// dozen cycles.  Measured on B200 (tools/ab_bench.py): 0x989680 (the value CUTLASS uses) is 1-2 % SLOWER
// than plain polling for this kernel, burst and sustained, so the default is 0 = no hint.
#ifndef FA_TRY_WAIT_HINT_NS
#define FA_TRY_WAIT_HINT_NS 0
#endif
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
#if FA_TRY_WAIT_HINT_NS > 0
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(FA_TRY_WAIT_HINT_NS)
      : "memory");
#else
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
#endif
  return ok != 0;
}

#ifndef FA_SYNC_TIMEOUT_NS
#define FA_SYNC_TIMEOUT_NS 4000000000ull  // a deadlock becomes a trap (launch failure), never a hang
#endif

// Non-blocking probe of a phase (mbarrier.test_wait never suspends the thread).
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

// FA_TEST_WAIT_FIRST: probe with the non-blocking test_wait before falling into the try_wait loop.  Most waits
// of the MMA-issuing thread find their phase complete; the plain probe is a little cheaper than a try_wait that
// may suspend: +0.5 % at N=16384 in two A/B pairs (1427 / 1426 vs 1420 / 1419 TFLOPS burst, round 2).
#ifndef FA_TEST_WAIT_FIRST
#define FA_TEST_WAIT_FIRST 1
#endif

// Wait for the phase with the given parity to complete.  `tag` identifies the wait site in the
// deadlock report.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0) {
#if FA_TEST_WAIT_FIRST
  if (mbar_test_wait(bar, parity)) return;
#endif
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (((++spins) & 0x3ff) == 0 && globaltimer_ns() - t0 > FA_SYNC_TIMEOUT_NS) {
#ifdef FA_DEBUG_SYNC
      printf("[fa] mbarrier timeout: block (%d,%d,%d) thread %d tag %d parity %u\n", blockIdx.x,
             blockIdx.y, blockIdx.z, threadIdx.x, tag, parity);
#endif
      __trap();
    }
  }
}

// Whole-warp wait: lane 0 polls, the other lanes park at the warp barrier (32x fewer polls of the
// barrier word; __syncwarp orders the lanes' later accesses after lane 0's acquire).
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int tag = 0) {
#ifdef FA_WARP_WAIT
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity, tag);
  __syncwarp();
#else
  mbar_wait(bar, parity, tag);
#endif
}

// ---------------------------------------------------------------------------------------------
// proxies / fences
// ---------------------------------------------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (TMA store / UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// 4-D tiled prefetch global -> L2 (no shared-memory destination, no completion to wait for).  L2 is the
// coherence point of global memory, so a prefetch issued before pdl_wait() can never make a later load
// observe stale data; it only hides the HBM latency of the first tiles.
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// 4-D tiled load global -> smem, completion on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst_smem), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// CTA-pair form: the data lands in THIS CTA's shared memory, the transaction bytes are counted on a
// barrier that may live in the peer CTA (`bar_cluster` is a shared::cluster address, see mapa_shared)
__device__ __forceinline__ void tma_load_4d_2cta(uint32_t dst_smem, const CUtensorMap* map,
                                                 uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst_smem), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// 4-D tiled store smem -> global (bulk-group completion); out-of-bounds elements are not written
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src_smem, int c0,
                                             int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      ::"l"(map), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// 3-D tiled reduce-add smem -> global (the element type, fp32 here, comes from the tensor map):
// global[box] += smem[box], performed by the L2 atomic units on whole 128-byte lines
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, uint32_t src_smem, int c0,
                                                  int c1, int c2) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
      ::"l"(map), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// wait until the smem source of all committed bulk stores has been read
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// wait until at most one committed bulk group still has its smem source unread
__device__ __forceinline__ void tma_store_wait_read_1() {
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation
// ---------------------------------------------------------------------------------------------
// Whole warp must execute.  Writes the TMEM base address to *dst_smem.
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05: fences / waits / commit
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// mbarrier arrives (count 1) once every tcgen05.mma previously issued by this thread has completed.
// Implies tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// CTA pairs (thread-block cluster of 2, tcgen05 cta_group::2): two SMs work on one M = 256 tile, each
// holding its own 128 rows of A / D and HALF of B, so a B tile is fetched and read once per pair.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `saddr` (a shared::cta address of this CTA) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
#ifndef FA_PAIR_ARRIVE_RELEASE
#define FA_PAIR_ARRIVE_RELEASE 0
#endif
// Arrival on a barrier of (possibly) the peer CTA.  What the arrival publishes here is tensor-memory
// state, ordered by tcgen05.wait::st + tcgen05.fence::before_thread_sync on this side and
// tcgen05.fence::after_thread_sync on the waiter's, not generic-proxy memory, so the relaxed form is
// enough; the release.cluster form is kept behind FA_PAIR_ARRIVE_RELEASE for comparison.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
#if FA_PAIR_ARRIVE_RELEASE
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#else
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#endif
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// completion of every MMA issued so far by this thread -> one arrival on the barrier at the same
// shared-memory offset in each CTA of `cta_mask`
__device__ __forceinline__ void tc_commit_2cta(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(cta_mask)
      : "memory");
}
// issued by the leader CTA only; descriptors / TMEM addresses are offsets valid in both CTAs
__device__ __forceinline__ void umma_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ts_2cta(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05: descriptors
// ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (sm_100 "version 1"), SWIZZLE_128B.
//   bits [ 0,14) start address >> 4        bits [16,30) leading-dim byte offset >> 4
//   bits [32,46) stride-dim byte offset>>4 bits [46,48) version = 1
//   bits [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// K-major operand WITHOUT swizzle: 8-row x 16-byte core matrices (row r of a core matrix at +16 r bytes); `lbo_bytes` is the
// distance between the core matrices of one row group along K, `sbo_bytes` between row groups (8 rows) along M / N.
// Used for the one-k-step "bias" operands of the backward (a ones column against a column of -L_i / -D_i); pinned on the
// hardware by fa_umma_selftest mode 5.
__device__ __forceinline__ uint64_t make_smem_desc_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

// The same descriptor as two 32-bit halves.  The high word depends only on the layout (SBO,
// version, swizzle), so it is a compile-time constant; the low word is the 14-bit address field
// plus LBO << 16, and stepping the operand by `bytes` inside one tile is `lo + (bytes >> 4)` (no
// carry out of the address field for smem addresses < 256 KB).  Keeping both halves in uniform
// registers makes one tcgen05.mma cost a single UIADD3 + UTCHMMA on the issuing thread.
__host__ __device__ constexpr uint32_t smem_desc_hi_sw128(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
}
__host__ __device__ constexpr uint32_t smem_desc_hi_nosw(uint32_t sbo_bytes) {  // high word of make_smem_desc_nosw
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14);
}
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}

// Instruction descriptor for kind::f16 (fp16 / bf16 operands, fp32 accumulate).
//   [4,6) D format (1 = f32)   [7,10) A format   [10,13) B format  (0 = f16, 1 = bf16)
//   [15] A major (0 = K)       [16] B major (0 = K, 1 = MN)
//   [17,23) N >> 3             [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, bool is_bf16, bool a_mn_major,
                                                      bool b_mn_major) {
  return (1u << 4) | ((is_bf16 ? 1u : 0u) << 7) | ((is_bf16 ? 1u : 0u) << 10) |
         ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// tcgen05.mma  (issued by ONE thread)
// ---------------------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// split-descriptor forms (see smem_desc_lo / smem_desc_hi_sw128)
__device__ __forceinline__ void umma_ss2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi,
                                         uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo,
                                         uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// the same for a CTA pair (issued by the leader CTA only)
__device__ __forceinline__ void umma_ss2_2cta(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi,
                                              uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ts2_2cta(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo,
                                              uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05.ld / tcgen05.st, shape 32x32b: thread t of the warp touches TMEM lane
// (32 * (warp_id % 4) + t) and N consecutive 32-bit columns.
// ---------------------------------------------------------------------------------------------
#define FA_R4(a, i) "%" #i
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                 "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
                 "r"(r[7])
               : "memory");
}

__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
        "r"(r[15])
      : "memory");
}

__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
        "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
        "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]),
        "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
#undef FA_R4

__device__ __forceinline__ void st_shared_u16(uint32_t saddr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"(static_cast<uint16_t>(v)) : "memory");
}

// ---------------------------------------------------------------------------------------------
// 16-bit packing
// ---------------------------------------------------------------------------------------------
template <bool kBF16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if constexpr (kBF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}

}  // namespace fa
