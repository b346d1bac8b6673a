/*
 * fa_fwd_sm100.h — C ABI of the B200 (sm_100a) Flash-Attention-2 forward path.
 *
 * This is the drop-in boundary for the forward half of the reference's native extension:
 *
 *   reference interface                                   replaced by
 *   ---------------------------------------------------   ---------------------------------
 *   rocwmma_fattn/host.cpp:30-45   forward(q,k,v,Br,Bc,   fa_fwd_sm100()
 *        causal,scale,permute_NH)  (pybind11, dtype
 *        dispatch fp16 / bf16)
 *   rocwmma_fattn/kernel_fp16.cu:744-876  forward_fp16()  fa_fwd_sm100(dtype = FA_DTYPE_F16)
 *   rocwmma_fattn/kernel_bf16.cu:802-941  forward_bf16()  fa_fwd_sm100(dtype = FA_DTYPE_BF16)
 *   kernel_fp16.cu:854-863  printf-only launch errors     return code + fa_last_error()
 *
 * Differences from the reference, by design:
 *   - plain pointers, sizes and element strides; no torch types cross this boundary;
 *   - the caller owns every buffer (o, lse); the library allocates no tensors and makes no hidden
 *     padded / contiguous copies: both [B,H,N,D] and [B,N,H,D] are consumed in place through the
 *     stride arguments (the reference's `permute_NH` flag is subsumed by the strides);
 *   - Br/Bc tile sizes are internal to the kernels;
 *   - the launch goes to the CUDA stream the caller passes (the reference uses the legacy default
 *     stream) and is asynchronous.
 *
 * All functions are thread-safe; errors are reported per thread.
 */
#ifndef FA_FWD_SM100_H_
#define FA_FWD_SM100_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FA_ABI_VERSION 5

/* element types (reference: host.cpp:32-44 dispatches on torch::kFloat16 / torch::kBFloat16) */
#define FA_DTYPE_F16 0
#define FA_DTYPE_BF16 1

/* return codes */
#define FA_OK 0
#define FA_ERR_INVALID_ARG 1   /* bad shape / dtype / stride / null pointer */
#define FA_ERR_UNSUPPORTED 2   /* valid request this build cannot serve (e.g. head dim > 1024) */
#define FA_ERR_CUDA 3          /* a CUDA runtime / driver call failed */
#define FA_ERR_NO_DEVICE 4     /* no sm_100 device is current */

/* kernel selectors for fa_set_kernel() / fa_select_kernel() */
#define FA_KERNEL_AUTO 0
#define FA_KERNEL_SIMT 1       /* CUDA-core kernel, any head dim <= 1024 */
#define FA_KERNEL_TC1 2        /* tcgen05, one 128-row Q tile per CTA, P through TMEM */
/* 3 was FA_KERNEL_TC1_PSMEM (P through shared memory): removed in ABI 5, the value stays reserved */
#define FA_KERNEL_WS 4         /* tcgen05, warp-specialised, two Q tiles per CTA (the fast path) */
#define FA_KERNEL_SK 5         /* FA_KERNEL_WS made persistent: one CTA per SM, work split evenly over
                                  (query block, KV tile) items, a query block shared by any number of
                                  CTAs; non-causal, Nq % 256 == 0; falls back to FA_KERNEL_WS otherwise */
#define FA_KERNEL_WIDE 6       /* tcgen05, one Q tile per CTA with the score tile double-buffered: head dims
                                  129..256 (on CTA pairs above 192), and small or short-causal problems at head
                                  dims <= 128 */
#define FA_KERNEL_WS2 7        /* FA_KERNEL_WS on CTA pairs (cluster of two, cta_group::2): each SM fetches half of
                                  every K/V tile; non-causal (causal requests run FA_KERNEL_WS) */
/* 8 was FA_KERNEL_QUAD2 (four threads per query row, measured slower): removed in ABI 5, reserved */
#define FA_KERNEL_WS3 9        /* FA_KERNEL_WS2 with P in spare tensor memory, so that S_t(j+1) is issued ahead of
                                  O_t += P_t(j) V: head dims <= 64 (the default there for all but small or short-KV
                                  problems); a request at head dims 65..128 runs FA_KERNEL_WS */

/*
 * Attention forward on device buffers.  Replaces host.cpp:30-45 `forward` + kernel_*.cu
 * `forward_fp16/bf16`.
 *
 *   q        [B,H,Nq ,D] logical; element strides q_strides[4] in the order (b,h,n,d)
 *   k, v     [B,H,Nkv,D] logical; strides k_strides / v_strides
 *   o        [B,H,Nq ,D] logical; strides o_strides; written by the kernel (same dtype as q)
 *   lse      optional (may be NULL): contiguous fp32 [B,H,Nq]; receives
 *            L = max_j(s_ij)*log2(e) + log2(sum_j exp(s_ij - max)) with s = scale * q.k, i.e. the
 *            base-2 log-sum-exp the reference stores (kernel_fp16.cu:541-542)
 *   dtype    FA_DTYPE_F16 or FA_DTYPE_BF16 (q, k, v, o all share it)
 *   causal   non-zero: mask col > row (top-left aligned, kernel_fp16.cu:396-412)
 *   scale    softmax scale (the Python layer defaults it to D**-0.5, FlashAttn.py:63-64)
 *   stream   cudaStream_t to launch on (NULL = legacy default stream)
 *
 * The innermost stride (d) of every tensor must be 1.  Returns FA_OK or an error code; the launch
 * is asynchronous with respect to the host.
 */
int fa_fwd_sm100(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H,
                 int Nq, int Nkv, int D, const int64_t q_strides[4], const int64_t k_strides[4],
                 const int64_t v_strides[4], const int64_t o_strides[4], int dtype, int causal,
                 float scale, void* stream);

/*
 * Same computation with HOST buffers (contiguous [B,H,N,D]; pinned memory recommended).  Inputs are
 * staged to the current device head-group by head-group on internal streams so that the
 * host->device copy of group i+1 and the device->host copy of group i-1 overlap the kernel of
 * group i; returns after the last output byte has landed in `o`.  This is the end-to-end path
 * bench.py reports as `e2e`.  `lse` may be NULL.
 */
int fa_fwd_sm100_host(const void* q, const void* k, const void* v, void* o, float* lse, int B,
                      int H, int Nq, int Nkv, int D, int dtype, int causal, float scale);

/*
 * fa_fwd_sm100_host() without the final wait: returns once the copies and kernels are enqueued.  `q`, `k`,
 * `v`, `o` (and `lse`) must stay valid and untouched until fa_host_sync() returns.  Consecutive calls on a
 * device pipeline through two staging sets: the host->device copies of call i+1 run under the last kernel
 * and the device->host copy of call i, so a sequence of calls costs the PCIe time of its bytes plus one tail
 * instead of one tail per call.  An error drains everything in flight before it is returned.
 */
int fa_fwd_sm100_host_async(const void* q, const void* k, const void* v, void* o, float* lse, int B,
                            int H, int Nq, int Nkv, int D, int dtype, int causal, float scale);

/* Wait until every fa_fwd_sm100_host_async() call issued on the current device has delivered its output. */
int fa_host_sync(void);

/*
 * Attention backward on device buffers.  Replaces host.cpp:47-58 `backward` + kernel_*.cu
 * `backward_fp16/bf16` (kernel_fp16.cu:878-1028, kernel_bf16.cu:943-1092) and `bwd_kernel`
 * (kernel_fp16.cu:547-740).
 *
 *   q, k, v, o   the forward's inputs and output (o: [B,H,Nq,D], strides o_strides)
 *   d_o          gradient of the loss w.r.t. o, [B,H,Nq,D] logical, strides do_strides
 *   lse          the forward's base-2 log-sum-exp, contiguous fp32 [B,H,Nq] (fa_fwd_sm100 `lse`)
 *   dq, dk, dv   outputs, shaped / typed like q, k, v; strides dq_strides / dk_strides / dv_strides
 *   dq_accum     caller-owned fp32 workspace: contiguous [B,H,Nq,D] for D <= 128 (zeroed here; dQ is accumulated across key
 *                tiles with fp32 reduce-adds - the reference adds into a 16-bit dQ from several CTAs without atomics,
 *                kernel_fp16.cu:736), any non-null 16-byte-aligned buffer of at least 8 floats otherwise (unused)
 *   delta        caller-owned fp32 workspace, contiguous [B,H,Nq]; receives rowsum(dO o O)
 *                (what the reference recomputes in every CTA, kernel_fp16.cu:605-631)
 *
 * Kernels: D % 8 == 0 with 16-byte aligned pointers and strides runs on the tensor cores - D <= 128: pre-pass, the pipelined
 * kernel fa_bwd_ws.cuh, dQ conversion (3 launches); 128 < D <= 256: pre-pass + fa_bwd_wide.cuh, dQ written directly (D <= 192: one
 * launch for dV and dK together and one for dQ, 3 launches; above: one each, 4 launches).  Everything else the forward accepts (D up to 1024, any alignment) runs
 * the generic CUDA-core kernels (pre-pass + 2 launches).  D > 1024 returns FA_ERR_UNSUPPORTED.  Innermost strides must be 1.
 * All launches go to `stream`, asynchronous with respect to the host.
 */
int fa_bwd_sm100(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                 const float* lse, void* dq, void* dk, void* dv, float* dq_accum, float* delta, int B,
                 int H, int Nq, int Nkv, int D, const int64_t q_strides[4],
                 const int64_t k_strides[4], const int64_t v_strides[4], const int64_t o_strides[4],
                 const int64_t do_strides[4], const int64_t dq_strides[4],
                 const int64_t dk_strides[4], const int64_t dv_strides[4], int dtype, int causal,
                 float scale, void* stream);

/*
 * How fa_fwd_sm100_host() would split the flattened (batch, head) axis of this problem into pipeline chunks:
 * writes up to `cap` chunk sizes (heads per chunk, in order) to `out` and returns the number of chunks (their
 * sizes sum to B*H), or a negative error code.  Pure host logic: usable without a GPU.
 */
int fa_host_plan_chunks(int B, int H, int Nq, int Nkv, int D, int causal, int* out, int cap);

/* Release what the library caches on the current device: the staging buffers and streams of the host-buffer
 * path and the persistent kernel's per-stream workspaces.  Call it only when no launch of this library is in
 * flight and no CUDA graph that captured one will be replayed again.
 *
 * Workspace rule for FA_KERNEL_SK: partial results of query blocks shared by several CTAs go through a
 * workspace owned by the (device, stream) the launch was issued - or captured - on.  Launches on one stream
 * are ordered, so they share it safely; a graph captured on stream S must not be replayed concurrently with
 * other work of this library issued on S or with another graph captured on S (replay such graphs on S, or
 * capture each on its own stream).  A capture on a stream that has no workspace yet uses FA_KERNEL_WS. */
int fa_host_workspace_release(void);

/* Message describing the last error on the calling thread ("" if none). */
const char* fa_last_error(void);

/* FA_ABI_VERSION the library was built with. */
int fa_abi_version(void);

/*
 * Which kernel fa_fwd_sm100() would run for this problem (one of FA_KERNEL_*, never AUTO), or a
 * negative error code.  Pure host logic: usable without a GPU.
 */
int fa_select_kernel(int B, int H, int Nq, int Nkv, int D, const int64_t q_strides[4],
                     const int64_t k_strides[4], const int64_t v_strides[4],
                     const int64_t o_strides[4], int dtype, int causal, float scale);

/* Number of kernel launches issued by this library in this process (all threads). */
uint64_t fa_launch_count(void);

#ifdef __cplusplus
}
#endif

#endif /* FA_FWD_SM100_H_ */
