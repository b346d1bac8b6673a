// C-ABI host side of the B200 Flash-Attention-2 forward path (see include/fa_fwd_sm100.h).
//
// Replaces the reference's native host layer:
//   /root/reference/rocwmma_fattn/host.cpp:30-45          dtype dispatch            -> fa_fwd_sm100(dtype)
//   /root/reference/rocwmma_fattn/kernel_fp16.cu:744-876  forward_fp16 (pad, alloc,
//   /root/reference/rocwmma_fattn/kernel_bf16.cu:802-941  forward_bf16  grid, launch) -> plan + launch below
//
// What is different by design: no tensors are allocated or copied here (unaligned sequence lengths
// are handled by TMA out-of-bounds fill + in-kernel masks instead of F::pad, both memory layouts
// are consumed through TMA strides instead of .contiguous()), errors are returned instead of
// printf'd, and the launch goes to the caller's stream.
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <algorithm>
#include <cmath>
#include <limits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fa_fwd_sm100.h"
#include "../../include/fa_fwd_sm100_test.h"
#include "fa_bwd_simt.cuh"
#include "fa_bwd_tc.cuh"
#include "fa_bwd_ws.cuh"
#include "fa_bwd_wide.cuh"
#include "fa_fwd_simt.cuh"
#include "fa_fwd_tc.cuh"
#include "fa_fwd_ws.cuh"
#include "fa_fwd_sk.cuh"
#include "fa_fwd_wide.cuh"
#include "fa_fwd_wide2.cuh"
#include "fa_fwd_ws2.cuh"
#include "fa_fwd_ws3.cuh"
#include "umma_probe.cuh"
#include "umma2_probe.cuh"

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_forced_kernel{FA_KERNEL_AUTO};
std::atomic<int> g_wide_pairs{1};  // head dims 193..256: CTA-pair kernel (fa_set_wide_pairs)
// programmatic dependent launch for the forward kernels (fa_set_pdl / FA_NO_PDL=1 switch it off for A/B runs)
std::atomic<int> g_pdl{std::getenv("FA_NO_PDL") == nullptr ? 1 : 0};
#ifdef FA_TRACE
unsigned long long* g_trace = nullptr;  // debug builds only (tools/trace_ws.py)
#define FA_TP_TRACE , g_trace
#else
#define FA_TP_TRACE
#endif

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define FA_CUDA_TRY(expr)                                                                   \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      (void)cudaGetLastError();                                                             \
      return fail(FA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));        \
    }                                                                                       \
  } while (0)

// ---------------------------------------------------------------------------------------------
// driver entry point for tensor-map encoding (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) {
      (void)cudaGetLastError();
      return nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 4-D map over a logical [B,H,N,D] 16-bit tensor with element strides st[4] (b,h,n,d), box =
// {64 d-elements (one 128-byte swizzle row), box_rows, 1, 1}, SWIZZLE_128B, zero OOB fill.
int make_map(CUtensorMap* map, const void* base, int B, int H, int N, int D, const int64_t st[4],
             int dtype, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return fail(FA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(N),
                        static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(st[2]) * 2, static_cast<cuuint64_t>(st[1]) * 2,
                           static_cast<cuuint64_t>(st[0]) * 2};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(box_rows), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, dtype == FA_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                   4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf,
             "cuTensorMapEncodeTiled failed (CUresult %d) dims=[%d,%d,%d,%d] strides(elem)=[%lld,%lld,%lld]",
             static_cast<int>(r), D, N, H, B, static_cast<long long>(st[2]),
             static_cast<long long>(st[1]), static_cast<long long>(st[0]));
    return fail(FA_ERR_CUDA, buf);
  }
  return FA_OK;
}

// ---------------------------------------------------------------------------------------------
// problem description + validation + kernel choice (pure host logic)
// ---------------------------------------------------------------------------------------------
struct Problem {
  int B, H, Nq, Nkv, D, dtype, causal;
  float scale;
  int64_t qs[4], ks[4], vs[4], os[4];
};

// A stride of a size-1 dimension is meaningless (torch reports arbitrary values): replace it by a
// value that is always TMA-encodable.
void canon_strides(int64_t st[4], int B, int H, int N, int D) {
  (void)D;
  const int64_t safe = 8;
  if (N == 1) st[2] = safe;
  if (H == 1) st[1] = safe;
  if (B == 1) st[0] = safe;
}

int validate(Problem& p, const int64_t* qs, const int64_t* ks, const int64_t* vs,
             const int64_t* os) {
  if (qs == nullptr || ks == nullptr || vs == nullptr || os == nullptr)
    return fail(FA_ERR_INVALID_ARG, "stride arrays must not be null");
  if (p.B < 1 || p.H < 1 || p.Nq < 1 || p.Nkv < 1 || p.D < 1) {
    char buf[160];
    snprintf(buf, sizeof buf, "sizes must be positive: B=%d H=%d Nq=%d Nkv=%d D=%d", p.B, p.H, p.Nq,
             p.Nkv, p.D);
    return fail(FA_ERR_INVALID_ARG, buf);
  }
  if (p.dtype != FA_DTYPE_F16 && p.dtype != FA_DTYPE_BF16)
    return fail(FA_ERR_INVALID_ARG, "dtype must be FA_DTYPE_F16 or FA_DTYPE_BF16");
  if (!std::isfinite(p.scale)) return fail(FA_ERR_INVALID_ARG, "scale must be finite");
  memcpy(p.qs, qs, sizeof p.qs);
  memcpy(p.ks, ks, sizeof p.ks);
  memcpy(p.vs, vs, sizeof p.vs);
  memcpy(p.os, os, sizeof p.os);
  if (p.qs[3] != 1 || p.ks[3] != 1 || p.vs[3] != 1 || p.os[3] != 1)
    return fail(FA_ERR_INVALID_ARG, "the innermost (head-dim) stride of q, k, v and o must be 1");
  for (int i = 0; i < 3; ++i)
    if (p.qs[i] < 0 || p.ks[i] < 0 || p.vs[i] < 0 || p.os[i] < 0)
      return fail(FA_ERR_INVALID_ARG, "negative strides are not supported");
  if (p.D > fa::kSimtMaxD) return fail(FA_ERR_UNSUPPORTED, "head dim > 1024 is not supported");
  if (p.B > 65535 || p.H > 65535)
    return fail(FA_ERR_UNSUPPORTED, "B and H must be <= 65535 (CUDA grid limit)");
  canon_strides(p.qs, p.B, p.H, p.Nq, p.D);
  canon_strides(p.os, p.B, p.H, p.Nq, p.D);
  canon_strides(p.ks, p.B, p.H, p.Nkv, p.D);
  canon_strides(p.vs, p.B, p.H, p.Nkv, p.D);
  return FA_OK;
}

bool tma_ok_strides(const int64_t st[4]) {
  for (int i = 0; i < 3; ++i) {
    if (st[i] % 8 != 0) return false;                  // 16-byte multiple
    if (st[i] * 2 >= (int64_t(1) << 40)) return false;  // TMA stride limit
  }
  return true;
}

bool sk_eligible(const Problem& p, int n_sm);

// pointer-independent part of the choice
int choose_kernel_shape(const Problem& p) {
  const bool tc = (p.D % 8 == 0) && (p.D <= 256) && (p.scale > 0.f) && tma_ok_strides(p.qs) &&
                  tma_ok_strides(p.ks) && tma_ok_strides(p.vs) && tma_ok_strides(p.os);
  if (!tc) return FA_KERNEL_SIMT;
  if (p.D > 128) return FA_KERNEL_WIDE;  // two Q tiles no longer fit in TMEM: one tile, two S buffers
  if (p.Nq <= fa::kTileM) return FA_KERNEL_TC1;
  const long long tiles128 = static_cast<long long>(p.B) * p.H * ((p.Nq + fa::kTileM - 1) / fa::kTileM);
  // Head dims <= 64, non-causal: tensor memory has 128 spare columns there, so P gets its own region and
  // S_t(j+1) is issued while softmax_t(j) still runs (fa_fwd_ws3.cuh) - the S -> P -> PV -> S chain loses its
  // tensor-core tail.  Measured fp16 H=16 D=64: 841 vs 721 TFLOPS at N=16384, 640 vs 572 at N=4096.
  // (Not for short KV loops - cross-attention with Nkv = 77 is one tile: the pair's cluster start-up then costs
  // more than it saves, 21.0 vs 17.7 us at B=2 H=10 Nq=4096 D=64, tools/bench_sd_shapes.py.)
  // Causal too (the four Q tiles of a pair run in lock step over the tiles the last one needs): 688 vs 619 TFLOPS
  // at N=16384; small causal problems keep the 128-row grain of the one-tile kernel (415 vs 373 at N=4096).
  const long long blocks256 = static_cast<long long>(p.B) * p.H * ((p.Nq + 2 * fa::kTileM - 1) / (2 * fa::kTileM));
  // The persistent kernel first: its cost model (estimate_costs) knows both one-shot alternatives, so at head dims
  // <= 64 it takes the problems that lose a large part of their last round - the Stable-Diffusion shapes: B=2 H=10
  // N=4096 D=64 is 2.16 rounds of CTA pairs, 123 vs 147 us; B=2 H=20 N=1024 D=64 24.9 vs 31.4 us.
  if (sk_eligible(p, 148)) return FA_KERNEL_SK;  // re-checked against the real SM count at launch
  if (p.D <= 64 && tiles128 > 148 && p.Nkv >= 4 * fa::kTileN && (!p.causal || blocks256 >= 2 * 148))
    return FA_KERNEL_WS3;
  // Causal problems whose 256-row blocks all fit in one round: the launch lasts as long as its longest block, so
  // the one-tile arrangement's 128-row grain wins although it moves twice the K/V per Q tile.  Beyond one round
  // the two-tile kernel is ahead again - since round 2 causal launches run longest block first across heads
  // (fa::work_coords), which took N=4096 from 90 to 62 us.  fp16 H=16, wide / ws in us (profiles/r02_sweep_kernels
  // .json): D=128 N=2048 23.7 / 32.0, N=4096 72.3 / 62.0, N=8192 256 / 225; D=64 N=2048 20.7 / 30.8, N=4096 64.8 / 58.6.
  if (p.causal && blocks256 <= 148) return FA_KERNEL_WIDE;
  // Any mask, when every 128-row tile gets its own SM in one round: twice the CTAs of the two-tile kernel
  // and a shorter iteration (graph-timed, fp16 H=16 D=128: N=512 252 vs 179 TFLOPS, N=1024 654 vs 460;
  // at N=2048 - 256 tiles, two rounds - the two-tile kernel is back in front, 1081 vs 875).
  if (static_cast<long long>(p.B) * p.H * ((p.Nq + fa::kTileM - 1) / fa::kTileM) <= 148) return FA_KERNEL_WIDE;
  // (The two-tile kernel on CTA pairs, FA_KERNEL_WS2, was +1 % at N=16384 in round 1; with PDL and the L2
  // prefetch in the one-shot kernel it measures 2 % behind - 1590 vs 1556 us, profiles/r02_sweep_kernels.json -
  // so it is no longer selected automatically.  FA_WS2=1 in the environment restores the round-1 rule.)
  static const bool use_ws2 = std::getenv("FA_WS2") != nullptr;
  if (use_ws2 && !p.causal && p.Nkv >= 8192 &&
      static_cast<long long>(p.B) * p.H * ((p.Nq + 2 * fa::kTileM - 1) / (2 * fa::kTileM)) >= 4 * 148)
    return FA_KERNEL_WS2;
  return FA_KERNEL_WS;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------------------------------------
// tensor-map plan cache: encoding four maps costs a few microseconds, which matters at N = 512
// ---------------------------------------------------------------------------------------------
struct Plan {
  const void *q, *k, *v, *o;
  Problem p;
  int device;
  CUtensorMap mq, mk, mv, mo;
  CUtensorMap mk64;  // K with a 64-key box (CTA-pair kernels: each CTA loads half of a K tile)
  uint64_t stamp;
};

struct PlanCache {
  std::mutex mu;
  std::vector<Plan> plans;
  uint64_t clock = 0;
  static constexpr size_t kCap = 256;

  static bool same(const Plan& a, const void* q, const void* k, const void* v, const void* o,
                   const Problem& p, int device) {
    return a.q == q && a.k == k && a.v == v && a.o == o && a.device == device &&
           a.p.B == p.B && a.p.H == p.H && a.p.Nq == p.Nq && a.p.Nkv == p.Nkv && a.p.D == p.D &&
           a.p.dtype == p.dtype && memcmp(a.p.qs, p.qs, sizeof p.qs) == 0 &&
           memcmp(a.p.ks, p.ks, sizeof p.ks) == 0 && memcmp(a.p.vs, p.vs, sizeof p.vs) == 0 &&
           memcmp(a.p.os, p.os, sizeof p.os) == 0;
  }

  int get(const void* q, const void* k, const void* v, void* o, const Problem& p, int device,
          Plan* out) {
    std::lock_guard<std::mutex> lk(mu);
    ++clock;
    for (auto& pl : plans) {
      if (same(pl, q, k, v, o, p, device)) {
        pl.stamp = clock;
        *out = pl;
        return FA_OK;
      }
    }
    Plan pl;
    pl.q = q; pl.k = k; pl.v = v; pl.o = o; pl.p = p; pl.device = device; pl.stamp = clock;
    int rc;
    if ((rc = make_map(&pl.mq, q, p.B, p.H, p.Nq, p.D, p.qs, p.dtype, fa::kTileM))) return rc;
    if ((rc = make_map(&pl.mk, k, p.B, p.H, p.Nkv, p.D, p.ks, p.dtype, fa::kTileN))) return rc;
    if ((rc = make_map(&pl.mv, v, p.B, p.H, p.Nkv, p.D, p.vs, p.dtype, fa::kTileN))) return rc;
    if ((rc = make_map(&pl.mo, o, p.B, p.H, p.Nq, p.D, p.os, p.dtype, fa::kTileM))) return rc;
    if ((rc = make_map(&pl.mk64, k, p.B, p.H, p.Nkv, p.D, p.ks, p.dtype, fa::kTileN / 2))) return rc;
    if (plans.size() < kCap) {
      plans.push_back(pl);
    } else {
      size_t victim = 0;
      for (size_t i = 1; i < plans.size(); ++i)
        if (plans[i].stamp < plans[victim].stamp) victim = i;
      plans[victim] = pl;
    }
    *out = pl;
    return FA_OK;
  }
};
PlanCache g_plans;

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
// The dynamic shared-memory opt-in is per (kernel, device); `done` remembers the devices served.
template <typename K>
int set_smem(K kernel, int bytes, std::atomic<uint64_t>* done = nullptr, int device = 0) {
  const uint64_t bit = uint64_t(1) << (device & 63);
  if (done != nullptr && (done->load(std::memory_order_acquire) & bit)) return FA_OK;
  FA_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (done != nullptr) done->fetch_or(bit, std::memory_order_release);
  return FA_OK;
}

// One launch of a forward kernel.  With PDL the launch carries the programmatic-stream-serialization
// attribute: the kernel may become resident while its predecessor on the stream is still running and does
// its prologue (barrier init, TMEM allocation, descriptor / L2 prefetch) there; every kernel calls
// griddepcontrol.wait before it touches global memory, so stream order is preserved for the data.
// Cluster kernels carry their cluster shape at compile time (__cluster_dims__).
template <typename... KArgs, typename... Args>
int launch_fwd(void (*kernel)(KArgs...), dim3 grid, int threads, int smem, cudaStream_t stream,
               Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(static_cast<unsigned>(threads));
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl.load(std::memory_order_relaxed) ? 1 : 0;
  FA_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return FA_OK;
}

// Launch grid over (row blocks per head, H, B): 3-D as is for non-causal problems, flattened to one dimension
// for causal ones, where the kernels order the blocks longest first across heads (fa::work_coords).
dim3 block_grid(long long blocks_per_head, const Problem& p) {
  if (p.causal) return dim3(static_cast<unsigned>(blocks_per_head * p.H * p.B), 1, 1);
  return dim3(static_cast<unsigned>(blocks_per_head), static_cast<unsigned>(p.H), static_cast<unsigned>(p.B));
}

template <int kDP, bool kBF16, bool kCausal>
int launch_ws(const Plan& pl, float* lse, cudaStream_t stream) {
  const Problem& p = pl.p;
  auto kernel = fa::fa_fwd_ws_kernel<kDP, kBF16, kCausal>;
  constexpr int smem = fa::WsCfg<kDP>::kTotal;
  static std::atomic<uint64_t> configured{0};
  int rc = set_smem(kernel, smem, &configured, pl.device);
  if (rc) return rc;
  fa::TcParams tp{lse, p.Nq, p.Nkv, p.H, p.scale * 1.4426950408889634f, nullptr, nullptr, 0, 0, 0, 0 FA_TP_TRACE};
  const dim3 grid = block_grid((p.Nq + 2 * fa::kTileM - 1) / (2 * fa::kTileM), p);
  return launch_fwd(kernel, grid, fa::kWsThreads, smem, stream, pl.mq, pl.mk, pl.mv, pl.mo, tp);
}

// ---------------------------------------------------------------------------------------------
// persistent stream-K forward: per-(device, stream) workspace for the partial results of split units
// ---------------------------------------------------------------------------------------------
// One entry per (device, stream), always sized for the largest configuration (SM count + 1 slots of the
// head-dim-128 slot size, ~20 MB), so an entry is never re-allocated and its address - which CUDA graphs
// captured on that stream have baked in - stays valid until fa_host_workspace_release().  Launches on one
// stream are ordered, so one slot set per stream is enough; see include/fa_fwd_sm100.h for the rule this
// implies for graphs replayed on other streams.
struct SkWorkspace {
  int device = -1;
  cudaStream_t stream = nullptr;
  float* ws = nullptr;
  int* flags = nullptr;
  int slots = 0;
};
std::mutex g_sk_mu;
std::vector<SkWorkspace> g_sk_ws;
constexpr size_t kSkMaxWorkspaces = 64;

int sm_count(int device) {
  static std::mutex mu;
  static int cached[64];
  std::lock_guard<std::mutex> lk(mu);
  if (device < 64 && cached[device] > 0) return cached[device];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  if (device < 64) cached[device] = n;
  return n;
}

// Workspace for launches on `stream`, or nullptr if none can be provided right now (stream capture in
// progress with nothing cached, allocation failure, more than kSkMaxWorkspaces streams): the caller then
// uses the one-shot kernel.
SkWorkspace* get_sk_workspace(int device, cudaStream_t stream, int slots) {
  std::lock_guard<std::mutex> lk(g_sk_mu);
  for (auto& w : g_sk_ws)
    if (w.device == device && w.stream == stream && w.slots >= slots) return &w;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) {
    (void)cudaGetLastError();
    return nullptr;
  }
  if (g_sk_ws.size() >= kSkMaxWorkspaces) return nullptr;
  SkWorkspace w;
  w.device = device;
  w.stream = stream;
  w.slots = slots;
  const size_t ws_bytes = static_cast<size_t>(slots) * fa::kSkSlotFloatsMax * sizeof(float);
  const size_t flag_bytes = static_cast<size_t>(slots) * 2 * sizeof(int);
  // The flags are cleared ON THE LAUNCH STREAM (verified above not to be capturing): the first kernel that
  // reads them is ordered behind the memset whatever kind of stream this is.
  if (cudaMalloc(reinterpret_cast<void**>(&w.ws), ws_bytes) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&w.flags), flag_bytes) != cudaSuccess ||
      cudaMemsetAsync(w.flags, 0, flag_bytes, stream) != cudaSuccess) {
    (void)cudaGetLastError();
    if (w.ws) cudaFree(w.ws);
    if (w.flags) cudaFree(w.flags);
    return nullptr;
  }
  g_sk_ws.reserve(kSkMaxWorkspaces);  // pointers handed out stay valid
  g_sk_ws.push_back(w);
  return &g_sk_ws.back();
}

void release_sk_workspaces(int device) {
  std::lock_guard<std::mutex> lk(g_sk_mu);
  for (size_t i = 0; i < g_sk_ws.size();) {
    if (g_sk_ws[i].device == device) {
      cudaFree(g_sk_ws[i].ws);
      cudaFree(g_sk_ws[i].flags);
      g_sk_ws[i] = g_sk_ws.back();
      g_sk_ws.pop_back();
    } else {
      ++i;
    }
  }
}

// What the persistent kernel can run at all: non-causal, whole 256-row query blocks.
bool sk_possible(const Problem& p) { return !p.causal && p.Nq % (2 * fa::kTileM) == 0; }

// Cost model for the persistent kernel against the one-shot two-tile kernels, in units of one
// "step" = the time the two-tile kernel (ws) needs for one KV tile of a 256-row query block:
//   one-shot (ws)    ceil(U / n_sm) rounds of T steps, ~2 steps of fixed cost per CTA (launch, prologue, first
//                    loads, first softmax, epilogue, drain)
//   early-S (ws3)    head dims <= 64 only, on CTA pairs: ceil(pairs / (n_sm / 2)) rounds of 0.83 T steps (866 vs 721
//                    TFLOPS at N=16384) + ~4.5 steps per round (cluster start-up; B=2 H=10 N=4096 D=64: 147 vs 155 us)
//   persistent (sk)  ceil(U T / n_sm) steps, ~1 step per unit boundary inside a CTA (pipelined: round-2 A/B on
//                    whole units, 592 units on 148 SMs: 232 vs 236 us; 0.75 for loops of fewer than 4 KV tiles), a
//                    penalty when units are split between CTAs (the partial's round trip through L2 and the merge
//                    at the end of the range: ~8 steps at head dim 128, ~4 at head dims <= 64 where the partial is
//                    half the size; none when T = 1, where items are whole units), ~2 steps of fixed cost once
// Fitted to profiles/r02_sweep_kernels.json (fp16 H=16 D=128, sk vs ws in us): N=2048 38.4 / 31.2, 4096 110.7 /
// 112.1, 8192 397 / 431, 16384 1538 / 1545, and to the Stable-Diffusion shapes of tools/bench_sd_shapes.py
// (profiles/r02_bench_sd_shapes_kernels.txt, sk / ws3 / ws in us): B=2 H=10 N=4096 D=64 123 / 147 / 155,
// B=2 H=20 N=1024 D=64 24.9 / 31.4 / 31.6, B=2 H=8 N=4096 D=40 103 / 101 / 108, cross-attention (77 keys)
// B=2 H=10 Nq=4096 D=64 13.2 / 19.3 / 16.2, D=128 14.8 / - / 20.0.  With the early S issue in the persistent kernel at
// head dims <= 64 (third session): B=2 H=10 N=4096 D=64 113 / 148, B=1 H=10 N=4096 D=64 61 / 99, B=2 H=20 N=1024 D=64
// 23.5 / 31.7, B=2 H=8 N=4096 D=40 94.9 / 100.0, B=2 H=8 N=16384 D=40 1330 / 1283.
struct KernelCosts {
  double ws, sk, ws3;  // ws3 = +inf where that kernel does not apply
};
KernelCosts estimate_costs(const Problem& p, int n_sm) {
  const double T = static_cast<double>((p.Nkv + fa::kTileN - 1) / fa::kTileN);
  const long long U = static_cast<long long>(p.B) * p.H * ((p.Nq + 2 * fa::kTileM - 1) / (2 * fa::kTileM));
  KernelCosts k;
  k.ws = static_cast<double>((U + n_sm - 1) / n_sm) * (T + 2.0);
  const double per_cta = std::ceil(static_cast<double>(U) * T / n_sm);
  const bool split = (U % n_sm) != 0 && T > 1.0;
  const double boundary = (T < 4.0) ? 0.75 : 1.0;
  const double merge = (p.D <= 64) ? 4.0 : 8.0;
  // head dims <= 64: the persistent kernel issues S_t(g+1) ahead of PV_t(g) too (P in spare tensor memory), which makes
  // its step 0.89 of the two-tile kernel's (the early-S kernel on CTA pairs: 0.83)
  const double step = (p.D <= 64) ? 0.89 : 1.0;
  k.sk = step * per_cta + boundary * std::ceil(per_cta / T) + (split ? merge : 0.0) + 2.0;
  k.ws3 = std::numeric_limits<double>::infinity();
  if (p.D <= 64 && T >= 4.0 && n_sm >= 2) {
    const long long pairs = (U + 1) / 2, slots = n_sm / 2;
    k.ws3 = static_cast<double>((pairs + slots - 1) / slots) * (0.83 * T + 4.5);
  }
  return k;
}

// shape-only eligibility (the SM count defaults to a B200's 148 when no device is consulted)
bool sk_eligible(const Problem& p, int n_sm) {
  if (!sk_possible(p) || n_sm <= 0) return false;
  // every 128-row tile gets an SM of its own in one round: the one-tile kernel's territory (choose_kernel_shape)
  if (static_cast<long long>(p.B) * p.H * (p.Nq / fa::kTileM) <= n_sm) return false;
  const KernelCosts k = estimate_costs(p, n_sm);
  // head dims <= 64: measured only with more units than SMs (fewer: every unit is shared by several CTAs and the chain
  // of partials grows; the early-S kernel or the one-tile kernel serve those)
  if (p.D <= 64 && static_cast<long long>(p.B) * p.H * (p.Nq / (2 * fa::kTileM)) <= n_sm) return false;
  if (p.Nkv < 4 * fa::kTileN) {
    // Loops of a tile or two (cross-attention): nothing to split, but with more units than SMs a persistent CTA's
    // unit boundary is cheaper than a CTA turnover (B=2 H=10 Nq=4096 Nkv=77: 13.2 vs 16.2 us at D=64, 14.8 vs 20.0 us
    // at D=128; B=2 H=8 Nq=16384 D=40: 26.9 vs 48.6 us).  Only worth it beyond one round.
    const long long U = static_cast<long long>(p.B) * p.H * (p.Nq / (2 * fa::kTileM));
    if (U <= n_sm) return false;
  }
  return k.sk < 0.985 * std::min(k.ws, k.ws3);
}

template <int kDP, bool kBF16>
int launch_sk(const Plan& pl, float* lse, cudaStream_t stream, bool* launched) {
  const Problem& p = pl.p;
  *launched = false;
  const int n_sm = sm_count(pl.device);
  if (!sk_possible(p) || n_sm <= 0) return FA_OK;
  SkWorkspace* w = get_sk_workspace(pl.device, stream, n_sm + 1);
  if (w == nullptr) return FA_OK;
  auto kernel = fa::fa_fwd_sk_kernel<kDP, kBF16>;
  constexpr int smem = fa::SkCfg<kDP>::kTotal;
  static std::atomic<uint64_t> configured{0};
  int rc = set_smem(kernel, smem, &configured, pl.device);
  if (rc) return rc;
  const int T = (p.Nkv + fa::kTileN - 1) / fa::kTileN;
  const int P = p.Nq / (2 * fa::kTileM);
  const long long U = static_cast<long long>(p.B) * p.H * P;
  // one CTA per SM, fewer when there are fewer (unit, KV tile) items than SMs
  const int G = static_cast<int>(std::min<long long>(n_sm, U * T));
  // all but the last full round are data-parallel; the last G + U % G units are split evenly
  const int dp = (U % G == 0) ? static_cast<int>(U / G) : static_cast<int>(std::max<long long>(U / G - 1, 0));
  const long long W = (U - static_cast<long long>(dp) * G) * T;
  fa::TcParams tp{lse, p.Nq, p.Nkv, p.H, p.scale * 1.4426950408889634f, w->ws, w->flags, W, dp, T, P FA_TP_TRACE};
  rc = launch_fwd(kernel, dim3(G), fa::kWsThreads, smem, stream, pl.mq, pl.mk, pl.mv, pl.mo, tp);
  *launched = (rc == FA_OK);
  return rc;
}

template <int kDP, bool kBF16, bool kCausal>
int launch_tc1(const Plan& pl, float* lse, cudaStream_t stream) {
  const Problem& p = pl.p;
  auto kernel = fa::fa_fwd_tc1_kernel<kDP, kBF16, kCausal>;
  constexpr int smem = fa::Tc1Smem<kDP>::kTotal;
  static std::atomic<uint64_t> configured{0};
  int rc = set_smem(kernel, smem, &configured, pl.device);
  if (rc) return rc;
  fa::TcParams tp{lse, p.Nq, p.Nkv, p.H, p.scale * 1.4426950408889634f, nullptr, nullptr, 0, 0, 0, 0 FA_TP_TRACE};
  const dim3 grid = block_grid((p.Nq + fa::kTileM - 1) / fa::kTileM, p);
  return launch_fwd(kernel, grid, 128, smem, stream, pl.mq, pl.mk, pl.mv, pl.mo, tp);
}

template <int kDP, bool kBF16, bool kCausal>
int launch_wide(const Plan& pl, float* lse, cudaStream_t stream) {
  const Problem& p = pl.p;
  auto kernel = fa::fa_fwd_wide_kernel<kDP, kBF16, kCausal>;
  constexpr int smem = fa::WideCfg<kDP>::kTotal;
  static std::atomic<uint64_t> configured{0};
  int rc = set_smem(kernel, smem, &configured, pl.device);
  if (rc) return rc;
  fa::TcParams tp{lse, p.Nq, p.Nkv, p.H, p.scale * 1.4426950408889634f, nullptr, nullptr, 0, 0, 0, 0 FA_TP_TRACE};
  const dim3 grid = block_grid((p.Nq + fa::kTileM - 1) / fa::kTileM, p);
  return launch_fwd(kernel, grid, fa::kWideThreads, smem, stream, pl.mq, pl.mk, pl.mv, pl.mo, tp);
}

// two-tile kernel on CTA pairs (cluster of 2, cta_group::2): non-causal, head dims <= 128
template <int kDP, bool kBF16>
int launch_ws2(const Plan& pl, float* lse, cudaStream_t stream) {
  const Problem& p = pl.p;
  auto kernel = fa::fa_fwd_ws2_kernel<kDP, kBF16>;
  constexpr int smem = fa::Ws2Cfg<kDP>::kTotal;
  static std::atomic<uint64_t> configured{0};
  int rc = set_smem(kernel, smem, &configured, pl.device);
  if (rc) return rc;
  fa::TcParams tp{lse, p.Nq, p.Nkv, p.H, p.scale * 1.4426950408889634f, nullptr, nullptr, 0, 0, 0, 0 FA_TP_TRACE};
  const int blocks = (p.Nq + 2 * fa::kTileM - 1) / (2 * fa::kTileM);
  dim3 grid((blocks + 1) & ~1, p.H, p.B);  // whole pairs
  return launch_fwd(kernel, grid, fa::kWsThreads, smem, stream, pl.mq, pl.mk64, pl.mv, pl.mo, tp);
}

// two-tile kernel on CTA pairs with P through shared memory (S_t(j+1) issued ahead of PV_t(j)): non-causal
template <int kDP, bool kBF16, bool kCausal>
int launch_ws3(const Plan& pl, float* lse, cudaStream_t stream) {
  const Problem& p = pl.p;
  auto kernel = fa::fa_fwd_ws3_kernel<kDP, kBF16, kCausal>;
  constexpr int smem = fa::Ws3Cfg<kDP>::kTotal;
  static std::atomic<uint64_t> configured{0};
  int rc = set_smem(kernel, smem, &configured, pl.device);
  if (rc) return rc;
  fa::TcParams tp{lse, p.Nq, p.Nkv, p.H, p.scale * 1.4426950408889634f, nullptr, nullptr, 0, 0, 0, 0 FA_TP_TRACE};
  const int blocks = (p.Nq + 2 * fa::kTileM - 1) / (2 * fa::kTileM);
  const dim3 grid = block_grid((blocks + 1) & ~1, p);
  return launch_fwd(kernel, grid, fa::kWsThreads, smem, stream, pl.mq, pl.mk64, pl.mv, pl.mo, tp);
}

// CTA-pair kernel (cluster of 2, cta_group::2): padded head dim 192 or 256
template <int kDP, bool kBF16, bool kCausal>
int launch_wide2(const Plan& pl, float* lse, cudaStream_t stream) {
  const Problem& p = pl.p;
  auto kernel = fa::fa_fwd_wide2_kernel<kDP, kBF16, kCausal>;
  constexpr int smem = fa::Wide2Cfg<kDP>::kTotal;
  static std::atomic<uint64_t> configured{0};
  int rc = set_smem(kernel, smem, &configured, pl.device);
  if (rc) return rc;
  fa::TcParams tp{lse, p.Nq, p.Nkv, p.H, p.scale * 1.4426950408889634f, nullptr, nullptr, 0, 0, 0, 0 FA_TP_TRACE};
  const int tiles = (p.Nq + fa::kTileM - 1) / fa::kTileM;
  const dim3 grid = block_grid((tiles + 1) & ~1, p);  // whole pairs: an odd last tile gets a partner that is all padding
  return launch_fwd(kernel, grid, fa::kWideThreads, smem, stream, pl.mq, pl.mk64, pl.mv, pl.mo, tp);
}

template <int kDP, bool kBF16, bool kCausal>
int launch_tc_variant(int kernel, const Plan& pl, float* lse, cudaStream_t stream) {
  switch (kernel) {
    case FA_KERNEL_SK: {
      if constexpr (!kCausal) {
        bool launched = false;
        int rc = launch_sk<kDP, kBF16>(pl, lse, stream, &launched);
        if (rc || launched) return rc;
      }
      return launch_ws<kDP, kBF16, kCausal>(pl, lse, stream);  // not eligible / no workspace
    }
    case FA_KERNEL_WS: return launch_ws<kDP, kBF16, kCausal>(pl, lse, stream);
    case FA_KERNEL_WS3:
      // P outside the S columns needs spare tensor memory: head dims <= 64 (the shared-memory route at 128 was
      // measured 5 % slower than ws2 and is no longer instantiated)
      if constexpr (kDP == 64) return launch_ws3<kDP, kBF16, kCausal>(pl, lse, stream);
      return launch_ws<kDP, kBF16, kCausal>(pl, lse, stream);
    case FA_KERNEL_WS2:
      if constexpr (!kCausal) return launch_ws2<kDP, kBF16>(pl, lse, stream);
      return launch_ws<kDP, kBF16, kCausal>(pl, lse, stream);  // the pair kernel is non-causal only
    case FA_KERNEL_TC1: return launch_tc1<kDP, kBF16, kCausal>(pl, lse, stream);
  }
  return fail(FA_ERR_INVALID_ARG, "unknown tensor-core kernel selector");
}

int launch_tc(int kernel, const Plan& pl, float* lse, cudaStream_t stream) {
  const Problem& p = pl.p;
  const bool bf = p.dtype == FA_DTYPE_BF16;
  const bool ca = p.causal != 0;
#define FA_DISPATCH(DP)                                                              \
  do {                                                                               \
    if (bf) {                                                                        \
      if (ca) return launch_tc_variant<DP, true, true>(kernel, pl, lse, stream);     \
      return launch_tc_variant<DP, true, false>(kernel, pl, lse, stream);            \
    }                                                                                \
    if (ca) return launch_tc_variant<DP, false, true>(kernel, pl, lse, stream);      \
    return launch_tc_variant<DP, false, false>(kernel, pl, lse, stream);             \
  } while (0)
  if (p.D > 128) {
    if (kernel != FA_KERNEL_WIDE)
      return fail(FA_ERR_UNSUPPORTED, "head dims 129..256 are served by FA_KERNEL_WIDE (or FA_KERNEL_SIMT) only");
#define FA_DISPATCH_WIDE(DP)                                                  \
  do {                                                                        \
    if (bf) {                                                                 \
      if (ca) return launch_wide<DP, true, true>(pl, lse, stream);            \
      return launch_wide<DP, true, false>(pl, lse, stream);                   \
    }                                                                         \
    if (ca) return launch_wide<DP, false, true>(pl, lse, stream);             \
    return launch_wide<DP, false, false>(pl, lse, stream);                    \
  } while (0)
    // CTA pairs pay where the one-CTA kernel is bound by shared-memory traffic: padded head dim 256
    // (+14 %).  At 192 the pair kernel (Wide2Cfg<192> is supported and was measured) ties with the one-CTA
    // kernel - 1563 vs 1577 TFLOPS, one softmax group is the limit there - so it is not instantiated.
    if (p.D > 192 && g_wide_pairs.load(std::memory_order_relaxed)) {
      if (bf) return ca ? launch_wide2<256, true, true>(pl, lse, stream) : launch_wide2<256, true, false>(pl, lse, stream);
      return ca ? launch_wide2<256, false, true>(pl, lse, stream) : launch_wide2<256, false, false>(pl, lse, stream);
    }
    if (p.D <= 192) FA_DISPATCH_WIDE(192);
    FA_DISPATCH_WIDE(256);
  }
  if (kernel == FA_KERNEL_WIDE) {  // the one-tile arrangement at head dims <= 128
    // (The pair kernel was also instantiated and measured here - Wide2Cfg<64/128> - and loses to both this
    // kernel and the two-tile kernel: 669 / 1284 TFLOPS at D = 64 / 128, N=16384)
    if (p.D <= 64) FA_DISPATCH_WIDE(64);
    FA_DISPATCH_WIDE(128);
  }
  if (p.D <= 64) FA_DISPATCH(64);
  FA_DISPATCH(128);
#undef FA_DISPATCH
#undef FA_DISPATCH_WIDE
}

int launch_simt(const void* q, const void* k, const void* v, void* o, float* lse, const Problem& p,
                cudaStream_t stream) {
  fa::SimtParams sp;
  sp.q = q; sp.k = k; sp.v = v; sp.o = o; sp.lse = lse;
  sp.B = p.B; sp.H = p.H; sp.Nq = p.Nq; sp.Nkv = p.Nkv; sp.D = p.D;
  memcpy(sp.qs, p.qs, sizeof sp.qs);
  memcpy(sp.ks, p.ks, sizeof sp.ks);
  memcpy(sp.vs, p.vs, sizeof sp.vs);
  memcpy(sp.os, p.os, sizeof sp.os);
  sp.causal = p.causal;
  sp.scale_log2 = p.scale * 1.4426950408889634f;
  dim3 grid((p.Nq + fa::kSimtWarps - 1) / fa::kSimtWarps, p.H, p.B);
  const int smem = fa::kSimtWarps * p.D * static_cast<int>(sizeof(float));
  if (p.dtype == FA_DTYPE_BF16)
    fa::fa_fwd_simt_kernel<__nv_bfloat16><<<grid, fa::kSimtWarps * 32, smem, stream>>>(sp);
  else
    fa::fa_fwd_simt_kernel<__half><<<grid, fa::kSimtWarps * 32, smem, stream>>>(sp);
  FA_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return FA_OK;
}

// ---------------------------------------------------------------------------------------------
// backward launchers
// ---------------------------------------------------------------------------------------------
struct BwdMaps {
  CUtensorMap q, k, v, d_o, dk, dv, dq;
  CUtensorMap dq32;  // the same accumulator with a {32 columns, 32 rows} box (one per drain warp, fa_bwd_ws.cuh)
  // head dims 129..256 (fa_bwd_wide.cuh): the inputs again with 64-row boxes (streamed tiles) and dQ as a 16-bit output
  CUtensorMap q64, k64, v64, do64, dq16;
};

// 3-D fp32 map over the contiguous dq accumulator [B*H, Nq, DP]: box {32 columns = one 128-byte
// swizzle row, 128 rows, 1}; rows >= Nq of a partial tile are clipped by the reduce.
int make_map_dq(CUtensorMap* map, float* base, int BH, int Nq, int DP, int box_rows = 128) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return fail(FA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(DP), static_cast<cuuint64_t>(Nq),
                        static_cast<cuuint64_t>(BH)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(DP) * 4, static_cast<cuuint64_t>(Nq) * DP * 4};
  cuuint32_t box[3] = {32, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(dq accumulator) failed (CUresult %d) dims=[%d,%d,%d]",
             static_cast<int>(r), DP, Nq, BH);
    return fail(FA_ERR_CUDA, buf);
  }
  return FA_OK;
}

template <int kDP, bool kBF16, bool kCausal>
int launch_bwd_tc(const BwdMaps& m, const fa::BwdParams& bp, int B, int H, int Nkv, int device,
                  cudaStream_t stream) {
  dim3 grid((Nkv + fa::kTileN - 1) / fa::kTileN, H, B);
  auto kernel = fa::fa_bwd_tc_kernel<kDP, kBF16, kCausal>;
  constexpr int smem = fa::BwdSmem<kDP>::kTotal;
  static std::atomic<uint64_t> configured{0};
  int rc = set_smem(kernel, smem, &configured, device);
  if (rc) return rc;
  kernel<<<grid, 256, smem, stream>>>(m.q, m.k, m.v, m.d_o, m.dk, m.dv, m.dq, bp);
  FA_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return FA_OK;
}

template <int kDP, bool kBF16, bool kCausal>
int launch_bwd_ws(const BwdMaps& m, const fa::BwdParams& bp, int B, int H, int Nkv, int device,
                  cudaStream_t stream) {
  const unsigned n_j = static_cast<unsigned>((Nkv + fa::kTileN - 1) / fa::kTileN);
  // causal: 1-D grid, longest key tile first across groups of (batch, head) pairs of about four rounds of CTAs (see the kernel)
  const unsigned ctas = n_j * static_cast<unsigned>(H) * static_cast<unsigned>(B);
  // ... up to ten rounds of CTAs: beyond that (N=16384, H=16: 13.8 rounds) the natural order - a head's key tiles next to each
  // other, walking the same query tiles almost in step - measures better (1075 vs 1045 TFLOPS) than any longest-first order
  const dim3 grid = (kCausal && ctas <= 10u * 148u) ? dim3(ctas, 1, 1)
                                                    : dim3(n_j, static_cast<unsigned>(H), static_cast<unsigned>(B));
  fa::BwdParams bpl = bp;
  bpl.lpt_group = std::max(1, std::min(H * B, static_cast<int>((4u * 148u + n_j / 2) / n_j)));
  auto kernel = fa::fa_bwd_ws_kernel<kDP, kBF16, kCausal>;
  constexpr int smem = fa::BwdWsSmem<kDP>::kTotal;
  static std::atomic<uint64_t> configured{0};
  int rc = set_smem(kernel, smem, &configured, device);
  if (rc) return rc;
  kernel<<<grid, fa::kBwdWsThreads, smem, stream>>>(m.q, m.k, m.v, m.d_o, m.dk, m.dv, m.dq32, bpl);
  FA_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return FA_OK;
}

// 0 = automatic (the pipelined kernel fa_bwd_ws.cuh), 1 = the serial kernel fa_bwd_tc.cuh, 2 = fa_bwd_ws.cuh
std::atomic<int> g_bwd_kernel{0};

int dispatch_bwd_tc(const BwdMaps& m, const fa::BwdParams& bp, int B, int H, int Nkv, int D, int dtype,
                    int causal, int device, cudaStream_t stream) {
  const bool bf = dtype == FA_DTYPE_BF16;
  const bool ca = causal != 0;
  if (g_bwd_kernel.load() != 1) {
#define FA_BWD_WS_DISPATCH(DP)                                                                 \
  do {                                                                                         \
    if (bf) {                                                                                  \
      if (ca) return launch_bwd_ws<DP, true, true>(m, bp, B, H, Nkv, device, stream);          \
      return launch_bwd_ws<DP, true, false>(m, bp, B, H, Nkv, device, stream);                 \
    }                                                                                          \
    if (ca) return launch_bwd_ws<DP, false, true>(m, bp, B, H, Nkv, device, stream);           \
    return launch_bwd_ws<DP, false, false>(m, bp, B, H, Nkv, device, stream);                  \
  } while (0)
    if (D <= 64) FA_BWD_WS_DISPATCH(64);
    FA_BWD_WS_DISPATCH(128);
#undef FA_BWD_WS_DISPATCH
  }
#define FA_BWD_DISPATCH(DP)                                                                    \
  do {                                                                                         \
    if (bf) {                                                                                  \
      if (ca) return launch_bwd_tc<DP, true, true>(m, bp, B, H, Nkv, device, stream);          \
      return launch_bwd_tc<DP, true, false>(m, bp, B, H, Nkv, device, stream);                 \
    }                                                                                          \
    if (ca) return launch_bwd_tc<DP, false, true>(m, bp, B, H, Nkv, device, stream);           \
    return launch_bwd_tc<DP, false, false>(m, bp, B, H, Nkv, device, stream);                  \
  } while (0)
  if (D <= 64) FA_BWD_DISPATCH(64);
  FA_BWD_DISPATCH(128);
#undef FA_BWD_DISPATCH
}

template <int kDP, bool kBF16, bool kCausal, int kMode>
int launch_bwd_wide_mode(const BwdMaps& m, const fa::BwdWideParams& wp, int B, int H, int n_rows, int device,
                         cudaStream_t stream) {
  dim3 grid(static_cast<unsigned>((n_rows + fa::kTileM - 1) / fa::kTileM), static_cast<unsigned>(H),
            static_cast<unsigned>(B));
  auto kernel = fa::fa_bwd_wide_kernel<kDP, kBF16, kCausal, kMode>;
  constexpr int smem = fa::BwdWideSmem<kDP, kMode>::kTotal;
  static std::atomic<uint64_t> configured{0};
  int rc = set_smem(kernel, smem, &configured, device);
  if (rc) return rc;
  if (kMode == fa::kBwdWideDV)
    kernel<<<grid, fa::kBwdWideThreads, smem, stream>>>(m.k, m.k, m.q64, m.do64, m.dv, m.dv, wp);
  else if (kMode == fa::kBwdWideDK)
    kernel<<<grid, fa::kBwdWideThreads, smem, stream>>>(m.k, m.v, m.q64, m.do64, m.dk, m.dk, wp);
  else if (kMode == fa::kBwdWideDKV)
    kernel<<<grid, fa::kBwdWideThreads, smem, stream>>>(m.k, m.v, m.q64, m.do64, m.dv, m.dk, wp);
  else
    kernel<<<grid, fa::kBwdWideThreads, smem, stream>>>(m.q, m.d_o, m.k64, m.v64, m.dq16, m.dq16, wp);
  FA_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return FA_OK;
}

// head dims 129..256: dV, dK and dQ by three launches of fa_bwd_wide_kernel (see fa_bwd_wide.cuh)
template <int kDP, bool kBF16, bool kCausal>
int launch_bwd_wide(const BwdMaps& m, fa::BwdWideParams wp, int B, int H, int Nq, int Nkv, float scale, int device,
                    cudaStream_t stream) {
  int rc;
  if constexpr (kDP <= 192) {  // dV and dK in one launch: both accumulators fit in tensor memory next to the scores
    wp.out_scale = scale;
    if ((rc = launch_bwd_wide_mode<kDP, kBF16, kCausal, fa::kBwdWideDKV>(m, wp, B, H, Nkv, device, stream))) return rc;
  } else {
    wp.out_scale = 1.f;
    if ((rc = launch_bwd_wide_mode<kDP, kBF16, kCausal, fa::kBwdWideDV>(m, wp, B, H, Nkv, device, stream))) return rc;
    wp.out_scale = scale;
    if ((rc = launch_bwd_wide_mode<kDP, kBF16, kCausal, fa::kBwdWideDK>(m, wp, B, H, Nkv, device, stream))) return rc;
  }
  wp.out_scale = scale;
  return launch_bwd_wide_mode<kDP, kBF16, kCausal, fa::kBwdWideDQ>(m, wp, B, H, Nq, device, stream);
}

int dispatch_bwd_wide(const BwdMaps& m, const fa::BwdWideParams& wp, int B, int H, int Nq, int Nkv, int D, int dtype,
                      int causal, float scale, int device, cudaStream_t stream) {
  const bool bf = dtype == FA_DTYPE_BF16;
  const bool ca = causal != 0;
#define FA_BWD_WIDE_DISPATCH(DP)                                                                         \
  do {                                                                                                   \
    if (bf) {                                                                                            \
      if (ca) return launch_bwd_wide<DP, true, true>(m, wp, B, H, Nq, Nkv, scale, device, stream);       \
      return launch_bwd_wide<DP, true, false>(m, wp, B, H, Nq, Nkv, scale, device, stream);              \
    }                                                                                                    \
    if (ca) return launch_bwd_wide<DP, false, true>(m, wp, B, H, Nq, Nkv, scale, device, stream);        \
    return launch_bwd_wide<DP, false, false>(m, wp, B, H, Nq, Nkv, scale, device, stream);               \
  } while (0)
  if (D <= 192) FA_BWD_WIDE_DISPATCH(192);
  FA_BWD_WIDE_DISPATCH(256);
#undef FA_BWD_WIDE_DISPATCH
}

// Tensor maps of one backward call, cached like the forward's Plan (encoding seven maps costs ~10 us of
// host time per call; training loops present the same buffers again and again through the caching allocator).
struct BwdPlan {
  const void* ptr[8];  // q, k, v, dO, dK, dV, dq accumulator (head dims <= 128), dQ (head dims > 128)
  int B, H, Nq, Nkv, D, dtype, device;
  int64_t st[7][4];
  BwdMaps m;
  uint64_t stamp;
};
struct BwdPlanCache {
  std::mutex mu;
  std::vector<BwdPlan> plans;
  uint64_t clock = 0;
  static constexpr size_t kCap = 64;
  int get(const BwdPlan& key, BwdMaps* out) {
    std::lock_guard<std::mutex> lk(mu);
    ++clock;
    for (auto& pl : plans) {
      if (memcmp(pl.ptr, key.ptr, sizeof key.ptr) == 0 && pl.B == key.B && pl.H == key.H && pl.Nq == key.Nq &&
          pl.Nkv == key.Nkv && pl.D == key.D && pl.dtype == key.dtype && pl.device == key.device &&
          memcmp(pl.st, key.st, sizeof key.st) == 0) {
        pl.stamp = clock;
        *out = pl.m;
        return FA_OK;
      }
    }
    BwdPlan pl = key;
    pl.stamp = clock;
    int rc;
    const int B = key.B, H = key.H, Nq = key.Nq, Nkv = key.Nkv, D = key.D, dt = key.dtype;
    if ((rc = make_map(&pl.m.q, key.ptr[0], B, H, Nq, D, key.st[0], dt, fa::kTileM))) return rc;
    if ((rc = make_map(&pl.m.k, key.ptr[1], B, H, Nkv, D, key.st[1], dt, fa::kTileN))) return rc;
    if ((rc = make_map(&pl.m.v, key.ptr[2], B, H, Nkv, D, key.st[2], dt, fa::kTileN))) return rc;
    if ((rc = make_map(&pl.m.d_o, key.ptr[3], B, H, Nq, D, key.st[3], dt, fa::kTileM))) return rc;
    if ((rc = make_map(&pl.m.dk, key.ptr[4], B, H, Nkv, D, key.st[4], dt, fa::kTileN))) return rc;
    if ((rc = make_map(&pl.m.dv, key.ptr[5], B, H, Nkv, D, key.st[5], dt, fa::kTileN))) return rc;
    if (D <= 128) {
      if ((rc = make_map_dq(&pl.m.dq, static_cast<float*>(const_cast<void*>(key.ptr[6])), B * H, Nq, D))) return rc;
      if ((rc = make_map_dq(&pl.m.dq32, static_cast<float*>(const_cast<void*>(key.ptr[6])), B * H, Nq, D, 32))) return rc;
    } else {
      if ((rc = make_map(&pl.m.q64, key.ptr[0], B, H, Nq, D, key.st[0], dt, fa::kWideT))) return rc;
      if ((rc = make_map(&pl.m.k64, key.ptr[1], B, H, Nkv, D, key.st[1], dt, fa::kWideT))) return rc;
      if ((rc = make_map(&pl.m.v64, key.ptr[2], B, H, Nkv, D, key.st[2], dt, fa::kWideT))) return rc;
      if ((rc = make_map(&pl.m.do64, key.ptr[3], B, H, Nq, D, key.st[3], dt, fa::kWideT))) return rc;
      if ((rc = make_map(&pl.m.dq16, key.ptr[7], B, H, Nq, D, key.st[6], dt, fa::kTileM))) return rc;
    }
    if (plans.size() < kCap) {
      plans.push_back(pl);
    } else {
      size_t victim = 0;
      for (size_t i = 1; i < plans.size(); ++i)
        if (plans[i].stamp < plans[victim].stamp) victim = i;
      plans[victim] = pl;
    }
    *out = pl.m;
    return FA_OK;
  }
};
BwdPlanCache g_bwd_plans;

int check_device(int* device_out) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return fail(FA_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
  }
  static std::mutex mu;
  static int cc_major[64];
  static bool known[64];
  {
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 64 && !known[dev]) {
      int major = 0;
      FA_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
      cc_major[dev] = major;
      known[dev] = true;
    }
  }
  if (dev < 64 && cc_major[dev] != 10)
    return fail(FA_ERR_NO_DEVICE, "the current device is not sm_100 (compute capability 10.x)");
  *device_out = dev;
  return FA_OK;
}

int resolve_kernel(const Problem& p, const void* q, const void* k, const void* v, const void* o,
                   int* kernel_out) {
  int kernel = choose_kernel_shape(p);
  if (kernel != FA_KERNEL_SIMT && q != nullptr &&
      !(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o)))
    kernel = FA_KERNEL_SIMT;
  const int forced = g_forced_kernel.load();
  if (forced != FA_KERNEL_AUTO) {
    if (forced != FA_KERNEL_SIMT && kernel == FA_KERNEL_SIMT)
      return fail(FA_ERR_UNSUPPORTED,
                  "forced tensor-core kernel cannot serve this problem (needs D % 8 == 0, D <= 256, "
                  "scale > 0, 16-byte aligned pointers and strides)");
    kernel = forced;
  }
  *kernel_out = kernel;
  return FA_OK;
}

int run_device(const void* q, const void* k, const void* v, void* o, float* lse, Problem& p,
               cudaStream_t stream) {
  int dev;
  int rc = check_device(&dev);
  if (rc) return rc;
  int kernel;
  if ((rc = resolve_kernel(p, q, k, v, o, &kernel))) return rc;
  if (kernel == FA_KERNEL_SIMT) return launch_simt(q, k, v, o, lse, p, stream);
  Plan pl;
  if ((rc = g_plans.get(q, k, v, o, p, dev, &pl))) return rc;
  pl.p.causal = p.causal;
  pl.p.scale = p.scale;
  return launch_tc(kernel, pl, lse, stream);
}

// ---------------------------------------------------------------------------------------------
// host-buffer path: chunk planner (pure host logic, exported as fa_host_plan_chunks for the CPU tests)
// ---------------------------------------------------------------------------------------------
// Heads per chunk of the pipelined H2D -> kernel -> D2H path over the flattened (b,h) axis.  The H2D stream
// is busy from the first byte to the last, so a call takes
//   T(n) = H2D_total + n * c_issue + (compute_total + D2H_total) / n
// (the tail of the last chunk is all that is not hidden; c_issue = the ~10 runtime calls a chunk costs on
// the host, ~40 us measured on B200 boxes: tools/e2e_probe.py).  n = sqrt(tail / c_issue) minimises it: 1 chunk
// at N=512, 8 at N=16384 for the sweep.  Then the last chunk is halved while that shortens the exposed tail
// (its kernel + its D2H) by more than the ~25 us of copy start-up and event hand-offs an extra chunk costs on
// the device; a CTA's run time is set by Nkv alone (~1.6 us per 128-key tile), so the kernel term stops
// shrinking once the chunk no longer fills the SMs.  FA_HOST_CHUNKS=n overrides the count (no tail split;
// "0" = planner count without the tail split).
std::vector<size_t> plan_host_chunks(size_t heads, int Nq, int Nkv, int D, int causal) {
  std::vector<size_t> chunk_heads;
  const size_t q_head = size_t(Nq) * D * 2;
  const double flops = 4.0 * double(heads) * Nq * Nkv * D * (causal ? 0.5 : 1.0);
  const double tail_us = flops / 1.0e9 + double(heads * q_head) / 50.0e3 + 12.0;
  double n = std::sqrt(tail_us / 40.0);
  const char* e = std::getenv("FA_HOST_CHUNKS");
  if (e != nullptr && std::atof(e) > 0) n = std::atof(e);
  size_t nc = n < 1.0 ? 1 : static_cast<size_t>(n + 0.5);
  if (nc > heads) nc = heads;
  const size_t hg = (heads + nc - 1) / nc;
  for (size_t h0 = 0; h0 < heads; h0 += hg) chunk_heads.push_back(h0 + hg <= heads ? hg : heads - h0);
  auto tail = [&](size_t nh) {
    const double cta_us = 10.0 + 1.6 * double((Nkv + 127) / 128) * (causal ? 0.5 : 1.0);
    const double waves = std::ceil(double(nh) * double((Nq + 255) / 256) / 148.0);
    return cta_us * waves + double(nh * q_head) / 50.0e3;
  };
  while (e == nullptr && chunk_heads.back() >= 2 &&
         tail(chunk_heads.back()) - tail(chunk_heads.back() / 2) > 25.0) {
    const size_t last = chunk_heads.back();
    chunk_heads.back() = last - last / 2;
    chunk_heads.push_back(last / 2);
  }
  return chunk_heads;
}

// ---------------------------------------------------------------------------------------------
// host-buffer path: per-device workspace + three streams
// ---------------------------------------------------------------------------------------------
// Two staging sets per device, used alternately: the H2D copies of call i+1 may run under the tail (last
// kernel + D2H) of call i, which is what fa_fwd_sm100_host_async() is for.
struct HostSet {
  void *dq = nullptr, *dk = nullptr, *dv = nullptr, *dout = nullptr;
  float* dlse = nullptr;
  size_t cap_q = 0, cap_kv = 0, cap_lse = 0;
  cudaEvent_t done = nullptr;  // recorded on s_out behind the last D2H of the call that used the set
  bool used = false;
};
struct HostWs {
  bool init = false;
  cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
  HostSet set[2];
  uint64_t calls = 0;
  std::vector<cudaEvent_t> ev_in, ev_run;
  std::mutex mu;  // one host call at a time per device (the workspace and its streams are per device);
                  // calls on different devices - the multi-GPU shard driven from one process - run concurrently
};
HostWs g_ws[64];

void host_drain(HostWs& w) {
  if (w.s_in) cudaStreamSynchronize(w.s_in);
  if (w.s_run) cudaStreamSynchronize(w.s_run);
  if (w.s_out) cudaStreamSynchronize(w.s_out);
  (void)cudaGetLastError();
}

int ws_release(HostWs& w) {
  host_drain(w);
  for (HostSet& st : w.set) {
    if (st.dq) cudaFree(st.dq);
    if (st.dk) cudaFree(st.dk);
    if (st.dv) cudaFree(st.dv);
    if (st.dout) cudaFree(st.dout);
    if (st.dlse) cudaFree(st.dlse);
    if (st.done) cudaEventDestroy(st.done);
    st = HostSet();
  }
  for (auto e : w.ev_in) cudaEventDestroy(e);
  for (auto e : w.ev_run) cudaEventDestroy(e);
  if (w.s_in) cudaStreamDestroy(w.s_in);
  if (w.s_run) cudaStreamDestroy(w.s_run);
  if (w.s_out) cudaStreamDestroy(w.s_out);
  w.init = false;  // (not `w = HostWs()`: the mutex the caller holds lives in w)
  w.s_in = w.s_run = w.s_out = nullptr;
  w.calls = 0;
  w.ev_in.clear();
  w.ev_run.clear();
  return FA_OK;
}

// Enqueue one host-buffer forward on the device's three streams (caller holds w.mu).  On return with
// FA_OK the work is in flight; `o` (and `lse`) are complete once s_out has drained.
int host_enqueue(HostWs& w, const void* q, const void* k, const void* v, void* o, float* lse, int B, int H,
                 int Nq, int Nkv, int D, int dtype, int causal, float scale) {
  int rc;
  if (!w.init) {
    FA_CUDA_TRY(cudaStreamCreateWithFlags(&w.s_in, cudaStreamNonBlocking));
    FA_CUDA_TRY(cudaStreamCreateWithFlags(&w.s_run, cudaStreamNonBlocking));
    FA_CUDA_TRY(cudaStreamCreateWithFlags(&w.s_out, cudaStreamNonBlocking));
    w.init = true;
  }
  const size_t heads = size_t(B) * H;
  const size_t q_head = size_t(Nq) * D * 2, kv_head = size_t(Nkv) * D * 2;
  const size_t bytes_q = heads * q_head, bytes_kv = heads * kv_head;
  const size_t bytes_lse = heads * Nq * sizeof(float);
  HostSet& st = w.set[w.calls & 1];
  if (st.cap_q < bytes_q || st.cap_kv < bytes_kv || (lse != nullptr && st.cap_lse < bytes_lse)) {
    host_drain(w);  // growing frees buffers that calls still in flight may be using
    if (st.cap_q < bytes_q) {
      if (st.dq) cudaFree(st.dq);
      if (st.dout) cudaFree(st.dout);
      st.dq = st.dout = nullptr; st.cap_q = 0;
      FA_CUDA_TRY(cudaMalloc(&st.dq, bytes_q));
      FA_CUDA_TRY(cudaMalloc(&st.dout, bytes_q));
      st.cap_q = bytes_q;
    }
    if (st.cap_kv < bytes_kv) {
      if (st.dk) cudaFree(st.dk);
      if (st.dv) cudaFree(st.dv);
      st.dk = st.dv = nullptr; st.cap_kv = 0;
      FA_CUDA_TRY(cudaMalloc(&st.dk, bytes_kv));
      FA_CUDA_TRY(cudaMalloc(&st.dv, bytes_kv));
      st.cap_kv = bytes_kv;
    }
    if (lse != nullptr && st.cap_lse < bytes_lse) {
      if (st.dlse) cudaFree(st.dlse);
      st.dlse = nullptr; st.cap_lse = 0;
      FA_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&st.dlse), bytes_lse));
      st.cap_lse = bytes_lse;
    }
  }
  if (st.done == nullptr) FA_CUDA_TRY(cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming));
  // the set's previous user (two calls ago) must have copied its result out before we overwrite the buffers
  if (st.used) FA_CUDA_TRY(cudaStreamWaitEvent(w.s_in, st.done, 0));

  // chunk over the flattened (b,h) axis: heads are independent and contiguous in [B,H,N,D]
  const std::vector<size_t> chunk_heads = plan_host_chunks(heads, Nq, Nkv, D, causal);
  const size_t n_chunks = chunk_heads.size();
  while (w.ev_in.size() < n_chunks) {
    cudaEvent_t e1, e2;
    FA_CUDA_TRY(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
    FA_CUDA_TRY(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
    w.ev_in.push_back(e1);
    w.ev_run.push_back(e2);
  }

  // FA_HOST_TIMING=1: print host enqueue time and the device timeline per call (diagnostic; synchronises)
  static const bool timing = std::getenv("FA_HOST_TIMING") != nullptr;
  static cudaEvent_t t_ev[2] = {nullptr, nullptr};
  static std::vector<cudaEvent_t> t_chunk;
  const auto t_cpu0 = std::chrono::steady_clock::now();
  if (timing) {
    if (t_ev[0] == nullptr) { cudaEventCreate(&t_ev[0]); cudaEventCreate(&t_ev[1]); }
    cudaEventRecord(t_ev[0], w.s_in);
  }
  const char* hq = static_cast<const char*>(q);
  const char* hk = static_cast<const char*>(k);
  const char* hv = static_cast<const char*>(v);
  char* ho = static_cast<char*>(o);
  size_t h0 = 0;
  for (size_t c = 0; c < n_chunks; h0 += chunk_heads[c], ++c) {
    const size_t nh = chunk_heads[c];
    char* dq = static_cast<char*>(st.dq) + h0 * q_head;
    char* dk = static_cast<char*>(st.dk) + h0 * kv_head;
    char* dv = static_cast<char*>(st.dv) + h0 * kv_head;
    char* dout = static_cast<char*>(st.dout) + h0 * q_head;
    FA_CUDA_TRY(cudaMemcpyAsync(dq, hq + h0 * q_head, nh * q_head, cudaMemcpyHostToDevice, w.s_in));
    FA_CUDA_TRY(cudaMemcpyAsync(dk, hk + h0 * kv_head, nh * kv_head, cudaMemcpyHostToDevice, w.s_in));
    FA_CUDA_TRY(cudaMemcpyAsync(dv, hv + h0 * kv_head, nh * kv_head, cudaMemcpyHostToDevice, w.s_in));
    FA_CUDA_TRY(cudaEventRecord(w.ev_in[c], w.s_in));
    FA_CUDA_TRY(cudaStreamWaitEvent(w.s_run, w.ev_in[c], 0));
    if (timing) {
      while (t_chunk.size() < 4 * (c + 1)) { cudaEvent_t e; cudaEventCreate(&e); t_chunk.push_back(e); }
      cudaEventRecord(t_chunk[4 * c + 0], w.s_in);
      cudaEventRecord(t_chunk[4 * c + 1], w.s_run);
    }
    // the chunk is a [1, nh, N, D] problem
    Problem p{};
    p.B = 1; p.H = static_cast<int>(nh); p.Nq = Nq; p.Nkv = Nkv; p.D = D; p.dtype = dtype;
    p.causal = causal; p.scale = scale;
    const int64_t cqs[4] = {int64_t(nh) * Nq * D, int64_t(Nq) * D, D, 1};
    const int64_t cks[4] = {int64_t(nh) * Nkv * D, int64_t(Nkv) * D, D, 1};
    if ((rc = validate(p, cqs, cks, cks, cqs))) return rc;
    float* dl = (lse != nullptr) ? st.dlse + h0 * Nq : nullptr;
    if ((rc = run_device(dq, dk, dv, dout, dl, p, w.s_run))) return rc;
    FA_CUDA_TRY(cudaEventRecord(w.ev_run[c], w.s_run));
    if (timing) cudaEventRecord(t_chunk[4 * c + 2], w.s_run);
    FA_CUDA_TRY(cudaStreamWaitEvent(w.s_out, w.ev_run[c], 0));
    FA_CUDA_TRY(cudaMemcpyAsync(ho + h0 * q_head, dout, nh * q_head, cudaMemcpyDeviceToHost, w.s_out));
    if (lse != nullptr)
      FA_CUDA_TRY(cudaMemcpyAsync(lse + h0 * Nq, dl, nh * Nq * sizeof(float),
                                  cudaMemcpyDeviceToHost, w.s_out));
    if (timing) cudaEventRecord(t_chunk[4 * c + 3], w.s_out);
  }
  FA_CUDA_TRY(cudaEventRecord(st.done, w.s_out));
  st.used = true;
  ++w.calls;
  if (timing) {
    const auto t_cpu1 = std::chrono::steady_clock::now();
    cudaEventRecord(t_ev[1], w.s_out);
    cudaStreamSynchronize(w.s_out);
    const auto t_cpu2 = std::chrono::steady_clock::now();
    float dev_ms = 0.f;
    cudaEventElapsedTime(&dev_ms, t_ev[0], t_ev[1]);
    fprintf(stderr, "[fa_fwd_sm100_host] N=%d chunks=%zu enqueue %.3f ms, device span %.3f ms, total %.3f ms\n",
            Nq, n_chunks, std::chrono::duration<double, std::milli>(t_cpu1 - t_cpu0).count(), dev_ms,
            std::chrono::duration<double, std::milli>(t_cpu2 - t_cpu0).count());
    for (size_t c = 0; c < n_chunks; ++c) {
      float t[4];
      for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], t_ev[0], t_chunk[4 * c + i]);
      fprintf(stderr, "    chunk %zu: H2D done %.3f, kernel start %.3f, kernel done %.3f, D2H done %.3f\n", c,
              t[0], t[1], t[2], t[3]);
    }
  }
  return FA_OK;
}

int host_call(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Nq,
              int Nkv, int D, int dtype, int causal, float scale, bool wait) {
  if (q == nullptr || k == nullptr || v == nullptr || o == nullptr)
    return fail(FA_ERR_INVALID_ARG, "q, k, v and o must not be null");
  Problem chk{};
  chk.B = B; chk.H = H; chk.Nq = Nq; chk.Nkv = Nkv; chk.D = D; chk.dtype = dtype;
  chk.causal = causal; chk.scale = scale;
  const int64_t qs[4] = {int64_t(H) * Nq * D, int64_t(Nq) * D, D, 1};
  const int64_t ks[4] = {int64_t(H) * Nkv * D, int64_t(Nkv) * D, D, 1};
  int rc = validate(chk, qs, ks, ks, qs);
  if (rc) return rc;
  int dev;
  if ((rc = check_device(&dev))) return rc;
  if (dev >= 64) return fail(FA_ERR_UNSUPPORTED, "device ordinal >= 64");
  HostWs& w = g_ws[dev];
  std::lock_guard<std::mutex> lk(w.mu);
  rc = host_enqueue(w, q, k, v, o, lse, B, H, Nq, Nkv, D, dtype, causal, scale);
  if (rc) {
    // copies already enqueued still read and write the caller's buffers: drain before reporting the error,
    // and never leave half a call behind for the next one
    const std::string keep = g_err;
    host_drain(w);
    g_err = keep;
    return rc;
  }
  if (wait) FA_CUDA_TRY(cudaStreamSynchronize(w.s_out));
  return FA_OK;
}

}  // namespace

// =================================================================================================
// exported C ABI
// =================================================================================================
extern "C" {

#ifdef FA_TRACE
void fa_trace_set(void* buf) { g_trace = static_cast<unsigned long long*>(buf); }
#endif

int fa_abi_version(void) { return FA_ABI_VERSION; }

const char* fa_last_error(void) { return g_err.c_str(); }

uint64_t fa_launch_count(void) { return g_launches.load(); }

int fa_set_bwd_kernel(int kernel) { return g_bwd_kernel.exchange(kernel); }

int fa_set_kernel(int kernel) {
  if (kernel < FA_KERNEL_AUTO || kernel > FA_KERNEL_WS3 || kernel == 3 || kernel == 8) return -FA_ERR_INVALID_ARG;
  return g_forced_kernel.exchange(kernel);
}

int fa_host_plan_chunks(int B, int H, int Nq, int Nkv, int D, int causal, int* out, int cap) {
  if (B < 1 || H < 1 || Nq < 1 || Nkv < 1 || D < 1 || out == nullptr || cap < 1)
    return -fail(FA_ERR_INVALID_ARG, "fa_host_plan_chunks: bad argument");
  const std::vector<size_t> c = plan_host_chunks(size_t(B) * H, Nq, Nkv, D, causal);
  for (size_t i = 0; i < c.size() && i < size_t(cap); ++i) out[i] = static_cast<int>(c[i]);
  return static_cast<int>(c.size());
}

int fa_set_wide_pairs(int enable) { return g_wide_pairs.exchange(enable ? 1 : 0); }

int fa_set_pdl(int enable) { return g_pdl.exchange(enable ? 1 : 0); }

int fa_select_kernel(int B, int H, int Nq, int Nkv, int D, const int64_t q_strides[4],
                     const int64_t k_strides[4], const int64_t v_strides[4],
                     const int64_t o_strides[4], int dtype, int causal, float scale) {
  Problem p{};
  p.B = B; p.H = H; p.Nq = Nq; p.Nkv = Nkv; p.D = D; p.dtype = dtype; p.causal = causal;
  p.scale = scale;
  int rc = validate(p, q_strides, k_strides, v_strides, o_strides);
  if (rc) return -rc;
  int kernel;
  if ((rc = resolve_kernel(p, nullptr, nullptr, nullptr, nullptr, &kernel))) return -rc;
  return kernel;
}

int fa_fwd_sm100(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H,
                 int Nq, int Nkv, int D, const int64_t q_strides[4], const int64_t k_strides[4],
                 const int64_t v_strides[4], const int64_t o_strides[4], int dtype, int causal,
                 float scale, void* stream) {
  if (q == nullptr || k == nullptr || v == nullptr || o == nullptr)
    return fail(FA_ERR_INVALID_ARG, "q, k, v and o must not be null");
  Problem p{};
  p.B = B; p.H = H; p.Nq = Nq; p.Nkv = Nkv; p.D = D; p.dtype = dtype; p.causal = causal;
  p.scale = scale;
  int rc = validate(p, q_strides, k_strides, v_strides, o_strides);
  if (rc) return rc;
  return run_device(q, k, v, o, lse, p, static_cast<cudaStream_t>(stream));
}

int fa_fwd_sm100_host(const void* q, const void* k, const void* v, void* o, float* lse, int B,
                      int H, int Nq, int Nkv, int D, int dtype, int causal, float scale) {
  return host_call(q, k, v, o, lse, B, H, Nq, Nkv, D, dtype, causal, scale, true);
}

int fa_fwd_sm100_host_async(const void* q, const void* k, const void* v, void* o, float* lse, int B,
                            int H, int Nq, int Nkv, int D, int dtype, int causal, float scale) {
  return host_call(q, k, v, o, lse, B, H, Nq, Nkv, D, dtype, causal, scale, false);
}

int fa_host_sync(void) {
  int dev;
  int rc = check_device(&dev);
  if (rc) return rc;
  if (dev >= 64) return FA_OK;
  HostWs& w = g_ws[dev];
  std::lock_guard<std::mutex> lk(w.mu);
  if (w.s_out != nullptr) FA_CUDA_TRY(cudaStreamSynchronize(w.s_out));
  return FA_OK;
}

int fa_bwd_sm100(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                 const float* lse, void* dq, void* dk, void* dv, float* dq_accum, float* delta, int B,
                 int H, int Nq, int Nkv, int D, const int64_t q_strides[4],
                 const int64_t k_strides[4], const int64_t v_strides[4], const int64_t o_strides[4],
                 const int64_t do_strides[4], const int64_t dq_strides[4],
                 const int64_t dk_strides[4], const int64_t dv_strides[4], int dtype, int causal,
                 float scale, void* stream) {
  if (q == nullptr || k == nullptr || v == nullptr || o == nullptr || d_o == nullptr ||
      lse == nullptr || dq == nullptr || dk == nullptr || dv == nullptr || dq_accum == nullptr ||
      delta == nullptr)
    return fail(FA_ERR_INVALID_ARG, "fa_bwd_sm100: every buffer pointer must be non-null");
  if (do_strides == nullptr || dq_strides == nullptr || dk_strides == nullptr || dv_strides == nullptr)
    return fail(FA_ERR_INVALID_ARG, "stride arrays must not be null");
  Problem p{};
  p.B = B; p.H = H; p.Nq = Nq; p.Nkv = Nkv; p.D = D; p.dtype = dtype; p.causal = causal;
  p.scale = scale;
  int rc = validate(p, q_strides, k_strides, v_strides, o_strides);
  if (rc) return rc;
  int64_t dos[4], dqs[4], dks[4], dvs[4];
  memcpy(dos, do_strides, sizeof dos);
  memcpy(dqs, dq_strides, sizeof dqs);
  memcpy(dks, dk_strides, sizeof dks);
  memcpy(dvs, dv_strides, sizeof dvs);
  if (dos[3] != 1 || dqs[3] != 1 || dks[3] != 1 || dvs[3] != 1)
    return fail(FA_ERR_INVALID_ARG, "the innermost (head-dim) stride of dO, dQ, dK and dV must be 1");
  canon_strides(dos, B, H, Nq, D);
  canon_strides(dqs, B, H, Nq, D);
  canon_strides(dks, B, H, Nkv, D);
  canon_strides(dvs, B, H, Nkv, D);
  const bool tma_ok = (D % 8 == 0) && tma_ok_strides(p.qs) && tma_ok_strides(p.ks) && tma_ok_strides(p.vs) &&
                      tma_ok_strides(dos) && tma_ok_strides(dks) && tma_ok_strides(dvs) && aligned16(q) &&
                      aligned16(k) && aligned16(v) && aligned16(d_o) && aligned16(dk) && aligned16(dv);
  const bool wide_tc = tma_ok && D > 128 && D <= 256 && tma_ok_strides(dqs) && aligned16(dq) &&
                       g_bwd_kernel.load() != 3;  // fa_set_bwd_kernel(3) forces the CUDA-core kernels (tests)
  const bool tc = (tma_ok && D <= 128 && aligned16(dq_accum)) || wide_tc;
  if (!tc && D > fa::kSimtMaxD) return fail(FA_ERR_UNSUPPORTED, "head dim > 1024 is not supported");
  int dev;
  if ((rc = check_device(&dev))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  const int64_t rows = int64_t(B) * H * Nq;
  const unsigned blocks = static_cast<unsigned>((rows + 7) / 8);
  // 1. delta = rowsum(dO o O), dq_accum = 0 (the generic path does not use the accumulator)
  const int acc_ld = (tc && !wide_tc) ? D : 0;
  // 16-byte vector version when every row of O and dO (and of the accumulator) can be read that way
  const bool vec_rows = (D % 8 == 0) && aligned16(o) && aligned16(d_o) && aligned16(dq_accum) && tma_ok_strides(p.os) &&
                        tma_ok_strides(dos) && (acc_ld % 4 == 0);
  if (vec_rows) {
    const bool narrow = D <= 128;  // 16 lanes per row, two rows per warp
    const unsigned vblocks = static_cast<unsigned>((rows + (narrow ? 15 : 7)) / (narrow ? 16 : 8));
#define FA_DELTA_VEC(T, LANES)                                                                                  \
  fa::fa_bwd_delta_vec_kernel<T, LANES><<<vblocks, 256, 0, st>>>(static_cast<const T*>(o), static_cast<const T*>(d_o), \
                                                                 delta, dq_accum, B, H, Nq, D, acc_ld, p.os[0], \
                                                                 p.os[1], p.os[2], dos[0], dos[1], dos[2])
    if (dtype == FA_DTYPE_BF16) {
      if (narrow) FA_DELTA_VEC(__nv_bfloat16, 16); else FA_DELTA_VEC(__nv_bfloat16, 32);
    } else {
      if (narrow) FA_DELTA_VEC(__half, 16); else FA_DELTA_VEC(__half, 32);
    }
#undef FA_DELTA_VEC
  } else if (dtype == FA_DTYPE_BF16) {
    fa::fa_bwd_delta_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(o), static_cast<const __nv_bfloat16*>(d_o), delta, dq_accum,
        B, H, Nq, D, acc_ld, p.os[0], p.os[1], p.os[2], dos[0], dos[1], dos[2]);
  } else {
    fa::fa_bwd_delta_kernel<__half><<<blocks, 256, 0, st>>>(
        static_cast<const __half*>(o), static_cast<const __half*>(d_o), delta, dq_accum, B, H, Nq, D,
        acc_ld, p.os[0], p.os[1], p.os[2], dos[0], dos[1], dos[2]);
  }
  FA_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);

  if (!tc) {
    // generic CUDA-core path (fa_bwd_simt.cuh): head dims 129..1024, head dims not a multiple of 8,
    // unaligned pointers or strides - everything the forward's generic kernel accepts
    fa::SimtBwdParams sp;
    sp.q = q; sp.k = k; sp.v = v; sp.d_o = d_o; sp.dq = dq; sp.dk = dk; sp.dv = dv;
    sp.lse = lse; sp.delta = delta;
    sp.B = B; sp.H = H; sp.Nq = Nq; sp.Nkv = Nkv; sp.D = D;
    memcpy(sp.qs, p.qs, sizeof sp.qs);
    memcpy(sp.ks, p.ks, sizeof sp.ks);
    memcpy(sp.vs, p.vs, sizeof sp.vs);
    memcpy(sp.dos, dos, sizeof dos);
    memcpy(sp.dqs, dqs, sizeof dqs);
    memcpy(sp.dks, dks, sizeof dks);
    memcpy(sp.dvs, dvs, sizeof dvs);
    sp.causal = causal;
    sp.scale = scale;
    sp.scale_log2 = scale * 1.4426950408889634f;
    const int smem = fa::kSimtWarps * 2 * D * static_cast<int>(sizeof(float));
    dim3 gq((Nq + fa::kSimtWarps - 1) / fa::kSimtWarps, H, B), gk((Nkv + fa::kSimtWarps - 1) / fa::kSimtWarps, H, B);
    if (dtype == FA_DTYPE_BF16) {
      fa::fa_bwd_simt_dq_kernel<__nv_bfloat16><<<gq, fa::kSimtWarps * 32, smem, st>>>(sp);
      fa::fa_bwd_simt_dkv_kernel<__nv_bfloat16><<<gk, fa::kSimtWarps * 32, smem, st>>>(sp);
    } else {
      fa::fa_bwd_simt_dq_kernel<__half><<<gq, fa::kSimtWarps * 32, smem, st>>>(sp);
      fa::fa_bwd_simt_dkv_kernel<__half><<<gk, fa::kSimtWarps * 32, smem, st>>>(sp);
    }
    FA_CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(2, std::memory_order_relaxed);
    return FA_OK;
  }

  // 2. main kernel
  BwdMaps m;
  BwdPlan key{};
  const void* ptrs[8] = {q, k, v, d_o, dk, dv, wide_tc ? nullptr : dq_accum, wide_tc ? dq : nullptr};
  memcpy(key.ptr, ptrs, sizeof ptrs);
  key.B = B; key.H = H; key.Nq = Nq; key.Nkv = Nkv; key.D = D; key.dtype = dtype; key.device = dev;
  memcpy(key.st[0], p.qs, sizeof p.qs);
  memcpy(key.st[1], p.ks, sizeof p.ks);
  memcpy(key.st[2], p.vs, sizeof p.vs);
  memcpy(key.st[3], dos, sizeof dos);
  memcpy(key.st[4], dks, sizeof dks);
  memcpy(key.st[5], dvs, sizeof dvs);
  memcpy(key.st[6], dqs, sizeof dqs);
  if ((rc = g_bwd_plans.get(key, &m))) return rc;
  if (wide_tc) {
    fa::BwdWideParams wp{lse, delta, Nq, Nkv, H, scale * 1.4426950408889634f, 1.f};
    return dispatch_bwd_wide(m, wp, B, H, Nq, Nkv, D, dtype, causal, scale, dev, st);
  }
  fa::BwdParams bp{lse, delta, dq_accum, Nq, Nkv, H, D, scale * 1.4426950408889634f, scale, 1, nullptr};
#ifdef FA_TRACE
  bp.trace = g_trace;
#endif
  if ((rc = dispatch_bwd_tc(m, bp, B, H, Nkv, D, dtype, causal, dev, st))) return rc;

  // 3. dQ = scale * dq_accum
  if ((D % 8 == 0) && aligned16(dq) && tma_ok_strides(dqs)) {
    const unsigned vblocks = static_cast<unsigned>((rows + 15) / 16);  // D <= 128 here: 16 lanes per row
    if (dtype == FA_DTYPE_BF16)
      fa::fa_bwd_dq_convert_vec_kernel<__nv_bfloat16, 16><<<vblocks, 256, 0, st>>>(
          dq_accum, static_cast<__nv_bfloat16*>(dq), B, H, Nq, D, D, dqs[0], dqs[1], dqs[2], scale);
    else
      fa::fa_bwd_dq_convert_vec_kernel<__half, 16><<<vblocks, 256, 0, st>>>(
          dq_accum, static_cast<__half*>(dq), B, H, Nq, D, D, dqs[0], dqs[1], dqs[2], scale);
  } else if (dtype == FA_DTYPE_BF16) {
    fa::fa_bwd_dq_convert_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
        dq_accum, static_cast<__nv_bfloat16*>(dq), B, H, Nq, D, D, dqs[0], dqs[1], dqs[2], scale);
  } else {
    fa::fa_bwd_dq_convert_kernel<__half><<<blocks, 256, 0, st>>>(
        dq_accum, static_cast<__half*>(dq), B, H, Nq, D, D, dqs[0], dqs[1], dqs[2], scale);
  }
  FA_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return FA_OK;
}

int fa_host_workspace_release(void) {
  int dev;
  int rc = check_device(&dev);
  if (rc) return rc;
  if (dev >= 64) return FA_OK;
  {
    std::lock_guard<std::mutex> lk(g_ws[dev].mu);
    ws_release(g_ws[dev]);
  }
  release_sk_workspaces(dev);
  return FA_OK;
}

int fa_umma_selftest(const void* a, const void* b, float* out, int dtype, int mode, uint32_t lbo,
                     uint32_t sbo, void* stream) {
  if (a == nullptr || b == nullptr || out == nullptr)
    return fail(FA_ERR_INVALID_ARG, "a, b and out must not be null");
  if (mode < 0 || mode > 5) return fail(FA_ERR_INVALID_ARG, "mode must be 0..5");
  if (dtype != FA_DTYPE_F16 && dtype != FA_DTYPE_BF16)
    return fail(FA_ERR_INVALID_ARG, "dtype must be FA_DTYPE_F16 or FA_DTYPE_BF16");
  int dev;
  int rc = check_device(&dev);
  if (rc) return rc;
  if (lbo == 0 && sbo == 0) { lbo = 16384; sbo = 1024; }
  if (mode == 5) lbo = 0xE5u;  // marks the extra un-swizzled k-step for the kernel (mode 5 = mode 0 + that step)
  const int64_t st[4] = {128 * 128, 128 * 128, 128, 1};
  CUtensorMap ma, mb;
  if ((rc = make_map(&ma, a, 1, 1, 128, 128, st, dtype, 128))) return rc;
  if ((rc = make_map(&mb, b, 1, 1, 128, 128, st, dtype, 128))) return rc;
  const int smem = 65536 + 128 + 1024 + 8192;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == FA_DTYPE_BF16) {
    if ((rc = set_smem(fa::umma_probe_kernel<true>, smem))) return rc;
    fa::umma_probe_kernel<true><<<1, 128, smem, s>>>(ma, mb, static_cast<const uint16_t*>(a), out,
                                                     mode, lbo, sbo);
  } else {
    if ((rc = set_smem(fa::umma_probe_kernel<false>, smem))) return rc;
    fa::umma_probe_kernel<false><<<1, 128, smem, s>>>(ma, mb, static_cast<const uint16_t*>(a), out,
                                                      mode, lbo, sbo);
  }
  FA_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return FA_OK;
}

int fa_umma2_selftest(const void* a, const void* b, float* out, int dtype, int mode, void* stream) {
  if (a == nullptr || b == nullptr || out == nullptr)
    return fail(FA_ERR_INVALID_ARG, "a, b and out must not be null");
  if (mode < 0 || mode > 1) return fail(FA_ERR_INVALID_ARG, "mode must be 0 or 1");
  if (dtype != FA_DTYPE_F16 && dtype != FA_DTYPE_BF16)
    return fail(FA_ERR_INVALID_ARG, "dtype must be FA_DTYPE_F16 or FA_DTYPE_BF16");
  int dev;
  int rc = check_device(&dev);
  if (rc) return rc;
  const int smem = 65536 + 128 + 1024;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == FA_DTYPE_BF16) {
    if ((rc = set_smem(fa::umma2_probe_kernel<true>, smem))) return rc;
    fa::umma2_probe_kernel<true><<<2, 128, smem, s>>>(static_cast<const uint16_t*>(a),
                                                      static_cast<const uint16_t*>(b), out, mode);
  } else {
    if ((rc = set_smem(fa::umma2_probe_kernel<false>, smem))) return rc;
    fa::umma2_probe_kernel<false><<<2, 128, smem, s>>>(static_cast<const uint16_t*>(a),
                                                       static_cast<const uint16_t*>(b), out, mode);
  }
  FA_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return FA_OK;
}

}  // extern "C"
