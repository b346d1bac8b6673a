/* A plain-C consumer of the drop-in boundary: includes include/fa_fwd_sm100.h, links libfa_fwd_sm100.so
 * and nothing else (no CUDA headers, no torch, no Python).  It is what a maintainer of the reference
 * would write to call the B200 path from host.cpp's side of the fence (host.cpp:30-45 `forward`).
 *
 *   consumer --abi      print the ABI version and the kernel fa_select_kernel() picks (no GPU needed)
 *   consumer            run BASELINE config 1 (fp16, B=1 H=2 N=128 D=64) and a ragged bf16 causal case
 *                       through fa_fwd_sm100_host() and compare with a double-precision softmax(QK^T)V
 *                       computed here; also checks the error contract.  Exit code 0 = all good.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fa_fwd_sm100.h"

static uint16_t f2h(float f) { /* round-to-nearest-even float -> IEEE half, normal range only */
  uint32_t x;
  memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  int32_t e = (int32_t)((x >> 23) & 0xff) - 127 + 15;
  uint32_t m = x & 0x7fffffu;
  if (e <= 0) return (uint16_t)sign;
  if (e >= 31) return (uint16_t)(sign | 0x7c00u);
  uint32_t h = sign | ((uint32_t)e << 10) | (m >> 13);
  uint32_t rem = m & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
  return (uint16_t)h;
}
static float h2f(uint16_t h) {
  uint32_t sign = ((uint32_t)h & 0x8000u) << 16, e = (h >> 10) & 0x1f, m = h & 0x3ffu, x;
  if (e == 0) {
    float v = (float)m * (1.0f / 16777216.0f); /* subnormal: m * 2^-24 */
    return sign ? -v : v;
  }
  x = sign | ((e - 15 + 127) << 23) | (m << 13);
  float f;
  memcpy(&f, &x, 4);
  return f;
}
static uint16_t f2b(float f) { /* float -> bfloat16, round-to-nearest-even */
  uint32_t x;
  memcpy(&x, &f, 4);
  x += 0x7fffu + ((x >> 16) & 1u);
  return (uint16_t)(x >> 16);
}
static float b2f(uint16_t b) {
  uint32_t x = (uint32_t)b << 16;
  float f;
  memcpy(&f, &x, 4);
  return f;
}

static uint32_t rng_state = 12345u;
static float urand(void) { /* U[0,1) like the reference's torch.rand inputs (bench_with_sdpa.py:207-209) */
  rng_state = rng_state * 1664525u + 1013904223u;
  return (float)(rng_state >> 8) * (1.0f / 16777216.0f);
}

static int run_case(int B, int H, int Nq, int Nkv, int D, int dtype, int causal, double tol) {
  size_t nq = (size_t)B * H * Nq * D, nk = (size_t)B * H * Nkv * D;
  uint16_t *q = malloc(nq * 2), *k = malloc(nk * 2), *v = malloc(nk * 2), *o = malloc(nq * 2);
  float* lse = malloc((size_t)B * H * Nq * sizeof(float));
  for (size_t i = 0; i < nq; ++i) q[i] = dtype ? f2b(urand()) : f2h(urand());
  for (size_t i = 0; i < nk; ++i) k[i] = dtype ? f2b(urand()) : f2h(urand());
  for (size_t i = 0; i < nk; ++i) v[i] = dtype ? f2b(urand()) : f2h(urand());
  float scale = 1.0f / sqrtf((float)D);
  int rc = fa_fwd_sm100_host(q, k, v, o, lse, B, H, Nq, Nkv, D, dtype, causal, scale);
  if (rc != FA_OK) {
    fprintf(stderr, "fa_fwd_sm100_host failed (%d): %s\n", rc, fa_last_error());
    return 1;
  }
  double worst = 0.0, worst_lse = 0.0;
  double* p = malloc((size_t)Nkv * sizeof(double));
  for (int bh = 0; bh < B * H; ++bh)
    for (int i = 0; i < Nq; ++i) {
      const uint16_t* qi = q + ((size_t)bh * Nq + i) * D;
      int lim = causal ? (i + 1 < Nkv ? i + 1 : Nkv) : Nkv; /* top-left aligned, kernel_fp16.cu:396-412 */
      double mx = -1e300, sum = 0.0;
      for (int j = 0; j < lim; ++j) {
        const uint16_t* kj = k + ((size_t)bh * Nkv + j) * D;
        double s = 0.0;
        for (int d = 0; d < D; ++d) s += (double)(dtype ? b2f(qi[d]) : h2f(qi[d])) * (double)(dtype ? b2f(kj[d]) : h2f(kj[d]));
        p[j] = s * scale;
        if (p[j] > mx) mx = p[j];
      }
      for (int j = 0; j < lim; ++j) { p[j] = exp(p[j] - mx); sum += p[j]; }
      for (int d = 0; d < D; ++d) {
        double acc = 0.0;
        for (int j = 0; j < lim; ++j) acc += p[j] * (double)(dtype ? b2f(v[((size_t)bh * Nkv + j) * D + d]) : h2f(v[((size_t)bh * Nkv + j) * D + d]));
        uint16_t got = o[((size_t)bh * Nq + i) * D + d];
        double err = fabs(acc / sum - (double)(dtype ? b2f(got) : h2f(got)));
        if (err > worst) worst = err;
      }
      double l2 = (mx + log(sum)) * 1.4426950408889634; /* base-2 LSE, kernel_fp16.cu:541-542 */
      double e2 = fabs(l2 - (double)lse[(size_t)bh * Nq + i]);
      if (e2 > worst_lse) worst_lse = e2;
    }
  printf("B=%d H=%d Nq=%d Nkv=%d D=%d %s causal=%d: max|o-ref|=%.3e (tol %.1e)  max|lse-ref|=%.3e\n", B, H, Nq,
         Nkv, D, dtype ? "bf16" : "fp16", causal, worst, tol, worst_lse);
  free(p); free(q); free(k); free(v); free(o); free(lse);
  return !(worst <= tol && worst_lse <= 5e-3);
}

int main(int argc, char** argv) {
  if (fa_abi_version() != FA_ABI_VERSION) {
    fprintf(stderr, "ABI mismatch: header %d, library %d\n", FA_ABI_VERSION, fa_abi_version());
    return 2;
  }
  if (argc > 1 && strcmp(argv[1], "--abi") == 0) {
    const int64_t st[4] = {16 * 16384 * 128, 16384 * 128, 128, 1};
    printf("abi %d kernel_for_sweep_point_16384 %d\n", fa_abi_version(),
           fa_select_kernel(1, 16, 16384, 16384, 128, st, st, st, st, FA_DTYPE_F16, 0, 0.088388f));
    return 0;
  }
  int bad = 0;
  bad |= run_case(1, 2, 128, 128, 64, FA_DTYPE_F16, 0, 1e-3);    /* BASELINE config 1 */
  bad |= run_case(2, 3, 300, 77, 128, FA_DTYPE_BF16, 1, 8e-3);   /* ragged, cross-attention length, causal */
  bad |= run_case(1, 2, 200, 333, 160, FA_DTYPE_F16, 0, 1e-3);   /* head dim > 128: the wide kernel */
  /* error contract: non-zero code + message, nothing written */
  uint16_t dummy[8] = {0};
  int rc = fa_fwd_sm100_host(NULL, dummy, dummy, dummy, NULL, 1, 1, 1, 1, 8, FA_DTYPE_F16, 0, 1.0f);
  if (rc != FA_ERR_INVALID_ARG || strlen(fa_last_error()) == 0) {
    fprintf(stderr, "null q: expected FA_ERR_INVALID_ARG with a message, got %d '%s'\n", rc, fa_last_error());
    bad = 1;
  }
  rc = fa_fwd_sm100_host(dummy, dummy, dummy, dummy, NULL, 1, 1, 1, 1, 8, 7, 0, 1.0f);
  if (rc != FA_ERR_INVALID_ARG) {
    fprintf(stderr, "dtype 7: expected FA_ERR_INVALID_ARG, got %d\n", rc);
    bad = 1;
  }
  fa_host_workspace_release();
  printf(bad ? "FAILED\n" : "OK\n");
  return bad;
}
