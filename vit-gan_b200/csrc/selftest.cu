// selftest.cu -- torch-free self test of the tcgen05 GEMM against the CUDA-core GEMM of this library.
// vg_selftest_tcgen05 is exported from the .so; `vg_selftest` (main below, built with -DVG_SELFTEST_MAIN)
// is the bring-up binary run on the GPU box.
#include <math.h>
#include <stdlib.h>
#include <vector>

#include "common.cuh"

using namespace vg;

namespace {
__global__ void fill_bf16(bf16* p, int64_t n, uint32_t seed, float scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    p[i] = __float2bfloat16_rn(((float)(h & 0xFFFF) / 65536.0f - 0.5f) * scale);
  }
}
}  // namespace

extern "C" int vg_selftest_tcgen05(int M, int N, int K, int trans_a, int trans_b, float tol, float* max_err_out) {
  const int64_t a_rows = trans_a ? K : M, a_cols = trans_a ? M : K;
  const int64_t b_rows = trans_b ? N : K, b_cols = trans_b ? K : N;
  const int64_t lda = (a_cols + 7) / 8 * 8, ldb = (b_cols + 7) / 8 * 8;
  bf16 *A = nullptr, *B = nullptr;
  float *C1 = nullptr, *C2 = nullptr;
  const size_t cbytes = sizeof(float) * (size_t)M * N;
  if (cudaMalloc(&A, sizeof(bf16) * a_rows * lda) || cudaMalloc(&B, sizeof(bf16) * b_rows * ldb) || cudaMalloc(&C1, cbytes) ||
      cudaMalloc(&C2, cbytes)) {
    set_error("selftest: cudaMalloc failed");
    return VG_ERR_LAUNCH;
  }
  fill_bf16<<<256, 256>>>(A, a_rows * lda, 17u, 2.0f);
  fill_bf16<<<256, 256>>>(B, b_rows * ldb, 91u, 2.0f);
  cudaMemset(C1, 0, cbytes);
  cudaMemset(C2, 0, cbytes);
  vg_gemm_args g = {};
  g.ab_dtype = VG_BF16; g.c_dtype = VG_F32; g.trans_a = trans_a; g.trans_b = trans_b; g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.ldc = N;
  g.accumulate = (trans_a && !trans_b) ? 1 : 0;       // weight-gradient form exercises split-K + atomics
  int rc;
  g.path = VG_GEMM_TCGEN05; g.C = C1;
  rc = vg_gemm(&g, nullptr);
  if (rc == VG_OK) { g.path = VG_GEMM_SIMT; g.C = C2; rc = vg_gemm(&g, nullptr); }
  cudaError_t e = cudaDeviceSynchronize();
  float max_err = INFINITY;
  if (rc == VG_OK && e == cudaSuccess) {
    std::vector<float> h1((size_t)M * N), h2((size_t)M * N);
    cudaMemcpy(h1.data(), C1, cbytes, cudaMemcpyDeviceToHost);
    cudaMemcpy(h2.data(), C2, cbytes, cudaMemcpyDeviceToHost);
    double num = 0, den = 0;
    for (size_t i = 0; i < h1.size(); ++i) { num = fmax(num, fabs((double)h1[i] - h2[i])); den = fmax(den, fabs((double)h2[i])); }
    max_err = (float)(num / (den > 0 ? den : 1));
  } else if (e != cudaSuccess) {
    set_error("selftest: device error: %s", cudaGetErrorString(e));
    rc = VG_ERR_LAUNCH;
  }
  cudaFree(A); cudaFree(B); cudaFree(C1); cudaFree(C2);
  if (max_err_out) *max_err_out = max_err;
  if (rc != VG_OK) return rc;
  if (!(max_err <= tol)) { set_error("selftest: max rel err %g > tol %g", max_err, tol); return VG_ERR_LAUNCH; }
  return VG_OK;
}

// average device time per launch (us) of `iters` back-to-back launches (no host work in between)
extern "C" float vg_time_gemm(int path, int M, int N, int K, int trans_a, int trans_b, int c_f32, int act, int iters) {
  const int64_t a_rows = trans_a ? K : M, a_cols = trans_a ? M : K;
  const int64_t b_rows = trans_b ? N : K, b_cols = trans_b ? K : N;
  const int64_t lda = (a_cols + 7) / 8 * 8, ldb = (b_cols + 7) / 8 * 8;
  bf16 *A = nullptr, *B = nullptr; void *C = nullptr, *P = nullptr; float* bias = nullptr;
  const size_t cb = (c_f32 ? 4 : 2) * (size_t)M * N;
  cudaMalloc(&A, sizeof(bf16) * a_rows * lda); cudaMalloc(&B, sizeof(bf16) * b_rows * ldb); cudaMalloc(&C, cb); cudaMalloc(&P, cb);
  cudaMalloc(&bias, 4 * (size_t)N);
  fill_bf16<<<256, 256>>>(A, a_rows * lda, 17u, 2.0f);
  fill_bf16<<<256, 256>>>(B, b_rows * ldb, 91u, 2.0f);
  cudaMemset(C, 0, cb); cudaMemset(bias, 0, 4 * (size_t)N);
  vg_gemm_args g = {};
  g.path = path; g.ab_dtype = VG_BF16; g.c_dtype = c_f32 ? VG_F32 : VG_BF16; g.trans_a = trans_a; g.trans_b = trans_b;
  g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = N;
  g.accumulate = (trans_a && !trans_b && c_f32) ? 1 : 0;
  if (!g.accumulate) { g.bias = bias; g.act = act % 100; if (act == VG_ACT_GELU) { g.c_pre = P; g.ldpre = N; } }
  for (int i = 0; i < 5; ++i) vg_gemm(&g, nullptr);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  int rc = 0;
  for (int i = 0; i < iters && !rc; ++i) rc = vg_gemm(&g, nullptr);
  cudaEventRecord(e1);
  cudaError_t e = cudaEventSynchronize(e1);
  float ms = -1.f;
  if (!rc && e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
  cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(P); cudaFree(bias);
  return ms < 0 ? -1.f : ms * 1000.f / iters;
}

#ifdef VG_SELFTEST_MAIN
#include <string.h>
#include <unistd.h>
int main(int argc, char** argv) {
  alarm(120);   // never hang a GPU box
  struct Case { int M, N, K, ta, tb; const char* name; };
  const Case cases[] = {
      {128, 128, 64, 0, 1, "fwd  1 tile 1 kblock"},
      {128, 128, 256, 0, 1, "fwd  1 tile 4 kblocks"},
      {33280, 384, 128, 0, 1, "fwd  C2 qkv"},
      {1000, 200, 72, 0, 1, "fwd  ragged M/N/K"},
      {128, 128, 64, 0, 0, "dgrad 1 tile (B MN-major)"},
      {33280, 128, 384, 0, 0, "dgrad C2 qkv"},
      {1000, 200, 72, 0, 0, "dgrad ragged"},
      {128, 128, 64, 1, 0, "wgrad 1 tile (A,B MN-major)"},
      {384, 128, 33280, 1, 0, "wgrad C2 qkv split-K"},
      {200, 72, 1000, 1, 0, "wgrad ragged"},
      {128, 128, 64, 1, 1, "A MN-major, B K-major"},
  };
  if (argc > 1 && !strcmp(argv[1], "--time")) {
    struct T { int M, N, K, ta, tb, cf32, act; const char* name; };
    const T ts[] = {
        {512, 128, 128, 0, 1, 0, 0, "fwd tiny 512x128x128"},
        {33280, 384, 128, 0, 1, 0, 0, "fwd C2 qkv"},
        {33280, 256, 128, 0, 1, 0, 1, "fwd C2 fc1+gelu(+pre)"},
        {33280, 256, 128, 0, 1, 0, 0, "fc1 shape, no act"},
        {33280, 256, 128, 0, 1, 0, 101, "fc1 shape, gelu no pre"},
        {33280, 256, 128, 0, 1, 0, 102, "fc1 shape, tanh"},
        {33280, 256, 128, 0, 1, 0, 104, "fc1 shape, sigmoid"},
        {33280, 128, 256, 0, 1, 0, 0, "fwd C2 fc2"},
        {33280, 128, 384, 0, 0, 0, 0, "dgrad C2 qkv"},
        {384, 128, 33280, 1, 0, 1, 0, "wgrad C2 qkv"},
        {8192, 8192, 8192, 0, 1, 0, 0, "fwd 8192^3"},
        {65792, 2304, 768, 0, 1, 0, 0, "fwd C4 qkv (B=256)"},
    };
    for (const T& t : ts) {
      const int iters = (double)t.M * t.N * t.K > 1e11 ? 5 : 200;
      const float us_tc = vg_time_gemm(VG_GEMM_TCGEN05, t.M, t.N, t.K, t.ta, t.tb, t.cf32, t.act, iters);
      const float us_si = (double)t.M * t.N * t.K > 1e11 ? -1.f : vg_time_gemm(VG_GEMM_SIMT, t.M, t.N, t.K, t.ta, t.tb, t.cf32, t.act, 20);
      const double fl = 2.0 * t.M * t.N * t.K;
      printf("%-26s tcgen05 %9.2f us %8.1f TFLOP/s | simt %9.2f us\n", t.name, us_tc, fl / us_tc * 1e-6, us_si);
      fflush(stdout);
    }
    return 0;
  }
  const char* only = argc > 1 ? argv[1] : nullptr;
  int fails = 0;
  for (const Case& c : cases) {
    if (only && !strstr(c.name, only)) continue;
    float err = -1.f;
    const int rc = vg_selftest_tcgen05(c.M, c.N, c.K, c.ta, c.tb, 1e-3f, &err);
    printf("%-34s M=%-6d N=%-4d K=%-6d rc=%d max_rel_err=%.3e %s %s\n", c.name, c.M, c.N, c.K, rc, err, rc == 0 ? "PASS" : "FAIL",
           rc == 0 ? "" : vg_last_error());
    fflush(stdout);
    if (rc) ++fails;
    if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("device in error state, stopping\n"); return 2; }
  }
  printf("selftest: %d failure(s)\n", fails);
  return fails ? 1 : 0;
}
#endif
