// spectral.cu -- sigma_max of small weight matrices by power iteration.
// Replaces the 3 full torch.svd calls per attention head per forward of src/v1/attention.py:54-64
// (the reference then takes python max() over the singular values).  Memory-bound: one CTA per matrix,
// the matrix (108 x 432 fp32 = 187 KB for the v1 discriminator) is staged ONCE in shared memory and
// every iteration runs out of SMEM; algorithmic HBM bytes = rows*cols*4 per matrix per call.
// A persistent left vector u (per matrix) warm-starts the iteration across training steps.
#include "common.cuh"

namespace vg {
namespace {

constexpr int NT = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NT / 32; ++i) s += red[i];
  return s;
}

template <bool IN_SMEM>
__global__ void __launch_bounds__(NT)
sigma_max_kernel(const float* const* __restrict__ mats, int rows, int cols, float* __restrict__ u_state, int n_iters,
                 float* __restrict__ sigma_out) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float red[NT / 32];
  float* u = sm;                    // [rows]
  float* v = u + ((rows + 3) & ~3); // [cols]
  float* Ws = v + ((cols + 3) & ~3);
  const float* Wg = mats[blockIdx.x];
  const float* W = Wg;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (IN_SMEM) {
    for (int i = tid; i < rows * cols; i += NT) Ws[i] = Wg[i];
    W = Ws;
  }
  float* us = u_state + (size_t)blockIdx.x * rows;
  float nrm = 0.f;
  for (int r = tid; r < rows; r += NT) { const float x = us[r]; u[r] = x; nrm = fmaf(x, x, nrm); }
  nrm = block_sum(nrm, red);
  if (nrm == 0.f) {   // cold start: fixed, matrix-independent vector with no special symmetry
    float n2 = 0.f;
    for (int r = tid; r < rows; r += NT) { const float x = 1.0f + 0.37f * __sinf(1.7f * (float)r + 0.3f); u[r] = x; n2 = fmaf(x, x, n2); }
    nrm = block_sum(n2, red);
  }
  const float inv0 = rsqrtf(nrm);
  __syncthreads();
  for (int r = tid; r < rows; r += NT) u[r] *= inv0;
  __syncthreads();

  float sigma = 0.f;
  for (int it = 0; it < n_iters; ++it) {
    // v = W^T u  (thread per column; consecutive threads -> consecutive addresses)
    float n2 = 0.f;
    for (int c = tid; c < cols; c += NT) {
      float a = 0.f;
      for (int r = 0; r < rows; ++r) a = fmaf(W[(size_t)r * cols + c], u[r], a);
      v[c] = a;
      n2 = fmaf(a, a, n2);
    }
    n2 = block_sum(n2, red);
    const float invv = n2 > 0.f ? rsqrtf(n2) : 0.f;
    for (int c = tid; c < cols; c += NT) v[c] *= invv;
    __syncthreads();
    // u = W v  (warp per row)
    float m2 = 0.f;
    for (int r = wid; r < rows; r += NT / 32) {
      float a = 0.f;
      for (int c = lane; c < cols; c += 32) a = fmaf(W[(size_t)r * cols + c], v[c], a);
      a = warp_sum(a);
      if (lane == 0) { u[r] = a; m2 = fmaf(a, a, m2); }
    }
    m2 = block_sum(m2, red);
    sigma = sqrtf(m2);
    const float invu = sigma > 0.f ? 1.f / sigma : 0.f;
    for (int r = tid; r < rows; r += NT) u[r] *= invu;
    __syncthreads();
  }
  for (int r = tid; r < rows; r += NT) us[r] = u[r];
  if (tid == 0) sigma_out[blockIdx.x] = sigma;
}

}  // namespace
}  // namespace vg

using namespace vg;

extern "C" int vg_sigma_max(const float* const* mats, int n_mats, int rows, int cols, float* u_state, int n_iters,
                            float* sigma_out, void* stream) {
  VG_REQUIRE(n_mats > 0 && rows > 0 && cols > 0 && n_iters > 0, VG_ERR_SHAPE, "sigma_max: bad sizes");
  const size_t vec_bytes = sizeof(float) * (size_t)(((rows + 3) & ~3) + ((cols + 3) & ~3));
  const size_t full = vec_bytes + sizeof(float) * (size_t)rows * cols;
  cudaStream_t st = as_stream(stream);
  if (full <= 220 * 1024) {
    if (full > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(sigma_max_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)full);
      VG_REQUIRE(e == cudaSuccess, VG_ERR_LAUNCH, "sigma_max: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    sigma_max_kernel<true><<<n_mats, NT, full, st>>>(mats, rows, cols, u_state, n_iters, sigma_out);
  } else {
    VG_REQUIRE(vec_bytes <= 48 * 1024, VG_ERR_SHAPE, "sigma_max: vectors do not fit shared memory");
    sigma_max_kernel<false><<<n_mats, NT, vec_bytes, st>>>(mats, rows, cols, u_state, n_iters, sigma_out);
  }
  return check_launch("sigma_max");
}
