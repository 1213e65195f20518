// common.cuh -- shared device/host helpers for libvitgan_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vitgan_b200.h"

namespace vg {

// ---------------------------------------------------------------- error state (thread local: backward runs on autograd threads)
void set_error(const char* fmt, ...);
int check_launch(const char* what);   // cudaPeekAtLastError -> vg_status

#define VG_REQUIRE(cond, code, ...)            \
  do {                                         \
    if (!(cond)) {                             \
      vg::set_error(__VA_ARGS__);              \
      return (code);                           \
    }                                          \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
int num_sms();

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Kernels launched through launch_pdl() may begin (prologue: barrier init, TMEM alloc, descriptor prefetch) while the
// previous kernel of the stream is still draining.  Contract: every such kernel executes pdl_wait() before its first
// access to global memory and pdl_trigger() as early as it likes.  Launching a kernel WITHOUT the attribute after one
// that triggered early is always safe (plain stream order).  VG_PDL=0 disables the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---------------------------------------------------------------- dtype helpers
typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4-element vector load/store with conversion to/from float (16 B for fp32, 8 B for bf16)
template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec4<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[4]) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------- activations (fp32 math)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float dgelu_erf(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// erf for the bf16 tensor-core epilogues: Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below bf16 resolution):
// 7 FMA-pipe instructions + 2 MUFU (rcp, ex2) instead of libdevice erff's ~25 -- the GELU epilogue of a 128x256 tile was
// issue-bound (5.3 us of epilogue against 3.3 us of MMA per tile at E = 768).  The fp32 parity path keeps erff.
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * ax * ax));
  return copysignf(fmaf(-p, e, 1.0f), x);
}
// erf(x / sqrt2) as ONE MUFU: tanh(x * (A0 + A1 x^2 + A2 x^4)), minimax-fitted on [0, 6] (|erf error| <= 3.7e-5, i.e. |GELU error|
// <= 5.5e-5 + the 2^-11 of tanh.approx -- below half a bf16 ulp of any |GELU| > 0.03; the far negative tail, |GELU| < 5e-3, keeps
// an ABSOLUTE error < 8e-4).  7 instructions per GELU instead of 15 (A&S erf_fast) or ~25 (erff): the fc1 / dgrad-fc2 epilogues
// were issue-bound at 2.7-3.7 us per 128x128 tile (profiles/r02_gemm_timeline.txt).
// The argument is clamped to [-8, 8] (erf(8/sqrt2) == 1 in fp32): the fitted polynomial turns over beyond |x| ~ 11.
__device__ __forceinline__ float erf_sqrt2_tanh(float x) {
  const float xc = fminf(fmaxf(x, -8.0f), 8.0f);
  const float x2 = xc * xc;
  const float poly = fmaf(fmaf(-3.1580704e-4f, x2, 3.6798256e-2f), x2, 7.9771783e-1f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(xc * poly));
  return t;
}
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float hx = 0.5f * x;
  return fmaf(hx, erf_sqrt2_tanh(x), hx);
}
// GELU'(x) = Phi(x) + x phi(x), as the exact derivative of the approximation above (no second MUFU for the exp):
// 0.5 (1 + t) + 0.5 x (1 - t^2) u'(x), u' = A0 + 3 A1 x^2 + 5 A2 x^4; |error| <= 1.4e-4 against the erf form.
__device__ __forceinline__ float dgelu_erf_fast(float x) {
  const float xc = fminf(fmaxf(x, -8.0f), 8.0f);
  const float x2 = xc * xc;
  const float poly = fmaf(fmaf(-3.1580704e-4f, x2, 3.6798256e-2f), x2, 7.9771783e-1f);
  const float dpoly = fmaf(fmaf(-1.5790352e-3f, x2, 1.1039477e-1f), x2, 7.9771783e-1f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(xc * poly));
  const float cdf = fmaf(0.5f, t, 0.5f);
  return fmaf(0.5f * xc * dpoly, fmaf(-t, t, 1.0f), cdf);
}

// epilogue activation shared by the SIMT and tcgen05 GEMMs; `a` is the aux value (ignored when unused)
__device__ __forceinline__ float apply_act(int act, float v, float a, float prm) {
  switch (act) {
    case VG_ACT_GELU: return gelu_erf(v);
    case VG_ACT_TANH: return tanhf(v);
    case VG_ACT_SIN: return sinf(prm * v);
    case VG_ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
    case VG_ACT_MUL_DGELU: return v * dgelu_erf(a);
    case VG_ACT_MUL_DTANH: return v * (1.0f - a * a);
    case VG_ACT_MUL_DSIN: return v * prm * cosf(prm * a);
    case VG_ACT_MUL_DSIGMOID: return v * a * (1.0f - a);
    default: return v;
  }
}
// compile-time activation; FAST selects hardware approximations (bf16 tensor-core path, 2e-2 tolerance),
// otherwise the accurate libdevice functions (fp32 parity path, 1e-4 tolerance)
template <int ACT, bool FAST>
__device__ __forceinline__ float act_t(float v, float a, float prm) {
  if (ACT == VG_ACT_GELU) return FAST ? gelu_erf_fast(v) : gelu_erf(v);
  if (ACT == VG_ACT_TANH) {
    if (FAST) { float t; asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(v)); return t; }
    return tanhf(v);
  }
  if (ACT == VG_ACT_SIN) return FAST ? __sinf(prm * v) : sinf(prm * v);
  if (ACT == VG_ACT_SIGMOID) return FAST ? __fdividef(1.0f, 1.0f + __expf(-v)) : 1.0f / (1.0f + expf(-v));
  if (ACT == VG_ACT_MUL_DGELU) return v * (FAST ? dgelu_erf_fast(a) : dgelu_erf(a));
  if (ACT == VG_ACT_MUL_DTANH) return v * (1.0f - a * a);
  if (ACT == VG_ACT_MUL_DSIN) return v * prm * (FAST ? __cosf(prm * a) : cosf(prm * a));
  if (ACT == VG_ACT_MUL_DSIGMOID) return v * a * (1.0f - a);
  return v;
}
// run `body.template operator()<ACT>()` for the runtime activation code (one switch per tile, not per element)
#define VG_ACT_SWITCH(act, CALL)                                   \
  switch (act) {                                                   \
    case VG_ACT_GELU: { constexpr int ACT = VG_ACT_GELU; CALL; } break;                 \
    case VG_ACT_TANH: { constexpr int ACT = VG_ACT_TANH; CALL; } break;                 \
    case VG_ACT_SIN: { constexpr int ACT = VG_ACT_SIN; CALL; } break;                   \
    case VG_ACT_SIGMOID: { constexpr int ACT = VG_ACT_SIGMOID; CALL; } break;           \
    case VG_ACT_MUL_DGELU: { constexpr int ACT = VG_ACT_MUL_DGELU; CALL; } break;       \
    case VG_ACT_MUL_DTANH: { constexpr int ACT = VG_ACT_MUL_DTANH; CALL; } break;       \
    case VG_ACT_MUL_DSIN: { constexpr int ACT = VG_ACT_MUL_DSIN; CALL; } break;         \
    case VG_ACT_MUL_DSIGMOID: { constexpr int ACT = VG_ACT_MUL_DSIGMOID; CALL; } break; \
    default: { constexpr int ACT = VG_ACT_NONE; CALL; } break;                          \
  }

__host__ __device__ __forceinline__ bool act_needs_aux(int act) { return act >= VG_ACT_MUL_DGELU; }

// row remaps of the GEMM epilogue (see vg_gemm_args)
__device__ __forceinline__ int64_t out_row(int m, int c_row_group) {
  return c_row_group > 0 ? (int64_t)m + m / c_row_group + 1 : (int64_t)m;
}
__device__ __forceinline__ int64_t res_row(int m, int mod, int off) {
  return mod > 0 ? (int64_t)(m % mod) + off : (int64_t)m;
}

// ---------------------------------------------------------------- cross-CTA column reduction with bounded contention
// gridDim CTAs each hold `ncols` partial sums in smem.  Same-address global atomics serialise at the L2 (~70 ns each,
// measured: 296 CTAs on one address = 20 us), so the partials go to one of R replicated accumulator rows
// (row = blockIdx % R -> contention gridDim/R), and the LAST CTA to finish (ticket counter) folds the R rows into
// `out` (+=) in a fixed order and re-zeroes them.  ws = persistent zero-initialised [R][ncols] scratch, `counter`
// a zero-initialised ticket; both are left zeroed for the next call.  Call with all threads of the CTA.
__device__ __forceinline__ void cta_replica_reduce(float* __restrict__ ws, int R, unsigned* __restrict__ counter,
                                                   const float* __restrict__ smem_vals, int ncols, float* const* outs,
                                                   const int* out_offsets, int n_outs) {
  __shared__ unsigned s_ticket;
  float* mine = ws + (size_t)(blockIdx.x % R) * ncols;
  for (int i = threadIdx.x; i < ncols; i += blockDim.x) {
    const float v = smem_vals[i];
    if (v != 0.f) atomicAdd(mine + i, v);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(counter, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  for (int i = threadIdx.x; i < ncols; i += blockDim.x) {
    float acc = 0.f;
    for (int r0 = 0; r0 < R; r0 += 16) {        // 16 independent L2 loads in flight (a dependent chain costs ~0.5 us per replica)
      float t[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) t[j] = (r0 + j < R) ? __ldcg(ws + (size_t)(r0 + j) * ncols + i) : 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) acc += t[j];
#pragma unroll
      for (int j = 0; j < 16; ++j) if (r0 + j < R) ws[(size_t)(r0 + j) * ncols + i] = 0.f;
    }
    int o = 0;
    while (o + 1 < n_outs && i >= out_offsets[o + 1]) ++o;       // which output array this column belongs to
    if (outs[o] != nullptr) outs[o][i - out_offsets[o]] += acc;
  }
  if (threadIdx.x == 0) *counter = 0u;
}

// launchers implemented per translation unit
int gemm_simt_launch(const vg_gemm_args& a, cudaStream_t st);
int gemm_tc_launch(const vg_gemm_args& a, cudaStream_t st);
bool gemm_tc_supported(const vg_gemm_args& a, const char** why);
void gemm_tc_set_trace(unsigned long long* p);

}  // namespace vg
