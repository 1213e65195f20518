// norm.cu -- LayerNorm and self-modulated LayerNorm (SLN), forward and backward.  HBM-bound:
// one warp per row, 4-wide vector loads, row kept in registers, warp-shuffle reductions, fp32 math.
// Algorithmic bytes: fwd 2*rows*E*sizeof(T); bwd 4*rows*E*sizeof(T) (dy, x, [dres], dx).
//   LayerNorm : src/v2/modules.py:168,172,225 ; src/v1/transformer.py:18-19 (eps 1e-5, biased var, affine)
//   SLN       : src/v1/spectral_layer_norm.py:19-20  y = gamma_s * w * LN(h) + beta_s * w
#include <stdlib.h>

#include "common.cuh"

namespace vg {
namespace {

constexpr int MAXV = 8;            // up to 8 float4 per lane -> E <= 1024 (kernels are templated on NV <= MAXV)
constexpr int WARPS = 8;           // warps per CTA
constexpr int MAXE = MAXV * 128;


template <typename T, int NV>
__device__ __forceinline__ void load_row(const T* p, int E, int lane, float (&v)[NV][4]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < E) Vec4<T>::load(p + c, v[i]);
    else { v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.f; }
  }
}
template <typename T, int NV>
__device__ __forceinline__ void store_row(T* p, int E, int lane, const float (&v)[NV][4]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < E) Vec4<T>::store(p + c, v[i]);
  }
}
template <int NV>
__device__ __forceinline__ void load_vec(const float* p, int E, int lane, float (&v)[NV][4]) {
  load_row<float, NV>(p, E, lane, v);
}

template <int NV>
__device__ __forceinline__ void row_stats(const float (&x)[NV][4], int E, int lane, float& mean, float& rstd, float eps) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (x[i][0] + x[i][1]) + (x[i][2] + x[i][3]);
  mean = warp_sum(s) / (float)E;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < E) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float d = x[i][j] - mean; q = fmaf(d, d, q); }
    }
  }
  rstd = rsqrtf(warp_sum(q) / (float)E + eps);
}

// ------------------------------------------------------------------------------------------------ LN forward
// RPI rows per warp iteration (4 for E <= 128): all loads of the batch are issued before any math so that each
// lane keeps RPI*NV 16-byte requests in flight (HBM latency, not the shuffle reductions, bounds this kernel).
template <typename T, int NV>
__global__ void __launch_bounds__(WARPS * 32)
ln_fwd_kernel(int64_t rows, int E, const T* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ beta, T* __restrict__ y, float* __restrict__ mean_out,
              float* __restrict__ rstd_out, float eps) {
  constexpr int RPI = NV == 1 ? 4 : (NV == 2 ? 2 : 1);
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * WARPS;
  float g[NV][4], b[NV][4];
  load_vec(gamma, E, lane, g);
  load_vec(beta, E, lane, b);
  for (int64_t r0 = warp * RPI; r0 < rows; r0 += nwarps * RPI) {
    float v[RPI][NV][4];
#pragma unroll
    for (int q = 0; q < RPI; ++q) {
      const int64_t r = min(r0 + q, rows - 1);
      load_row<T, NV>(x + r * E, E, lane, v[q]);
    }
#pragma unroll
    for (int q = 0; q < RPI; ++q) {
      const int64_t r = r0 + q;
      if (r >= rows) break;
      float mean, rstd;
      row_stats(v[q], E, lane, mean, rstd, eps);
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) v[q][i][j] = fmaf((v[q][i][j] - mean) * rstd, g[i][j], b[i][j]);
      store_row<T, NV>(y + r * E, E, lane, v[q]);
      if (lane == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
    }
  }
}

// flush per-warp column partials: smem atomics, then one global atomic per column per CTA
template <int NV>
__device__ __forceinline__ void flush_cols(float* sm, float* gl, const float (&acc)[NV][4], int E, int lane) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < E) {
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&sm[c + j], acc[i][j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ LN backward
template <typename T, int NV>
__global__ void __launch_bounds__(WARPS * 32)
ln_bwd_kernel(int64_t rows, int E, const T* __restrict__ dy, const T* __restrict__ x,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
              const T* __restrict__ dres, T* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
              float* __restrict__ dres_colsum, float* __restrict__ dx_colsum, float* __restrict__ ws, int ws_rows, unsigned* __restrict__ counter) {
  // dres_colsum / dx_colsum (optional): column sums of the skip-path gradient and of the produced dx -- these are the
  // bias gradients of the Linear layers on either side of the norm (fc2 / out-proj), obtained here for free.
  constexpr int RPI = NV == 1 ? 4 : (NV == 2 ? 2 : 1);
  __shared__ float s_dg[MAXE], s_db[MAXE];
  pdl_trigger();
  for (int i = threadIdx.x; i < E; i += blockDim.x) { s_dg[i] = 0.f; s_db[i] = 0.f; }
  __syncthreads();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * WARPS;
  float g[NV][4], adg[NV][4], adb[NV][4], adr[NV][4], adx[NV][4];
  load_vec(gamma, E, lane, g);
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { adg[i][j] = 0.f; adb[i][j] = 0.f; adr[i][j] = 0.f; adx[i][j] = 0.f; }
  const float invE = 1.0f / (float)E;
  for (int64_t r0 = warp * RPI; r0 < rows; r0 += nwarps * RPI) {
    float xv[RPI][NV][4], dv[RPI][NV][4], rv[RPI][NV][4];
    float mu[RPI], rs[RPI];
#pragma unroll
    for (int q = 0; q < RPI; ++q) {        // issue every load of the batch first
      const int64_t r = min(r0 + q, rows - 1);
      load_row<T, NV>(x + r * E, E, lane, xv[q]);
      load_row<T, NV>(dy + r * E, E, lane, dv[q]);
      if (dres != nullptr) load_row<T, NV>(dres + r * E, E, lane, rv[q]);
      mu[q] = mean[r]; rs[q] = rstd[r];
    }
#pragma unroll
    for (int q = 0; q < RPI; ++q) {
      const int64_t r = r0 + q;
      if (r >= rows) break;
      float c1 = 0.f, c2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xh = (xv[q][i][j] - mu[q]) * rs[q];     // padded lanes: dy = 0, gamma = 0 -> contribute nothing
          const float gd = dv[q][i][j] * g[i][j];
          adg[i][j] = fmaf(dv[q][i][j], xh, adg[i][j]);
          adb[i][j] += dv[q][i][j];
          c1 += gd;
          c2 = fmaf(gd, xh, c2);
          xv[q][i][j] = xh; dv[q][i][j] = gd;
        }
      c1 = warp_sum(c1) * invE;
      c2 = warp_sum(c2) * invE;
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float t = rs[q] * (dv[q][i][j] - c1 - xv[q][i][j] * c2);
          if (dres != nullptr) adr[i][j] += rv[q][i][j];
          dv[q][i][j] = (dres != nullptr) ? rv[q][i][j] + t : t;
          adx[i][j] += dv[q][i][j];
        }
      store_row<T, NV>(dx + r * E, E, lane, dv[q]);
    }
  }
  if (dgamma == nullptr) return;      // dx only (dgrad-only pass): no column reductions at all
  if (ws != nullptr) {
    // workspace path: [dgamma | dbeta | colsum(dres) | colsum(dx)] -> replicated accumulators, folded by the last CTA
    __shared__ float s_all[4 * MAXE];
    for (int i = threadIdx.x; i < 4 * E; i += blockDim.x) s_all[i] = 0.f;
    __syncthreads();
    flush_cols(s_all, nullptr, adg, E, lane);
    flush_cols(s_all + E, nullptr, adb, E, lane);
    flush_cols(s_all + 2 * E, nullptr, adr, E, lane);
    flush_cols(s_all + 3 * E, nullptr, adx, E, lane);
    __syncthreads();
    float* outs[4] = {dgamma, dbeta, dres_colsum, dx_colsum};
    const int offs[4] = {0, E, 2 * E, 3 * E};
    cta_replica_reduce(ws, ws_rows, counter, s_all, 4 * E, outs, offs, 4);
    return;
  }
  flush_cols(s_dg, dgamma, adg, E, lane);
  flush_cols(s_db, dbeta, adb, E, lane);
  __syncthreads();
  for (int i = threadIdx.x; i < E; i += blockDim.x) { atomicAdd(&dgamma[i], s_dg[i]); atomicAdd(&dbeta[i], s_db[i]); }
  if (dres_colsum != nullptr || dx_colsum != nullptr) {      // second round through the same smem accumulators
    __syncthreads();
    for (int i = threadIdx.x; i < E; i += blockDim.x) { s_dg[i] = 0.f; s_db[i] = 0.f; }
    __syncthreads();
    flush_cols(s_dg, dres_colsum, adr, E, lane);
    flush_cols(s_db, dx_colsum, adx, E, lane);
    __syncthreads();
    for (int i = threadIdx.x; i < E; i += blockDim.x) {
      if (dres_colsum != nullptr) atomicAdd(&dres_colsum[i], s_dg[i]);
      if (dx_colsum != nullptr) atomicAdd(&dx_colsum[i], s_db[i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ LN backward, wide rows (bf16)
// E > 256 (the scaled config's 768, v1's 384 / 432).  The generic kernel above keeps ONE row per warp in flight (3 tensors x E in
// fp32 registers + four E-wide column accumulators = 255 registers, 1 CTA / SM): 8 rows = 37 KB of loads in flight per SM, below
// the ~44 KB that 6.5 TB/s x 1 us of latency needs -- it measured 30 % of the HBM roofline at E = 768.  Here the operands stay
// PACKED (bf16) in registers until they are used and the next row of the warp is requested before the current one is reduced
// (two static buffers, loop unrolled by two), so every warp has two rows in flight; x-hat and gamma*dy are recomputed in the
// second pass instead of being kept.  Same outputs as ln_bwd_kernel (dx, dgamma, dbeta, optional colsum(dres) / colsum(dx)).
template <int NV>
struct PackedRow { uint2 x[NV], dy[NV], dr[NV]; float mu, rs; };

// COLS: also accumulate colsum(dres) and colsum(dx) (the bias gradients of the Linear layers on either side).  Without them the
// 2 x 4 NV accumulator registers they need hold a third row instead: the kernel is bound by the latency of its global loads
// (ncu: IPC 1.6, a third of the stall cycles on the L1TEX scoreboard, 8 warps per SM at 255 registers), so rows in flight are
// what it converts into bandwidth.
template <int NV, bool COLS>
__global__ void __launch_bounds__(WARPS * 32, 1)
ln_bwd_wide_kernel(int64_t rows, int E, const bf16* __restrict__ dy, const bf16* __restrict__ x,
                   const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                   const bf16* __restrict__ dres, bf16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                   float* __restrict__ dres_colsum, float* __restrict__ dx_colsum, float* __restrict__ ws, int ws_rows, unsigned* __restrict__ counter) {
  __shared__ float s_all[4 * MAXE];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * WARPS;
  const bool acc = dgamma != nullptr, has_res = dres != nullptr;
  float g[NV][4], adg[NV][4], adb[NV][4], adr[NV][4], adx[NV][4];
  load_vec(gamma, E, lane, g);
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { adg[i][j] = 0.f; adb[i][j] = 0.f; adr[i][j] = 0.f; adx[i][j] = 0.f; }
  const float invE = 1.0f / (float)E;

  auto fetch = [&](PackedRow<NV>& b, int64_t r) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < E) {
        b.x[i] = __ldg(reinterpret_cast<const uint2*>(x + r * E + c));
        b.dy[i] = __ldg(reinterpret_cast<const uint2*>(dy + r * E + c));
        b.dr[i] = has_res ? __ldg(reinterpret_cast<const uint2*>(dres + r * E + c)) : make_uint2(0u, 0u);
      } else {
        b.x[i] = b.dy[i] = b.dr[i] = make_uint2(0u, 0u);
      }
    }
    b.mu = __ldg(mean + r); b.rs = __ldg(rstd + r);
  };
  auto unpack = [](uint2 u, float (&v)[4]) {
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xFFFF0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xFFFF0000u);
  };
  auto process = [&](const PackedRow<NV>& b, int64_t r) {
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float xv[4], dv[4];
      unpack(b.x[i], xv); unpack(b.dy[i], dv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = (xv[j] - b.mu) * b.rs;              // padded lanes: dy = 0, gamma = 0 -> contribute nothing
        const float gd = dv[j] * g[i][j];
        adg[i][j] = fmaf(dv[j], xh, adg[i][j]);
        adb[i][j] += dv[j];
        c1 += gd;
        c2 = fmaf(gd, xh, c2);
      }
    }
    c1 = warp_sum(c1) * invE;
    c2 = warp_sum(c2) * invE;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      float xv[4], dv[4], rv[4], o[4];
      unpack(b.x[i], xv); unpack(b.dy[i], dv); unpack(b.dr[i], rv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = (xv[j] - b.mu) * b.rs;
        const float t = b.rs * (dv[j] * g[i][j] - c1 - xh * c2);
        o[j] = rv[j] + t;
        if (COLS) { adr[i][j] += rv[j]; adx[i][j] += o[j]; }
      }
      if (c < E) Vec4<bf16>::store(dx + r * E + c, o);
    }
  };

  PackedRow<NV> A, B;
  int64_t r = warp;
  if (COLS) {                        // two rows in flight
    if (r < rows) fetch(A, r);
    while (r < rows) {
      const int64_t rn = r + nwarps;
      if (rn < rows) fetch(B, rn);
      process(A, r);
      const int64_t r2 = rn + nwarps;
      if (r2 < rows) fetch(A, r2);
      if (rn < rows) process(B, rn);
      r = r2;
    }
  } else {                           // three rows in flight
    PackedRow<NV> Cq;
    if (r < rows) fetch(A, r);
    if (r + nwarps < rows) fetch(B, r + nwarps);
    while (r < rows) {
      const int64_t r1 = r + nwarps, r2 = r1 + nwarps, r3 = r2 + nwarps, r4 = r3 + nwarps;
      if (r2 < rows) fetch(Cq, r2);
      process(A, r);
      if (r3 < rows) fetch(A, r3);
      if (r1 < rows) process(B, r1);
      if (r4 < rows) fetch(B, r4);
      if (r2 < rows) process(Cq, r2);
      r = r3;
    }
  }
  if (!acc) return;                  // dx only (dgrad-only pass): no column reductions at all
  for (int i = threadIdx.x; i < 4 * E; i += blockDim.x) s_all[i] = 0.f;
  __syncthreads();
  flush_cols(s_all, nullptr, adg, E, lane);
  flush_cols(s_all + E, nullptr, adb, E, lane);
  if (COLS) {
    flush_cols(s_all + 2 * E, nullptr, adr, E, lane);
    flush_cols(s_all + 3 * E, nullptr, adx, E, lane);
  }
  __syncthreads();
  if (ws != nullptr) {               // replicated accumulators, folded by the last CTA (same layout as ln_bwd_kernel)
    float* outs[4] = {dgamma, dbeta, dres_colsum, dx_colsum};
    const int offs[4] = {0, E, 2 * E, 3 * E};
    cta_replica_reduce(ws, ws_rows, counter, s_all, 4 * E, outs, offs, 4);
    return;
  }
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    atomicAdd(&dgamma[i], s_all[i]);
    atomicAdd(&dbeta[i], s_all[E + i]);
    if (dres_colsum != nullptr) atomicAdd(&dres_colsum[i], s_all[2 * E + i]);
    if (dx_colsum != nullptr) atomicAdd(&dx_colsum[i], s_all[3 * E + i]);
  }
}

// ------------------------------------------------------------------------------------------------ LN backward, E = 768 (bf16): three warps per row
// ln_bwd_wide_kernel above gives every warp a whole row: 4 x 24 fp32 column accumulators + two packed rows per lane = 255
// registers, 8 warps per SM, and ~29 instructions per element because x-hat and gamma*dy cannot stay in registers between its
// two passes.  Here a row is shared by THREE warps (96 lanes x 8 columns = 768): 4 x 8 accumulators per lane, x-hat and gamma*dy
// kept between the passes, two rows per warp triple per iteration with the next pair's 16-byte loads already in flight -> ~150
// registers, 12 warps per SM (3 per scheduler instead of 2) and ~40 % fewer instructions.  The row sums c1 = mean(gamma dy),
// c2 = mean(gamma dy x-hat) meet in shared memory: one 96-thread named barrier per pair of rows (slots alternate by iteration).
// Same outputs and the same reduction tail as ln_bwd_wide_kernel.
struct TriRow { uint4 x, dy, dr; float mu, rs; };
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <bool COLS>
__global__ void __launch_bounds__(384, 1)
ln_bwd_tri_kernel(int64_t rows, const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ mean,
                  const float* __restrict__ rstd, const float* __restrict__ gamma, const bf16* __restrict__ dres, bf16* __restrict__ dx,
                  float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dres_colsum, float* __restrict__ dx_colsum,
                  float* __restrict__ ws, int ws_rows, unsigned* __restrict__ counter) {
  constexpr int E = 768, T = 4;
  __shared__ float s_all[4 * E];
  __shared__ float xch[T][2][2][3][2];            // [triple][iteration parity][row of the pair][warp of the triple][c1 | c2]
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tri = warp / 3, part = warp % 3;
  const int col = (part * 32 + lane) * 8;
  const bool acc = dgamma != nullptr, has_res = dres != nullptr;
  float g[8], adg[8], adb[8], adr[8], adx[8];
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { adg[j] = 0.f; adb[j] = 0.f; adr[j] = 0.f; adx[j] = 0.f; }
  const float invE = 1.0f / (float)E;
  const int64_t npairs = (rows + 1) / 2, stride = (int64_t)gridDim.x * T;

  auto fetch = [&](TriRow (&b)[2], int64_t p) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t r = 2 * p + i;
      if (r < rows) {
        b[i].x = __ldg(reinterpret_cast<const uint4*>(x + r * E + col));
        b[i].dy = __ldg(reinterpret_cast<const uint4*>(dy + r * E + col));
        b[i].dr = has_res ? __ldg(reinterpret_cast<const uint4*>(dres + r * E + col)) : make_uint4(0u, 0u, 0u, 0u);
        b[i].mu = __ldg(mean + r); b[i].rs = __ldg(rstd + r);
      } else {                                     // odd row count: the missing row contributes nothing and is not stored
        b[i].x = b[i].dy = b[i].dr = make_uint4(0u, 0u, 0u, 0u);
        b[i].mu = 0.f; b[i].rs = 0.f;
      }
    }
  };
  auto unpack = [](const uint4& u, float (&v)[8]) {
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xFFFF0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xFFFF0000u);
    v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xFFFF0000u);
    v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xFFFF0000u);
  };
  auto process = [&](const TriRow (&b)[2], int64_t p, int par) {
    float xh[2][8], gd[2][8], c1[2], c2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float xv[8], dv[8];
      unpack(b[i].x, xv); unpack(b[i].dy, dv);
      const float nmr = -b[i].mu * b[i].rs;
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[i][j] = fmaf(xv[j], b[i].rs, nmr);
        gd[i][j] = dv[j] * g[j];
        a1 += gd[i][j];
        a2 = fmaf(gd[i][j], xh[i][j], a2);
        adg[j] = fmaf(dv[j], xh[i][j], adg[j]);
        adb[j] += dv[j];
      }
      c1[i] = a1; c2[i] = a2;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) { c1[i] = warp_sum(c1[i]); c2[i] = warp_sum(c2[i]); }
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 2; ++i) { xch[tri][par][i][part][0] = c1[i]; xch[tri][par][i][part][1] = c2[i]; }
    }
    asm volatile("bar.sync %0, 96;" ::"r"(1 + tri) : "memory");
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      c1[i] = ((xch[tri][par][i][0][0] + xch[tri][par][i][1][0]) + xch[tri][par][i][2][0]) * invE;
      c2[i] = ((xch[tri][par][i][0][1] + xch[tri][par][i][1][1]) + xch[tri][par][i][2][1]) * invE;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t r = 2 * p + i;
      float rv[8], o[8];
      unpack(b[i].dr, rv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float u = fmaf(-xh[i][j], c2[i], gd[i][j] - c1[i]);
        o[j] = fmaf(b[i].rs, u, rv[j]);
        if (COLS) { adr[j] += rv[j]; adx[j] += o[j]; }
      }
      if (r < rows) {
        uint4 ov;
        ov.x = pack_bf16x2(o[0], o[1]); ov.y = pack_bf16x2(o[2], o[3]); ov.z = pack_bf16x2(o[4], o[5]); ov.w = pack_bf16x2(o[6], o[7]);
        *reinterpret_cast<uint4*>(dx + r * E + col) = ov;
      }
    }
  };

  TriRow A[2], B[2];
  int64_t p = (int64_t)tri * gridDim.x + blockIdx.x;          // neighbouring CTAs work on neighbouring row pairs
  if (p < npairs) fetch(A, p);
  while (p < npairs) {                                         // unrolled by two: static buffers, no register copies
    const int64_t pn = p + stride;
    if (pn < npairs) fetch(B, pn);
    process(A, p, 0);
    const int64_t p2 = pn + stride;
    if (p2 < npairs) fetch(A, p2);
    if (pn < npairs) process(B, pn, 1);
    p = p2;
  }
  if (!acc) return;                  // dx only (dgrad-only pass): no column reductions at all
  for (int i = threadIdx.x; i < 4 * E; i += blockDim.x) s_all[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&s_all[col + j], adg[j]);
    atomicAdd(&s_all[E + col + j], adb[j]);
    if (COLS) { atomicAdd(&s_all[2 * E + col + j], adr[j]); atomicAdd(&s_all[3 * E + col + j], adx[j]); }
  }
  __syncthreads();
  if (ws != nullptr) {               // replicated accumulators, folded by the last CTA (same layout as ln_bwd_kernel)
    float* outs[4] = {dgamma, dbeta, dres_colsum, dx_colsum};
    const int offs[4] = {0, E, 2 * E, 3 * E};
    cta_replica_reduce(ws, ws_rows, counter, s_all, 4 * E, outs, offs, 4);
    return;
  }
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    atomicAdd(&dgamma[i], s_all[i]);
    atomicAdd(&dbeta[i], s_all[E + i]);
    if (dres_colsum != nullptr) atomicAdd(&dres_colsum[i], s_all[2 * E + i]);
    if (dx_colsum != nullptr) atomicAdd(&dx_colsum[i], s_all[3 * E + i]);
  }
}

// ------------------------------------------------------------------------------------------------ LN forward, E = 768 (bf16): three warps per row
// ln_fwd_kernel gives every warp a whole 768-wide row (24 values + 48 gamma / beta registers per lane, ONE row in flight per warp,
// 16 warps per SM): 24 KB of loads in flight per SM, 64 % of the HBM roofline.  With a row shared by three warps (8 columns per
// lane, one 16-byte load) a lane needs ~80 registers, 24 warps fit, and every warp triple keeps four rows in flight plus the next
// four requested: the two-pass statistics (mean, then sum of squared deviations: the same arithmetic as row_stats) meet in shared
// memory on a 96-thread named barrier, twice per four rows.
__global__ void __launch_bounds__(384, 2)
ln_fwd_tri_kernel(int64_t rows, const bf16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                  bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, float eps) {
  constexpr int E = 768, T = 4, RB = 4;
  __shared__ float xch[T][2][RB][3];              // [triple][sum | sum of squared deviations][row of the batch][warp of the triple]
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tri = warp / 3, part = warp % 3;
  const int col = (part * 32 + lane) * 8;
  float g[8], b[8];
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + col)), b1 = __ldg(reinterpret_cast<const float4*>(beta + col + 4));
    g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
  }
  const float invE = 1.0f / (float)E;
  const int64_t nbatch = (rows + RB - 1) / RB, stride = (int64_t)gridDim.x * T;
  auto fetch = [&](uint4 (&raw)[RB], int64_t q) {
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int64_t r = q * RB + i;
      raw[i] = r < rows ? __ldg(reinterpret_cast<const uint4*>(x + r * E + col)) : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  uint4 nxt[RB];
  int64_t q = (int64_t)tri * gridDim.x + blockIdx.x;
  if (q < nbatch) fetch(nxt, q);
  for (; q < nbatch; q += stride) {
    float v[RB][8], st[RB];
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const uint4 u = nxt[i];
      v[i][0] = __uint_as_float(u.x << 16); v[i][1] = __uint_as_float(u.x & 0xFFFF0000u);
      v[i][2] = __uint_as_float(u.y << 16); v[i][3] = __uint_as_float(u.y & 0xFFFF0000u);
      v[i][4] = __uint_as_float(u.z << 16); v[i][5] = __uint_as_float(u.z & 0xFFFF0000u);
      v[i][6] = __uint_as_float(u.w << 16); v[i][7] = __uint_as_float(u.w & 0xFFFF0000u);
    }
    if (q + stride < nbatch) fetch(nxt, q + stride);           // the next four rows are in flight during both passes
#pragma unroll
    for (int i = 0; i < RB; ++i) st[i] = ((v[i][0] + v[i][1]) + (v[i][2] + v[i][3])) + ((v[i][4] + v[i][5]) + (v[i][6] + v[i][7]));
#pragma unroll
    for (int i = 0; i < RB; ++i) st[i] = warp_sum(st[i]);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < RB; ++i) xch[tri][0][i][part] = st[i];
    }
    asm volatile("bar.sync %0, 96;" ::"r"(1 + tri) : "memory");
    float mean[RB];
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      mean[i] = ((xch[tri][0][i][0] + xch[tri][0][i][1]) + xch[tri][0][i][2]) * invE;
      float s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { v[i][j] -= mean[i]; s2 = fmaf(v[i][j], v[i][j], s2); }
      st[i] = s2;
    }
#pragma unroll
    for (int i = 0; i < RB; ++i) st[i] = warp_sum(st[i]);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < RB; ++i) xch[tri][1][i][part] = st[i];
    }
    asm volatile("bar.sync %0, 96;" ::"r"(1 + tri) : "memory");
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int64_t r = q * RB + i;
      const float rstd = rsqrtf(((xch[tri][1][i][0] + xch[tri][1][i][1]) + xch[tri][1][i][2]) * invE + eps);
      if (r < rows) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(v[i][j] * rstd, g[j], b[j]);
        uint4 ov;
        ov.x = pack_bf16x2(o[0], o[1]); ov.y = pack_bf16x2(o[2], o[3]); ov.z = pack_bf16x2(o[4], o[5]); ov.w = pack_bf16x2(o[6], o[7]);
        *reinterpret_cast<uint4*>(y + r * E + col) = ov;
        if (part == 0 && lane == 0) { mean_out[r] = mean[i]; rstd_out[r] = rstd; }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ LN backward, E <= 128
// Same math as ln_bwd_kernel<T, 1>, restructured for memory-level parallelism: a warp keeps RPI = 8 (bf16) / 4 (fp32) rows
// of all three inputs in flight as PACKED registers (one 8/16-byte load per lane per tensor per row) before any conversion
// or shuffle, i.e. 6 KB per warp instead of 3 KB -- the generic kernel is latency-bound at ~1.4 TB/s.
template <typename T> struct Packed4;
template <> struct Packed4<float> {
  typedef float4 raw;
  static __device__ __forceinline__ void unpack(const raw& r, float (&v)[4]) { v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w; }
};
template <> struct Packed4<bf16> {
  typedef uint2 raw;
  static __device__ __forceinline__ void unpack(const raw& r, float (&v)[4]) {
    v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xFFFF0000u);
    v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xFFFF0000u);
  }
};

template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
ln_bwd_e128_kernel(int64_t rows, int E, const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const float* __restrict__ gamma, const T* __restrict__ dres,
                   T* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dres_colsum,
                   float* __restrict__ dx_colsum, float* __restrict__ ws, int ws_rows, unsigned* __restrict__ counter,
                   float* __restrict__ part) {
  // part != nullptr: deferred reductions -- the CTA's partial column sums go to part[blockIdx.x][4E] with plain stores
  typedef typename Packed4<T>::raw raw_t;
  constexpr int RPI = sizeof(T) == 2 ? 8 : 4;
  __shared__ float s_all[4 * 128];
  pdl_trigger();
  for (int i = threadIdx.x; i < 4 * 128; i += blockDim.x) s_all[i] = 0.f;
  __syncthreads();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int c = lane * 4;
  const bool on = c < E;
  const int64_t warp = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * WARPS;
  float g[4] = {0.f, 0.f, 0.f, 0.f};
  if (on) Vec4<float>::load(gamma + c, g);
  float adg[4] = {0.f, 0.f, 0.f, 0.f}, adb[4] = {0.f, 0.f, 0.f, 0.f}, adr[4] = {0.f, 0.f, 0.f, 0.f}, adx[4] = {0.f, 0.f, 0.f, 0.f};
  const float invE = 1.0f / (float)E;
  const raw_t zero = {};
  for (int64_t r0 = warp * RPI; r0 < rows; r0 += nwarps * RPI) {
    raw_t rx[RPI], rdy[RPI], rdr[RPI];
    float mu[RPI], rs[RPI];
#pragma unroll
    for (int q = 0; q < RPI; ++q) {        // all loads of the batch first (packed: 2 registers per bf16 load)
      const int64_t r = min(r0 + q, rows - 1);
      rx[q] = on ? *reinterpret_cast<const raw_t*>(x + r * E + c) : zero;
      rdy[q] = on ? *reinterpret_cast<const raw_t*>(dy + r * E + c) : zero;
      rdr[q] = (on && dres != nullptr) ? *reinterpret_cast<const raw_t*>(dres + r * E + c) : zero;
      mu[q] = __ldg(mean + r); rs[q] = __ldg(rstd + r);
    }
#pragma unroll
    for (int q = 0; q < RPI; ++q) {
      const int64_t r = r0 + q;
      if (r >= rows) break;
      float xv[4], dv[4], rv[4];
      Packed4<T>::unpack(rx[q], xv); Packed4<T>::unpack(rdy[q], dv); Packed4<T>::unpack(rdr[q], rv);
      float c1 = 0.f, c2 = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = on ? (xv[j] - mu[q]) * rs[q] : 0.f;
        const float gd = dv[j] * g[j];
        adg[j] = fmaf(dv[j], xh, adg[j]);
        adb[j] += dv[j];
        c1 += gd;
        c2 = fmaf(gd, xh, c2);
        xv[j] = xh; dv[j] = gd;
      }
      c1 = warp_sum(c1) * invE;
      c2 = warp_sum(c2) * invE;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float t = rs[q] * (dv[j] - c1 - xv[j] * c2);
        adr[j] += rv[j];
        dv[j] = rv[j] + t;           // rv is zero when there is no skip-path gradient
        adx[j] += dv[j];
      }
      if (on) Vec4<T>::store(dx + r * E + c, dv);
    }
  }
  if (dgamma == nullptr && part == nullptr) return;      // dx only (dgrad-only pass): no column reductions at all
  if (on) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&s_all[c + j], adg[j]); atomicAdd(&s_all[E + c + j], adb[j]);
      atomicAdd(&s_all[2 * E + c + j], adr[j]); atomicAdd(&s_all[3 * E + c + j], adx[j]);
    }
  }
  __syncthreads();
  if (part != nullptr) {
    for (int i = threadIdx.x; i < 4 * E; i += blockDim.x) part[(size_t)blockIdx.x * (4 * E) + i] = s_all[i];
    return;
  }
  if (ws != nullptr) {       // same [dgamma | dbeta | colsum(dres) | colsum(dx)] x E replica layout as ln_bwd_kernel
    float* outs[4] = {dgamma, dbeta, dres_colsum, dx_colsum};
    const int offs[4] = {0, E, 2 * E, 3 * E};
    cta_replica_reduce(ws, ws_rows, counter, s_all, 4 * E, outs, offs, 4);
  } else {
    for (int i = threadIdx.x; i < E; i += blockDim.x) {
      atomicAdd(&dgamma[i], s_all[i]); atomicAdd(&dbeta[i], s_all[E + i]);
      if (dres_colsum != nullptr) atomicAdd(&dres_colsum[i], s_all[2 * E + i]);
      if (dx_colsum != nullptr) atomicAdd(&dx_colsum[i], s_all[3 * E + i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ LN, E <= 128, 8 lanes per row
// The warp-per-row kernels above are ISSUE-bound at E = 128 (ncu: 145 / 170 warp instructions per row, issue slots 66 % busy
// at 3.5 TB/s): 4 elements per lane leave the shuffles, address math and stores un-amortised.  Here a row is owned by 8 lanes
// (16 elements each: two 8-element chunks 64 columns apart, so every load instruction of the warp is fully coalesced), a
// warp works on 4 rows at once and the reductions are 3-step shuffles inside the 8-lane group: ~4x fewer instructions per row.
template <typename T> struct Row8;
template <> struct Row8<bf16> {
  typedef uint4 raw;
  static __device__ __forceinline__ raw zero() { return make_uint4(0u, 0u, 0u, 0u); }
  static __device__ __forceinline__ raw load(const bf16* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void unpack(const raw& r, float* v) {
    v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xFFFF0000u); v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xFFFF0000u);
    v[4] = __uint_as_float(r.z << 16); v[5] = __uint_as_float(r.z & 0xFFFF0000u); v[6] = __uint_as_float(r.w << 16); v[7] = __uint_as_float(r.w & 0xFFFF0000u);
  }
  static __device__ __forceinline__ void store(bf16* p, const float* v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]), c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    uint4 t;
    t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b); t.z = *reinterpret_cast<uint32_t*>(&c); t.w = *reinterpret_cast<uint32_t*>(&d);
    *reinterpret_cast<uint4*>(p) = t;
  }
};
template <> struct Row8<float> {
  struct raw { float4 a, b; };
  static __device__ __forceinline__ raw zero() { raw r; r.a = make_float4(0.f, 0.f, 0.f, 0.f); r.b = r.a; return r; }
  static __device__ __forceinline__ raw load(const float* p) { raw r; r.a = *reinterpret_cast<const float4*>(p); r.b = *reinterpret_cast<const float4*>(p + 4); return r; }
  static __device__ __forceinline__ void unpack(const raw& r, float* v) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ void load_vec8(const float* p, bool on, float* v) {
  if (on) { const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; }
  else { for (int j = 0; j < 8; ++j) v[j] = 0.f; }
}

template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
ln_fwd_x8_kernel(int64_t rows, int E, const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 T* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, float eps) {
  typedef typename Row8<T>::raw raw_t;
  constexpr int U = sizeof(T) == 2 ? 4 : 2;               // row quads in flight per warp
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, sub = lane & 7, rq = lane >> 3;
  const int c0 = sub * 8, c1 = 64 + sub * 8;
  const bool on0 = c0 < E, on1 = c1 < E;
  float g[16], b[16];
  load_vec8(gamma + c0, on0, g); load_vec8(gamma + c1, on1, g + 8);
  load_vec8(beta + c0, on0, b); load_vec8(beta + c1, on1, b + 8);
  const int64_t warp = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * WARPS;
  const float invE = 1.0f / (float)E;
  for (int64_t r0 = warp * (4 * U); r0 < rows; r0 += nwarps * (4 * U)) {
    raw_t a[U][2];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = min(r0 + u * 4 + rq, rows - 1);
      a[u][0] = on0 ? Row8<T>::load(x + r * E + c0) : Row8<T>::zero();
      a[u][1] = on1 ? Row8<T>::load(x + r * E + c1) : Row8<T>::zero();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u * 4 + rq;
      float v[16];
      Row8<T>::unpack(a[u][0], v); Row8<T>::unpack(a[u][1], v + 8);
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) s += v[j];
      const float mean = group8_sum(s) * invE;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) { v[j] = ((j < 8) ? on0 : on1) ? v[j] - mean : 0.f; q = fmaf(v[j], v[j], q); }
      const float rstd = rsqrtf(group8_sum(q) * invE + eps);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j] * rstd, g[j], b[j]);
      if (r < rows) {
        if (on0) Row8<T>::store(y + r * E + c0, v);
        if (on1) Row8<T>::store(y + r * E + c1, v + 8);
        if (sub == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
ln_bwd_x8_kernel(int64_t rows, int E, const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
                 const float* __restrict__ rstd, const float* __restrict__ gamma, const T* __restrict__ dres,
                 T* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dres_colsum,
                 float* __restrict__ dx_colsum, float* __restrict__ ws, int ws_rows, unsigned* __restrict__ counter,
                 float* __restrict__ part) {
  typedef typename Row8<T>::raw raw_t;
  constexpr int U = sizeof(T) == 2 ? 2 : 1;
  __shared__ float s_all[4 * 128];
  pdl_trigger();
  for (int i = threadIdx.x; i < 4 * 128; i += blockDim.x) s_all[i] = 0.f;
  __syncthreads();
  pdl_wait();
  const int lane = threadIdx.x & 31, sub = lane & 7, rq = lane >> 3;
  const int c0 = sub * 8, c1 = 64 + sub * 8;
  const bool on0 = c0 < E, on1 = c1 < E;
  float g[16];
  load_vec8(gamma + c0, on0, g); load_vec8(gamma + c1, on1, g + 8);
  float adg[16], adb[16], adr[16], adx[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) { adg[j] = 0.f; adb[j] = 0.f; adr[j] = 0.f; adx[j] = 0.f; }
  const int64_t warp = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * WARPS;
  const float invE = 1.0f / (float)E;
  const bool has_res = dres != nullptr;
  for (int64_t r0 = warp * (4 * U); r0 < rows; r0 += nwarps * (4 * U)) {
    raw_t ax[U][2], ady[U][2], adrs[U][2];
    float mu[U], rs[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = min(r0 + u * 4 + rq, rows - 1);
      const int64_t o0 = r * E + c0, o1 = r * E + c1;
      ax[u][0] = on0 ? Row8<T>::load(x + o0) : Row8<T>::zero();
      ax[u][1] = on1 ? Row8<T>::load(x + o1) : Row8<T>::zero();
      ady[u][0] = on0 ? Row8<T>::load(dy + o0) : Row8<T>::zero();
      ady[u][1] = on1 ? Row8<T>::load(dy + o1) : Row8<T>::zero();
      adrs[u][0] = (on0 && has_res) ? Row8<T>::load(dres + o0) : Row8<T>::zero();
      adrs[u][1] = (on1 && has_res) ? Row8<T>::load(dres + o1) : Row8<T>::zero();
      mu[u] = __ldg(mean + r); rs[u] = __ldg(rstd + r);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u * 4 + rq;
      const bool live = r < rows;                        // a clamped duplicate row must not contribute to the column sums
      float xv[16], dv[16], rv[16];
      Row8<T>::unpack(ax[u][0], xv); Row8<T>::unpack(ax[u][1], xv + 8);
      Row8<T>::unpack(ady[u][0], dv); Row8<T>::unpack(ady[u][1], dv + 8);
      Row8<T>::unpack(adrs[u][0], rv); Row8<T>::unpack(adrs[u][1], rv + 8);
      float k1 = 0.f, k2 = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const bool on = (j < 8) ? on0 : on1;
        const float xh = (on && live) ? (xv[j] - mu[u]) * rs[u] : 0.f;
        const float d = live ? dv[j] : 0.f;
        const float gd = d * g[j];
        adg[j] = fmaf(d, xh, adg[j]);
        adb[j] += d;
        k1 += gd;
        k2 = fmaf(gd, xh, k2);
        xv[j] = xh; dv[j] = gd;
      }
      k1 = group8_sum(k1) * invE;
      k2 = group8_sum(k2) * invE;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float rj = live ? rv[j] : 0.f;
        const float t = rs[u] * (dv[j] - k1 - xv[j] * k2);
        adr[j] += rj;
        dv[j] = rj + t;
        adx[j] += live ? dv[j] : 0.f;
      }
      if (live) {
        if (on0) Row8<T>::store(dx + r * E + c0, dv);
        if (on1) Row8<T>::store(dx + r * E + c1, dv + 8);
      }
    }
  }
  if (dgamma == nullptr && part == nullptr) return;      // dx only
  // fold the four row groups of the warp (lanes l, l^8, l^16, l^24 own the same columns), then one smem atomic per column
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    adg[j] += __shfl_xor_sync(0xffffffffu, adg[j], 8); adg[j] += __shfl_xor_sync(0xffffffffu, adg[j], 16);
    adb[j] += __shfl_xor_sync(0xffffffffu, adb[j], 8); adb[j] += __shfl_xor_sync(0xffffffffu, adb[j], 16);
    adr[j] += __shfl_xor_sync(0xffffffffu, adr[j], 8); adr[j] += __shfl_xor_sync(0xffffffffu, adr[j], 16);
    adx[j] += __shfl_xor_sync(0xffffffffu, adx[j], 8); adx[j] += __shfl_xor_sync(0xffffffffu, adx[j], 16);
  }
  if (rq == 0) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = ((j < 8) ? c0 : c1) + (j & 7);
      if ((j < 8) ? on0 : on1) {
        atomicAdd(&s_all[c], adg[j]); atomicAdd(&s_all[E + c], adb[j]);
        atomicAdd(&s_all[2 * E + c], adr[j]); atomicAdd(&s_all[3 * E + c], adx[j]);
      }
    }
  }
  __syncthreads();
  if (part != nullptr) {
    for (int i = threadIdx.x; i < 4 * E; i += blockDim.x) part[(size_t)blockIdx.x * (4 * E) + i] = s_all[i];
    return;
  }
  if (ws != nullptr) {
    float* outs[4] = {dgamma, dbeta, dres_colsum, dx_colsum};
    const int offs[4] = {0, E, 2 * E, 3 * E};
    cta_replica_reduce(ws, ws_rows, counter, s_all, 4 * E, outs, offs, 4);
  } else {
    for (int i = threadIdx.x; i < E; i += blockDim.x) {
      atomicAdd(&dgamma[i], s_all[i]); atomicAdd(&dbeta[i], s_all[E + i]);
      if (dres_colsum != nullptr) atomicAdd(&dres_colsum[i], s_all[2 * E + i]);
      if (dx_colsum != nullptr) atomicAdd(&dx_colsum[i], s_all[3 * E + i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ deferred column reductions
// out_k[i] += sum_p part[p][k*E + i]: CTA = 32 columns x 8 row lanes, 4 independent loads in flight per thread.
__global__ void __launch_bounds__(256)
fold_partials_kernel(const float* __restrict__ part, int n_parts, int E, float* __restrict__ o0, float* __restrict__ o1,
                     float* __restrict__ o2, float* __restrict__ o3) {
  __shared__ float red[8][33];
  pdl_trigger();
  pdl_wait();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, ncols = 4 * E;
  const int col = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (col < ncols) {
    int p = ty;
    for (; p + 24 < n_parts; p += 32) {
      const float a = __ldg(part + (size_t)p * ncols + col), b = __ldg(part + (size_t)(p + 8) * ncols + col),
                  c = __ldg(part + (size_t)(p + 16) * ncols + col), d = __ldg(part + (size_t)(p + 24) * ncols + col);
      acc += (a + b) + (c + d);
    }
    for (; p < n_parts; p += 8) acc += __ldg(part + (size_t)p * ncols + col);
  }
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && col < ncols) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][tx];
    const int k = col / E, i = col - k * E;
    float* o = k == 0 ? o0 : (k == 1 ? o1 : (k == 2 ? o2 : o3));
    if (o != nullptr) atomicAdd(o + i, s);
  }
}

// ------------------------------------------------------------------------------------------------ SLN forward
template <typename T, int NV>
__global__ void __launch_bounds__(WARPS * 32)
sln_fwd_kernel(int64_t rows, int64_t h_rows, int F, const T* __restrict__ h, const T* __restrict__ w,
               const float* __restrict__ ln_g, const float* __restrict__ ln_b, const float* __restrict__ gamma_s,
               const float* __restrict__ beta_s, T* __restrict__ y, float* __restrict__ mean_out,
               float* __restrict__ rstd_out, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * WARPS;
  float g[NV][4], b[NV][4];
  load_vec(ln_g, F, lane, g);
  load_vec(ln_b, F, lane, b);
  const float gs = *gamma_s, bs = *beta_s;
  for (int64_t r = warp; r < rows; r += nwarps) {
    float hv[NV][4], wv[NV][4];
    load_row<T, NV>(h + (r % h_rows) * F, F, lane, hv);
    load_row<T, NV>(w + r * F, F, lane, wv);
    float mean, rstd;
    row_stats(hv, F, lane, mean, rstd, eps);
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float n = fmaf((hv[i][j] - mean) * rstd, g[i][j], b[i][j]);
        hv[i][j] = gs * wv[i][j] * n + bs * wv[i][j];      // same association as the reference expression
      }
    store_row<T, NV>(y + r * F, F, lane, hv);
    if (lane == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
  }
}

// ------------------------------------------------------------------------------------------------ SLN backward
template <typename T, int NV>
__global__ void __launch_bounds__(WARPS * 32)
sln_bwd_kernel(int64_t rows, int64_t h_rows, int F, const T* __restrict__ dy, const T* __restrict__ h,
               const T* __restrict__ w, const float* __restrict__ mean, const float* __restrict__ rstd,
               const float* __restrict__ ln_g, const float* __restrict__ ln_b, const float* __restrict__ gamma_s,
               const float* __restrict__ beta_s, const T* __restrict__ dh_res, const T* __restrict__ dw_res,
               void* __restrict__ dh_out, T* __restrict__ dw, float* __restrict__ dgamma_s,
               float* __restrict__ dbeta_s, float* __restrict__ dln_g, float* __restrict__ dln_b) {
  __shared__ float s_dg[MAXE], s_db[MAXE];
  __shared__ float s_scal[2];
  for (int i = threadIdx.x; i < F; i += blockDim.x) { s_dg[i] = 0.f; s_db[i] = 0.f; }
  if (threadIdx.x < 2) s_scal[threadIdx.x] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * WARPS;
  const bool bcast = h_rows < rows;
  float g[NV][4], b[NV][4], adg[NV][4], adb[NV][4];
  load_vec(ln_g, F, lane, g);
  load_vec(ln_b, F, lane, b);
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { adg[i][j] = 0.f; adb[i][j] = 0.f; }
  const float gs = *gamma_s, bs = *beta_s;
  const float invF = 1.0f / (float)F;
  float a_gs = 0.f, a_bs = 0.f;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const int64_t hr = r % h_rows;
    float hv[NV][4], wv[NV][4], dv[NV][4];
    load_row<T, NV>(h + hr * F, F, lane, hv);
    load_row<T, NV>(w + r * F, F, lane, wv);
    load_row<T, NV>(dy + r * F, F, lane, dv);
    const float mu = mean[r], rs = rstd[r];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = (hv[i][j] - mu) * rs;
        const float n = fmaf(xh, g[i][j], b[i][j]);      // LN output
        const float da = dv[i][j] * wv[i][j];            // d/d(gamma_s*n + beta_s)
        a_gs = fmaf(da, n, a_gs);
        a_bs += da;
        wv[i][j] = dv[i][j] * (gs * n + bs);             // dw
        const float dn = da * gs;                        // grad of LN output
        adg[i][j] = fmaf(dn, xh, adg[i][j]);
        adb[i][j] += dn;
        const float gd = dn * g[i][j];
        c1 += gd;
        c2 = fmaf(gd, xh, c2);
        hv[i][j] = xh; dv[i][j] = gd;
      }
    c1 = warp_sum(c1) * invF;
    c2 = warp_sum(c2) * invF;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dv[i][j] = rs * (dv[i][j] - c1 - hv[i][j] * c2);   // dh
    if (dw_res != nullptr) {
      float rv[NV][4];
      load_row<T, NV>(dw_res + r * F, F, lane, rv);
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) wv[i][j] += rv[i][j];
    }
    store_row<T, NV>(dw + r * F, F, lane, wv);
    if (bcast) {   // h is (S,F) shared by the whole batch: reduce over b with fp32 atomics
      float* dh32 = static_cast<float*>(dh_out) + hr * F;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < F) {
#pragma unroll
          for (int j = 0; j < 4; ++j) atomicAdd(&dh32[c + j], dv[i][j]);
        }
      }
    } else {
      if (dh_res != nullptr) {
        float rv[NV][4];
        load_row<T, NV>(dh_res + r * F, F, lane, rv);
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dv[i][j] += rv[i][j];
      }
      store_row<T, NV>(static_cast<T*>(dh_out) + r * F, F, lane, dv);
    }
  }
  flush_cols(s_dg, dln_g, adg, F, lane);
  flush_cols(s_db, dln_b, adb, F, lane);
  a_gs = warp_sum(a_gs);
  a_bs = warp_sum(a_bs);
  if (lane == 0) { atomicAdd(&s_scal[0], a_gs); atomicAdd(&s_scal[1], a_bs); }
  __syncthreads();
  for (int i = threadIdx.x; i < F; i += blockDim.x) { atomicAdd(&dln_g[i], s_dg[i]); atomicAdd(&dln_b[i], s_db[i]); }
  if (threadIdx.x == 0) { atomicAdd(dgamma_s, s_scal[0]); atomicAdd(dbeta_s, s_scal[1]); }
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  const int v = e ? atoi(e) : dflt;
  return v >= 1 ? v : dflt;
}
bool x8_enabled() {
  static const bool on = [] { const char* e = getenv("VG_LN_X8"); return !(e && e[0] == '0'); }();   // read once, thread-safe
  return on;
}
int grid_for_rows(int64_t rows, int ctas_per_sm) {
  const int64_t need = (rows + WARPS - 1) / WARPS;
  const int64_t cap = (int64_t)num_sms() * ctas_per_sm;
  return (int)max((int64_t)1, min(need, cap));
}

}  // namespace
}  // namespace vg

using namespace vg;

// pick the register tile: NV float4 per lane covers E <= NV*128
#define VG_NV_DISPATCH(E, CALL)            \
  do {                                     \
    if ((E) <= 128) { constexpr int NV = 1; CALL; }       \
    else if ((E) <= 256) { constexpr int NV = 2; CALL; }  \
    else if ((E) <= 512) { constexpr int NV = 4; CALL; }  \
    else { constexpr int NV = 8; CALL; }                  \
  } while (0)

#define VG_NORM_CHECK(E)                                                                                        \
  VG_REQUIRE((E) > 0 && (E) % 4 == 0 && (E) <= MAXE, VG_ERR_SHAPE, "norm: feature dim %d must be a multiple of 4 and <= %d", (E), MAXE)

extern "C" int vg_layernorm_fwd(int dtype, int64_t rows, int E, const void* x, const float* gamma, const float* beta,
                                void* y, float* mean, float* rstd, float eps, void* stream) {
  VG_NORM_CHECK(E);
  if (rows == 0) return VG_OK;
  if (E <= 128 && E % 8 == 0 && x8_enabled()) {
    const int rpi = dtype == VG_F32 ? 8 : 16;
    const int g8 = grid_for_rows((rows + rpi - 1) / rpi, env_int("VG_LN_FWD_CTAS_PER_SM", 4));
    if (dtype == VG_F32)
      launch_pdl(ln_fwd_x8_kernel<float>, dim3(g8), dim3(WARPS * 32), 0, as_stream(stream), rows, E, (const float*)x, gamma, beta, (float*)y, mean, rstd, eps);
    else
      launch_pdl(ln_fwd_x8_kernel<bf16>, dim3(g8), dim3(WARPS * 32), 0, as_stream(stream), rows, E, (const bf16*)x, gamma, beta, (bf16*)y, mean, rstd, eps);
    return check_launch("layernorm_fwd");
  }
  static const bool tri_off = [] { const char* e = getenv("VG_LN_TRI"); return e && e[0] == '0'; }();
  if (dtype == VG_BF16 && E == 768 && !tri_off && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    // three warps per row, 24 warps per SM (see ln_fwd_tri_kernel); VG_LN_TRI=0 keeps the one-warp-per-row kernel below
    const int gridt = (int)max((int64_t)1, min((rows + 15) / 16, (int64_t)2 * num_sms()));
    launch_pdl(ln_fwd_tri_kernel, dim3(gridt), dim3(384), 0, as_stream(stream), rows, (const bf16*)x, gamma, beta, (bf16*)y, mean, rstd, eps);
    return check_launch("layernorm_fwd");
  }
  const int grid = grid_for_rows(rows, 8);
  if (dtype == VG_F32)
    VG_NV_DISPATCH(E, (launch_pdl(ln_fwd_kernel<float, NV>, dim3(grid), dim3(WARPS * 32), 0, as_stream(stream), rows, E, (const float*)x, gamma, beta, (float*)y, mean, rstd, eps)));
  else
    VG_NV_DISPATCH(E, (launch_pdl(ln_fwd_kernel<bf16, NV>, dim3(grid), dim3(WARPS * 32), 0, as_stream(stream), rows, E, (const bf16*)x, gamma, beta, (bf16*)y, mean, rstd, eps)));
  return check_launch("layernorm_fwd");
}

extern "C" int vg_layernorm_bwd(int dtype, int64_t rows, int E, const void* dy, const void* x, const float* mean,
                                const float* rstd, const float* gamma, const void* dres, void* dx, float* dgamma,
                                float* dbeta, float* dres_colsum, float* dx_colsum, float* workspace, int ws_rows,
                                unsigned* counter, void* stream) {
  VG_NORM_CHECK(E);
  VG_REQUIRE(!(dres_colsum && !dres), VG_ERR_ARG, "layernorm_bwd: dres_colsum without dres");
  VG_REQUIRE((dgamma == nullptr) == (dbeta == nullptr) && !(dgamma == nullptr && (dres_colsum || dx_colsum)), VG_ERR_ARG,
             "layernorm_bwd: dgamma/dbeta must be given together (both NULL = dx only, no column sums)");
  VG_REQUIRE(!(workspace && (!counter || ws_rows < 1)), VG_ERR_ARG, "layernorm_bwd: workspace needs a counter and ws_rows >= 1");
  if (rows == 0) return VG_OK;
  static const int per_sm = [] { const char* e = getenv("VG_LN_BWD_CTAS_PER_SM"); const int v = e ? atoi(e) : 2; return v < 1 ? 2 : v; }();
  // opt-in (VG_LN_BWD_X8=1): fewer instructions per row but 193 registers; measured slower (19 vs 17 us at C2)
  static const int bwd_x8 = [] { const char* e = getenv("VG_LN_BWD_X8"); return (e && e[0] == '1') ? 1 : 0; }();
  if (E <= 128 && E % 8 == 0 && bwd_x8 == 1) {      // 8 lanes per row, 8 (bf16) / 4 (fp32) rows per warp in flight
    const int rpi = dtype == VG_F32 ? 4 : 8;
    const int g8 = grid_for_rows((rows + rpi - 1) / rpi, env_int("VG_LN_BWD_X8_CTAS_PER_SM", 2));
    if (dtype == VG_F32)
      launch_pdl(ln_bwd_x8_kernel<float>, dim3(g8), dim3(WARPS * 32), 0, as_stream(stream), rows, E, (const float*)dy, (const float*)x, mean,
                 rstd, gamma, (const float*)dres, (float*)dx, dgamma, dbeta, dres_colsum, dx_colsum, workspace, ws_rows, counter, (float*)nullptr);
    else
      launch_pdl(ln_bwd_x8_kernel<bf16>, dim3(g8), dim3(WARPS * 32), 0, as_stream(stream), rows, E, (const bf16*)dy, (const bf16*)x, mean,
                 rstd, gamma, (const bf16*)dres, (bf16*)dx, dgamma, dbeta, dres_colsum, dx_colsum, workspace, ws_rows, counter, (float*)nullptr);
    return check_launch("layernorm_bwd");
  }
  if (E <= 128) {      // specialised kernel: 8 (bf16) / 4 (fp32) rows per warp in flight
    const int rpi = dtype == VG_F32 ? 4 : 8;
    const int grid = grid_for_rows((rows + rpi - 1) / rpi, per_sm);
    if (dtype == VG_F32)
      launch_pdl(ln_bwd_e128_kernel<float>, dim3(grid), dim3(WARPS * 32), 0, as_stream(stream), rows, E, (const float*)dy, (const float*)x, mean,
                 rstd, gamma, (const float*)dres, (float*)dx, dgamma, dbeta, dres_colsum, dx_colsum, workspace, ws_rows, counter, (float*)nullptr);
    else
      launch_pdl(ln_bwd_e128_kernel<bf16>, dim3(grid), dim3(WARPS * 32), 0, as_stream(stream), rows, E, (const bf16*)dy, (const bf16*)x, mean,
                 rstd, gamma, (const bf16*)dres, (bf16*)dx, dgamma, dbeta, dres_colsum, dx_colsum, workspace, ws_rows, counter, (float*)nullptr);
    return check_launch("layernorm_bwd");
  }
  static const bool tri_off = [] { const char* e = getenv("VG_LN_TRI"); return e && e[0] == '0'; }();
  if (dtype == VG_BF16 && E == 768 && !tri_off && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) |
                                                      reinterpret_cast<uintptr_t>(dres)) & 15) == 0) {
    // three warps per row, 12 warps per SM (see ln_bwd_tri_kernel); VG_LN_TRI=0 keeps the one-warp-per-row kernel below
    const int gridt = (int)max((int64_t)1, min((rows + 7) / 8, (int64_t)num_sms()));
    if (dres_colsum != nullptr || dx_colsum != nullptr)
      launch_pdl(ln_bwd_tri_kernel<true>, dim3(gridt), dim3(384), 0, as_stream(stream), rows, (const bf16*)dy, (const bf16*)x, mean, rstd, gamma,
                 (const bf16*)dres, (bf16*)dx, dgamma, dbeta, dres_colsum, dx_colsum, workspace, ws_rows, counter);
    else
      launch_pdl(ln_bwd_tri_kernel<false>, dim3(gridt), dim3(384), 0, as_stream(stream), rows, (const bf16*)dy, (const bf16*)x, mean, rstd, gamma,
                 (const bf16*)dres, (bf16*)dx, dgamma, dbeta, dres_colsum, dx_colsum, workspace, ws_rows, counter);
    return check_launch("layernorm_bwd");
  }
  if (dtype == VG_BF16 && E > 256 && E <= 768 && E % 4 == 0 && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) |
                                                   reinterpret_cast<uintptr_t>(dres)) & 7) == 0) {
    // wide rows: packed operands, two rows per warp in flight, one CTA per SM
    const int gridw = grid_for_rows(rows, 1);
#define VG_LN_WIDE_(NV_, COLS_) launch_pdl(ln_bwd_wide_kernel<NV_, COLS_>, dim3(gridw), dim3(WARPS * 32), 0, as_stream(stream), rows, E, (const bf16*)dy, (const bf16*)x, mean, \
                                   rstd, gamma, (const bf16*)dres, (bf16*)dx, dgamma, dbeta, dres_colsum, dx_colsum, workspace, ws_rows, counter)
#define VG_LN_WIDE(NV_) do { if (dres_colsum != nullptr || dx_colsum != nullptr) VG_LN_WIDE_(NV_, true); else VG_LN_WIDE_(NV_, false); } while (0)
    if (E <= 384) VG_LN_WIDE(3); else if (E <= 512) VG_LN_WIDE(4); else VG_LN_WIDE(6);
#undef VG_LN_WIDE_
#undef VG_LN_WIDE
    return check_launch("layernorm_bwd");
  }
  int grid = grid_for_rows((rows + 3) / 4, per_sm);   // CTAs/SM x 8 warps x 4 rows x 3 tensors of 16 B loads in flight
  if (dtype == VG_F32)
    VG_NV_DISPATCH(E, (launch_pdl(ln_bwd_kernel<float, NV>, dim3(grid), dim3(WARPS * 32), 0, as_stream(stream), rows, E, (const float*)dy, (const float*)x, mean, rstd, gamma,
                                                                      (const float*)dres, (float*)dx, dgamma, dbeta, dres_colsum, dx_colsum, workspace, ws_rows, counter)));
  else
    VG_NV_DISPATCH(E, (launch_pdl(ln_bwd_kernel<bf16, NV>, dim3(grid), dim3(WARPS * 32), 0, as_stream(stream), rows, E, (const bf16*)dy, (const bf16*)x, mean, rstd, gamma,
                                                                     (const bf16*)dres, (bf16*)dx, dgamma, dbeta, dres_colsum, dx_colsum, workspace, ws_rows, counter)));
  return check_launch("layernorm_bwd");
}

extern "C" int vg_sln_fwd(int dtype, int64_t rows, int64_t h_rows, int F, const void* h, const void* w,
                          const float* ln_g, const float* ln_b, const float* gamma_s, const float* beta_s, void* y,
                          float* mean, float* rstd, float eps, void* stream) {
  VG_NORM_CHECK(F);
  VG_REQUIRE(h_rows > 0 && rows % h_rows == 0, VG_ERR_SHAPE, "sln: rows %lld not a multiple of h_rows %lld", (long long)rows, (long long)h_rows);
  if (rows == 0) return VG_OK;
  const int grid = grid_for_rows(rows, 8);
  if (dtype == VG_F32)
    VG_NV_DISPATCH(F, (sln_fwd_kernel<float, NV><<<grid, WARPS * 32, 0, as_stream(stream)>>>(rows, h_rows, F, (const float*)h, (const float*)w, ln_g, ln_b,
                                                                       gamma_s, beta_s, (float*)y, mean, rstd, eps)));
  else
    VG_NV_DISPATCH(F, (sln_fwd_kernel<bf16, NV><<<grid, WARPS * 32, 0, as_stream(stream)>>>(rows, h_rows, F, (const bf16*)h, (const bf16*)w, ln_g, ln_b,
                                                                      gamma_s, beta_s, (bf16*)y, mean, rstd, eps)));
  return check_launch("sln_fwd");
}

extern "C" int vg_sln_bwd(int dtype, int64_t rows, int64_t h_rows, int F, const void* dy, const void* h, const void* w,
                          const float* mean, const float* rstd, const float* ln_g, const float* ln_b,
                          const float* gamma_s, const float* beta_s, const void* dh_res, const void* dw_res, void* dh,
                          void* dw, float* dgamma_s, float* dbeta_s, float* dln_g, float* dln_b, void* stream) {
  VG_NORM_CHECK(F);
  VG_REQUIRE(h_rows > 0 && rows % h_rows == 0, VG_ERR_SHAPE, "sln: rows %lld not a multiple of h_rows %lld", (long long)rows, (long long)h_rows);
  VG_REQUIRE(!(h_rows < rows && dh_res), VG_ERR_ARG, "sln_bwd: dh_res unsupported with broadcast h");
  if (rows == 0) return VG_OK;
  const int grid = grid_for_rows(rows, 2);
  if (dtype == VG_F32)
    VG_NV_DISPATCH(F, (sln_bwd_kernel<float, NV><<<grid, WARPS * 32, 0, as_stream(stream)>>>(rows, h_rows, F, (const float*)dy, (const float*)h, (const float*)w,
        mean, rstd, ln_g, ln_b, gamma_s, beta_s, (const float*)dh_res, (const float*)dw_res, dh, (float*)dw, dgamma_s, dbeta_s, dln_g, dln_b)));
  else
    VG_NV_DISPATCH(F, (sln_bwd_kernel<bf16, NV><<<grid, WARPS * 32, 0, as_stream(stream)>>>(rows, h_rows, F, (const bf16*)dy, (const bf16*)h, (const bf16*)w,
        mean, rstd, ln_g, ln_b, gamma_s, beta_s, (const bf16*)dh_res, (const bf16*)dw_res, dh, (bf16*)dw, dgamma_s, dbeta_s, dln_g, dln_b)));
  return check_launch("sln_bwd");
}

extern "C" int vg_layernorm_bwd_partials(int dtype, int64_t rows, int E, const void* dy, const void* x, const float* mean,
                                         const float* rstd, const float* gamma, const void* dres, void* dx, float* partials,
                                         int max_parts, void* stream) {
  VG_NORM_CHECK(E);
  VG_REQUIRE(E <= 128, VG_ERR_UNSUPPORTED, "layernorm_bwd_partials: E <= 128 only (got %d)", E);
  VG_REQUIRE(partials != nullptr && max_parts >= 1, VG_ERR_ARG, "layernorm_bwd_partials: no partial buffer");
  VG_REQUIRE(rows > 0, VG_ERR_SHAPE, "layernorm_bwd_partials: empty input");
  const int rpi = dtype == VG_F32 ? 4 : 8;
  static const int px8 = [] { const char* e = getenv("VG_LN_BWD_X8"); return (e && e[0] == '1') ? 1 : 0; }();
  if (px8 == 1 && E % 8 == 0) {          // experiment: 8-lanes-per-row variant (half the instructions, 193 registers -> 1 CTA/SM)
    const int g8 = min(grid_for_rows((rows + rpi - 1) / rpi, env_int("VG_LN_BWD_X8_CTAS_PER_SM", 1)), max_parts);
    if (dtype == VG_F32)
      launch_pdl(ln_bwd_x8_kernel<float>, dim3(g8), dim3(WARPS * 32), 0, as_stream(stream), rows, E, (const float*)dy, (const float*)x, mean,
                 rstd, gamma, (const float*)dres, (float*)dx, (float*)nullptr, (float*)nullptr, (float*)nullptr, (float*)nullptr, (float*)nullptr, 0,
                 (unsigned*)nullptr, partials);
    else
      launch_pdl(ln_bwd_x8_kernel<bf16>, dim3(g8), dim3(WARPS * 32), 0, as_stream(stream), rows, E, (const bf16*)dy, (const bf16*)x, mean,
                 rstd, gamma, (const bf16*)dres, (bf16*)dx, (float*)nullptr, (float*)nullptr, (float*)nullptr, (float*)nullptr, (float*)nullptr, 0,
                 (unsigned*)nullptr, partials);
    const int rc8 = check_launch("layernorm_bwd_partials");
    return rc8 ? rc8 : g8;
  }
  const int grid = min(grid_for_rows((rows + rpi - 1) / rpi, 2), max_parts);
  if (dtype == VG_F32)
    launch_pdl(ln_bwd_e128_kernel<float>, dim3(grid), dim3(WARPS * 32), 0, as_stream(stream), rows, E, (const float*)dy, (const float*)x, mean,
               rstd, gamma, (const float*)dres, (float*)dx, (float*)nullptr, (float*)nullptr, (float*)nullptr, (float*)nullptr, (float*)nullptr, 0,
               (unsigned*)nullptr, partials);
  else
    launch_pdl(ln_bwd_e128_kernel<bf16>, dim3(grid), dim3(WARPS * 32), 0, as_stream(stream), rows, E, (const bf16*)dy, (const bf16*)x, mean,
               rstd, gamma, (const bf16*)dres, (bf16*)dx, (float*)nullptr, (float*)nullptr, (float*)nullptr, (float*)nullptr, (float*)nullptr, 0,
               (unsigned*)nullptr, partials);
  const int rc = check_launch("layernorm_bwd_partials");
  return rc ? rc : grid;
}

extern "C" int vg_fold_partials(const float* partials, int n_parts, int E, float* out0, float* out1, float* out2, float* out3,
                                void* stream) {
  VG_REQUIRE(partials != nullptr && n_parts >= 1 && E >= 1, VG_ERR_ARG, "fold_partials: bad arguments");
  launch_pdl(fold_partials_kernel, dim3((4 * E + 31) / 32), dim3(256), 0, as_stream(stream), partials, n_parts, E, out0, out1, out2, out3);
  return check_launch("fold_partials");
}
