// elementwise.cu -- HBM-bound helpers: dtype cast (+ spectral rescale), column sums (bias gradients),
// in-place add, row broadcast, fused Adam/AdamW.  All grid-stride, vectorised where alignment allows.
#include "common.cuh"

namespace vg {
namespace {

int grid1d(int64_t total, int block, int per_sm = 16) {
  const int64_t need = (total + block - 1) / block;
  return (int)max((int64_t)1, min(need, (int64_t)num_sms() * per_sm));
}

template <typename TS, typename TD>
__global__ void cast_scale_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int64_t n,
                                  const float* __restrict__ num, const float* __restrict__ den) {
  pdl_trigger();
  pdl_wait();
  float sc = 1.f;
  if (num) sc = *num;
  if (den) sc = sc / *den;
  const bool scaled = num || den;
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  const int64_t n4 = aligned ? n / 4 : 0;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = gtid; i < n4; i += gsz) {
    float v[4];
    Vec4<TS>::load(src + i * 4, v);
    if (scaled) { v[0] *= sc; v[1] *= sc; v[2] *= sc; v[3] *= sc; }
    Vec4<TD>::store(dst + i * 4, v);
  }
  for (int64_t i = n4 * 4 + gtid; i < n; i += gsz) {
    float v = to_f<TS>(src[i]);
    if (scaled) v *= sc;
    dst[i] = from_f<TD>(v);
  }
}

// n fp32 [rows, cols] tensors (device pointer array) -> one [n * rows_pad, cols_pad] tensor of TD, each block scaled by
// num[t] / den[t] and zero-padded to rows_pad x cols_pad: the grouped q/k/v weight of the v1 heads (src/v1/attention.py:46-48,
// 60-64 incl. the spectral rescale) with head widths padded to the tensor-core granularity, in ONE launch.
template <typename TD>
__global__ void pack_pad_kernel(const float* const* __restrict__ srcs, int n, int rows, int cols, int rows_pad, int cols_pad,
                                const float* __restrict__ num, const float* __restrict__ den, TD* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int64_t per = (int64_t)rows_pad * cols_pad, total = per * n;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = gtid; i < total; i += gsz) {
    const int t = (int)(i / per);
    const int64_t rem = i - (int64_t)t * per;
    const int r = (int)(rem / cols_pad), c = (int)(rem - (int64_t)r * cols_pad);
    float v = 0.f;
    if (r < rows && c < cols) {
      v = __ldg(srcs[t] + (int64_t)r * cols + c);
      if (num) v *= __ldg(num + t);
      if (den) v /= __ldg(den + t);
    }
    out[i] = from_f<TD>(v);
  }
}

// out[n] += sum_m x[m,n].  Vector version: thread (tx, ty) owns the 16-byte column group tx of the CTA's 32-group slab
// and the rows m0+ty, m0+ty+8, ...; a warp reads 512 contiguous bytes per row; 4 rows in flight per thread.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const T* __restrict__ x, int64_t M, int N, int64_t ld, float* __restrict__ out, int64_t rows_per_cta,
                  float* __restrict__ ws, int ws_rows, unsigned* __restrict__ counter) {
  constexpr int V = 16 / sizeof(T);
  __shared__ float red[8][32][V + 1];
  pdl_trigger();
  pdl_wait();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const bool wsmode = ws != nullptr;                   // workspace mode: 1-D grid over row slices, one 32-group column slab
  const int cg = (wsmode ? 0 : blockIdx.x * 32) + tx;  // column group
  const int n0 = cg * V;
  const int64_t m0 = (int64_t)(wsmode ? blockIdx.x : blockIdx.y) * rows_per_cta, m1 = min(M, m0 + rows_per_cta);
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.f;
  if (n0 < N) {
    int64_t m = m0 + ty;
    for (; m + 24 < m1; m += 32) {
      uint4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = *reinterpret_cast<const uint4*>(x + (m + 8 * u) * ld + n0);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const T* e = reinterpret_cast<const T*>(&t[u]);
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] += to_f<T>(e[j]);
      }
    }
    for (; m < m1; m += 8) {
      const uint4 t = *reinterpret_cast<const uint4*>(x + m * ld + n0);
      const T* e = reinterpret_cast<const T*>(&t);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] += to_f<T>(e[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) red[ty][tx][j] = acc[j];
  __syncthreads();
  if (ws != nullptr) {      // 1-D grid over row slices, all N columns per CTA (N <= 32*V): partials + last-CTA reduction
    __shared__ float s_cols[32 * V];
    if (ty == 0) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][tx][j];
        s_cols[tx * V + j] = s;
      }
    }
    __syncthreads();
    float* outs[1] = {out};
    const int offs[1] = {0};
    cta_replica_reduce(ws, ws_rows, counter, s_cols, N, outs, offs, 1);
    return;
  }
  if (ty == 0 && n0 < N) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += red[i][tx][j];
      atomicAdd(&out[n0 + j], s);
    }
  }
}

// scalar fallback (unaligned / odd N): CTA = (32-col slab) x (row slice); threads (32 cols x 8 row lanes)
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, int64_t M, int N, int64_t ld, float* __restrict__ out,
                              int64_t rows_per_cta) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const int64_t m0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t m1 = min(M, m0 + rows_per_cta);
  float acc = 0.f;
  if (n < N)
    for (int64_t m = m0 + ty; m < m1; m += 8) acc += to_f<T>(x[m * ld + n]);
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][tx];
    atomicAdd(&out[n], s);
  }
}

template <typename T>
__global__ void add_inplace_kernel(T* __restrict__ x, const T* __restrict__ y, int64_t n) {
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n / 4;
  for (int64_t i = gtid; i < n4; i += gsz) {
    float a[4], b[4];
    Vec4<T>::load(x + i * 4, a);
    Vec4<T>::load(y + i * 4, b);
    a[0] += b[0]; a[1] += b[1]; a[2] += b[2]; a[3] += b[3];
    Vec4<T>::store(x + i * 4, a);
  }
  for (int64_t i = n4 * 4 + gtid; i < n; i += gsz) x[i] = from_f<T>(to_f<T>(x[i]) + to_f<T>(y[i]));
}

template <typename T>
__global__ void broadcast_rows_kernel(const T* __restrict__ src, int64_t src_elems, T* __restrict__ dst, int64_t reps) {
  const int64_t total = src_elems * reps;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = gtid; i < total; i += gsz) dst[i] = src[i % src_elems];
}

template <typename T>
__global__ void copy_rows_kernel(int64_t rows, int cols, const T* __restrict__ src, int64_t lds, T* __restrict__ dst, int64_t ldd) {
  const int64_t total = rows * cols;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = gtid; i < total; i += gsz) {
    const int64_t r = i / cols; const int c = (int)(i % cols);
    dst[r * ldd + c] = src[r * lds + c];
  }
}

template <typename T, int ACT>
__device__ __forceinline__ void act_bwd_loop(int64_t n, const T* dy, const T* aux, float prm, T* out, int64_t gtid, int64_t gsz) {
  for (int64_t i = gtid; i < n; i += gsz) out[i] = from_f<T>(act_t<ACT, false>(to_f<T>(dy[i]), to_f<T>(aux[i]), prm));
}
template <typename T>
__global__ void act_bwd_kernel(int64_t n, const T* __restrict__ dy, const T* __restrict__ aux, int act, float prm, T* __restrict__ out) {
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
  VG_ACT_SWITCH(act, (act_bwd_loop<T, ACT>(n, dy, aux, prm, out, gtid, gsz)))
}

// torch.optim.Adam / AdamW single-tensor semantics (no amsgrad, no maximize):
//   AdamW: p *= 1 - lr*wd;                  Adam: g += wd*p
//   m = b1*m + (1-b1)*g; v = b2*v + (1-b2)*g*g
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float wd,
                            int decoupled, float grad_scale, const int* __restrict__ step_count, bf16* __restrict__ shadow) {
  const int t = *step_count + 1;     // the counter is bumped by a separate 1-thread kernel after all tensors
  const float bc1 = 1.f - powf(b1, (float)t);
  const float bc2s = sqrtf(1.f - powf(b2, (float)t));
  const float step_size = lr / bc1;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = gtid; i < n; i += gsz) {
    float pi = p[i], gi = g[i] * grad_scale;
    if (decoupled) pi *= 1.f - lr * wd; else gi += wd * pi;
    const float mi = b1 * m[i] + (1.f - b1) * gi;          // lerp form: m + (g-m)*(1-b1) in torch; same to 1 ulp
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    const float denom = sqrtf(vi) / bc2s + eps;
    pi -= step_size * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (shadow != nullptr) shadow[i] = __float2bfloat16_rn(pi);   // bf16 operand copy for the tensor-core GEMMs, kept in sync here
  }
}
__global__ void bump_kernel(int* c) { *c += 1; }


// ------------------------------------------------------------------------------------------------ loss head / sampling tail
// Fused CrossEntropyLoss (mean reduction, class-index targets: src/v2/training.py:159) over `rows/rows_per_group` groups of
// rows (e.g. the real half and the fake half of the merged discriminator pass): losses[g] = mean_{r in g} (lse(z_r) - z_r[t_r]);
// dlogits = d(sum_g losses[g]) / dz.  Replaces log_softmax + nll_loss forward/backward + the reductions (~10 launches) by one.
// The head is [B, 10]: one CTA is plenty.
__global__ void __launch_bounds__(1024)
softmax_ce_kernel(const float* __restrict__ logits, const int64_t* __restrict__ targets, int rows, int C, int rows_per_group,
                  float* __restrict__ losses, float* __restrict__ dlogits) {
  __shared__ float s_loss[64];
  const int n_groups = rows / rows_per_group;
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_loss[i] = 0.f;
  __syncthreads();
  const float inv = 1.0f / (float)rows_per_group;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const float* z = logits + (size_t)r * C;
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, z[c]);
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum += expf(z[c] - m);
    const int t = (int)targets[r];
    const float lse = m + logf(sum);
    const float rs = 1.0f / sum;
    for (int c = 0; c < C; ++c) dlogits[(size_t)r * C + c] = (expf(z[c] - m) * rs - (c == t ? 1.0f : 0.f)) * inv;
    atomicAdd(&s_loss[r / rows_per_group], (lse - z[t]) * inv);
  }
  __syncthreads();
  for (int g = threadIdx.x; g < n_groups; g += blockDim.x) losses[g] = s_loss[g];
}

// Fused nn.BCELoss (mean reduction, float targets: src/v1/gan.py:16-20, 222-252) over rows/rows_per_group groups of probabilities
// (the discriminator's sigmoid output, one value per row): losses[g] = mean_{r in g} -(t log p + (1 - t) log(1 - p)) with both
// logs clamped at -100 (torch's BCELoss), dprob = d(sum_g losses[g]) / dp = (p - t) / max(p (1 - p), 1e-12) / rows_per_group.
__global__ void __launch_bounds__(1024)
bce_kernel(const float* __restrict__ prob, const float* __restrict__ target, int rows, int rows_per_group,
           float* __restrict__ losses, float* __restrict__ dprob) {
  __shared__ float s_loss[64];
  const int n_groups = rows / rows_per_group;
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_loss[i] = 0.f;
  __syncthreads();
  const float inv = 1.0f / (float)rows_per_group;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const float p = prob[r], t = target[r];
    const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(log1pf(-p), -100.f);
    dprob[r] = (p - t) / fmaxf((1.0f - p) * p, 1e-12f) * inv;
    atomicAdd(&s_loss[r / rows_per_group], -(t * lp + (1.0f - t) * lq) * inv);
  }
  __syncthreads();
  for (int g = threadIdx.x; g < n_groups; g += blockDim.x) losses[g] = s_loss[g];
}

// utils.convert_to_uint8 (src/v2/utils.py:194-196): (images * 127.5 + 127.5).clamp(0, 255).to(uint8) -- two separately rounded
// fp32 operations (no FMA contraction) and truncation, so the bytes equal the reference's on identical inputs.
__device__ __forceinline__ uint8_t to_u8(float v) { return (uint8_t)fminf(fmaxf(__fadd_rn(__fmul_rn(v, 127.5f), 127.5f), 0.f), 255.f); }
template <typename T>
__global__ void denorm_u8_kernel(const T* __restrict__ x, int64_t n, uint8_t* __restrict__ out) {
  // grid-stride over groups of 4 elements (the launch grid is capped at 16 CTAs per SM), scalar tail by the last groups' owner
  const int64_t n4 = n / 4;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = gtid; g < n4; g += gsz) {
    float v[4];
    Vec4<T>::load(x + 4 * g, v);
    uchar4 o;
    o.x = to_u8(v[0]); o.y = to_u8(v[1]); o.z = to_u8(v[2]); o.w = to_u8(v[3]);
    *reinterpret_cast<uchar4*>(out + 4 * g) = o;
  }
  for (int64_t j = 4 * n4 + gtid; j < n; j += gsz) out[j] = to_u8(to_f<T>(x[j]));
}

}  // namespace
}  // namespace vg

using namespace vg;

extern "C" int vg_cast_scale(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, const float* num,
                             const float* den, void* stream) {
  if (n == 0) return VG_OK;
  cudaStream_t st = as_stream(stream);
  const int grid = grid1d((n + 3) / 4, 256);
  if (src_dtype == VG_F32 && dst_dtype == VG_BF16)
    launch_pdl(cast_scale_kernel<float, bf16>, dim3(grid), dim3(256), 0, st, (const float*)src, (bf16*)dst, n, num, den);
  else if (src_dtype == VG_F32 && dst_dtype == VG_F32)
    launch_pdl(cast_scale_kernel<float, float>, dim3(grid), dim3(256), 0, st, (const float*)src, (float*)dst, n, num, den);
  else if (src_dtype == VG_BF16 && dst_dtype == VG_F32)
    launch_pdl(cast_scale_kernel<bf16, float>, dim3(grid), dim3(256), 0, st, (const bf16*)src, (float*)dst, n, num, den);
  else if (src_dtype == VG_BF16 && dst_dtype == VG_BF16)
    launch_pdl(cast_scale_kernel<bf16, bf16>, dim3(grid), dim3(256), 0, st, (const bf16*)src, (bf16*)dst, n, num, den);
  else
    VG_REQUIRE(false, VG_ERR_ARG, "cast_scale: bad dtypes %d -> %d", src_dtype, dst_dtype);
  return check_launch("cast_scale");
}

extern "C" int vg_pack_pad(const void* const* srcs, int n, int rows, int cols, int rows_pad, int cols_pad, const float* num,
                           const float* den, void* dst, int dst_dtype, void* stream) {
  VG_REQUIRE(n > 0 && rows > 0 && cols > 0 && rows_pad >= rows && cols_pad >= cols, VG_ERR_SHAPE, "pack_pad: bad shape n=%d %dx%d -> %dx%d", n, rows,
             cols, rows_pad, cols_pad);
  cudaStream_t st = as_stream(stream);
  const int grid = grid1d((int64_t)n * rows_pad * cols_pad, 256);
  if (dst_dtype == VG_BF16)
    launch_pdl(pack_pad_kernel<bf16>, dim3(grid), dim3(256), 0, st, (const float* const*)srcs, n, rows, cols, rows_pad, cols_pad, num, den, (bf16*)dst);
  else if (dst_dtype == VG_F32)
    launch_pdl(pack_pad_kernel<float>, dim3(grid), dim3(256), 0, st, (const float* const*)srcs, n, rows, cols, rows_pad, cols_pad, num, den, (float*)dst);
  else
    VG_REQUIRE(false, VG_ERR_ARG, "pack_pad: bad dst dtype %d", dst_dtype);
  return check_launch("pack_pad");
}

extern "C" int vg_colsum(const void* x, int dtype, int64_t M, int N, int64_t ldx, float* out, float* workspace, int ws_rows,
                         unsigned* counter, void* stream) {
  if (M == 0 || N == 0) return VG_OK;
  const int V = dtype == VG_F32 ? 4 : 8;
  const bool vec = (N % V == 0) && (ldx % V == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  cudaStream_t st = as_stream(stream);
  if (vec && workspace && counter && ws_rows >= 1 && N <= 32 * V) {
    // low-contention path: blockIdx.x = row slice; workspace = persistent zeroed [ws_rows = R][N] replicated accumulators
    int ys = (int)max((int64_t)1, min((int64_t)(2 * num_sms()), (M + 63) / 64));
    const int64_t rows_per_cta = (M + ys - 1) / ys;
    ys = (int)((M + rows_per_cta - 1) / rows_per_cta);
    if (dtype == VG_F32) launch_pdl(colsum_vec_kernel<float>, dim3(dim3(ys, 1)), dim3(256), 0, st, (const float*)x, M, N, ldx, out, rows_per_cta, workspace, ws_rows, counter);
    else launch_pdl(colsum_vec_kernel<bf16>, dim3(dim3(ys, 1)), dim3(256), 0, st, (const bf16*)x, M, N, ldx, out, rows_per_cta, workspace, ws_rows, counter);
    return check_launch("colsum");
  }
  const int xs = vec ? (N / V + 31) / 32 : (N + 31) / 32;
  int ys = (int)max((int64_t)1, min((M + 127) / 128, (int64_t)(4 * num_sms() + xs - 1) / xs));
  const int64_t rows_per_cta = (M + ys - 1) / ys;
  ys = (int)((M + rows_per_cta - 1) / rows_per_cta);
  dim3 grid(xs, ys);
  if (vec) {
    if (dtype == VG_F32) launch_pdl(colsum_vec_kernel<float>, dim3(grid), dim3(256), 0, st, (const float*)x, M, N, ldx, out, rows_per_cta, nullptr, 0, nullptr);
    else launch_pdl(colsum_vec_kernel<bf16>, dim3(grid), dim3(256), 0, st, (const bf16*)x, M, N, ldx, out, rows_per_cta, nullptr, 0, nullptr);
  } else {
    if (dtype == VG_F32) colsum_kernel<float><<<grid, 256, 0, st>>>((const float*)x, M, N, ldx, out, rows_per_cta);
    else colsum_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, M, N, ldx, out, rows_per_cta);
  }
  return check_launch("colsum");
}

extern "C" int vg_add_inplace(int dtype, void* x, const void* y, int64_t n, void* stream) {
  if (n == 0) return VG_OK;
  const int grid = grid1d((n + 3) / 4, 256);
  if (dtype == VG_F32) add_inplace_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((float*)x, (const float*)y, n);
  else add_inplace_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>((bf16*)x, (const bf16*)y, n);
  return check_launch("add_inplace");
}

extern "C" int vg_broadcast_rows(int dtype, const void* src, int64_t src_rows, int64_t cols, void* dst, int64_t reps,
                                 void* stream) {
  const int64_t n = src_rows * cols;
  if (n == 0 || reps == 0) return VG_OK;
  const int grid = grid1d(n * reps, 256);
  if (dtype == VG_F32) broadcast_rows_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)src, n, (float*)dst, reps);
  else broadcast_rows_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>((const bf16*)src, n, (bf16*)dst, reps);
  return check_launch("broadcast_rows");
}

extern "C" int vg_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                            float beta1, float beta2, float eps, float weight_decay, int decoupled, float grad_scale,
                            int* step_count, void* bf16_shadow, void* stream) {
  VG_REQUIRE(step_count != nullptr, VG_ERR_ARG, "adam_step: step_count is NULL");
  cudaStream_t st = as_stream(stream);
  if (n > 0)
    adam_kernel<<<grid1d(n, 256), 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                                decoupled, grad_scale, step_count, (bf16*)bf16_shadow);
  bump_kernel<<<1, 1, 0, st>>>(step_count);
  return check_launch("adam_step");
}

extern "C" int vg_copy_rows(int dtype, int64_t rows, int cols, const void* src, int64_t ld_src, void* dst, int64_t ld_dst,
                            void* stream) {
  if (rows == 0 || cols == 0) return VG_OK;
  const int grid = grid1d(rows * cols, 256);
  if (dtype == VG_F32) copy_rows_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(rows, cols, (const float*)src, ld_src, (float*)dst, ld_dst);
  else copy_rows_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>(rows, cols, (const bf16*)src, ld_src, (bf16*)dst, ld_dst);
  return check_launch("copy_rows");
}

extern "C" int vg_act_backward(int dtype, int64_t n, const void* dy, const void* aux, int act, float act_param, void* out,
                               void* stream) {
  int bact;
  switch (act) {
    case VG_ACT_GELU: bact = VG_ACT_MUL_DGELU; break;
    case VG_ACT_TANH: bact = VG_ACT_MUL_DTANH; break;
    case VG_ACT_SIN: bact = VG_ACT_MUL_DSIN; break;
    case VG_ACT_SIGMOID: bact = VG_ACT_MUL_DSIGMOID; break;
    default: VG_REQUIRE(false, VG_ERR_ARG, "act_backward: activation %d has no derivative kernel", act);
  }
  if (n == 0) return VG_OK;
  const int grid = grid1d(n, 256);
  if (dtype == VG_F32) act_bwd_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(n, (const float*)dy, (const float*)aux, bact, act_param, (float*)out);
  else act_bwd_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>(n, (const bf16*)dy, (const bf16*)aux, bact, act_param, (bf16*)out);
  return check_launch("act_backward");
}

extern "C" int vg_softmax_ce(const float* logits, const int64_t* targets, int rows, int C, int rows_per_group, float* losses,
                             float* dlogits, void* stream) {
  VG_REQUIRE(logits && targets && losses && dlogits, VG_ERR_ARG, "softmax_ce: NULL argument");
  VG_REQUIRE(rows >= 1 && C >= 1 && rows_per_group >= 1 && rows % rows_per_group == 0 && rows / rows_per_group <= 64, VG_ERR_SHAPE,
             "softmax_ce: rows %d must be a multiple of rows_per_group %d with at most 64 groups", rows, rows_per_group);
  softmax_ce_kernel<<<1, 1024, 0, as_stream(stream)>>>(logits, targets, rows, C, rows_per_group, losses, dlogits);
  return check_launch("softmax_ce");
}

extern "C" int vg_bce(const float* prob, const float* target, int rows, int rows_per_group, float* losses, float* dprob, void* stream) {
  VG_REQUIRE(prob && target && losses && dprob, VG_ERR_ARG, "bce: NULL argument");
  VG_REQUIRE(rows >= 1 && rows_per_group >= 1 && rows % rows_per_group == 0 && rows / rows_per_group <= 64, VG_ERR_SHAPE,
             "bce: rows %d must be a multiple of rows_per_group %d with at most 64 groups", rows, rows_per_group);
  bce_kernel<<<1, 1024, 0, as_stream(stream)>>>(prob, target, rows, rows_per_group, losses, dprob);
  return check_launch("bce");
}

extern "C" int vg_denorm_u8(int dtype, const void* x, int64_t n, uint8_t* out, void* stream) {
  if (n == 0) return VG_OK;
  VG_REQUIRE(x && out, VG_ERR_ARG, "denorm_u8: NULL argument");
  VG_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0, VG_ERR_ALIGN, "denorm_u8: unaligned buffers");
  const int grid = grid1d((n + 3) / 4, 256);
  if (dtype == VG_F32) denorm_u8_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)x, n, out);
  else denorm_u8_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>((const bf16*)x, n, out);
  return check_launch("denorm_u8");
}
