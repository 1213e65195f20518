// embed.cu -- token construction around the patch-projection GEMM (HBM-bound gathers/scatters).
//   v2: conv2d(k=s=P) == im2col (non-overlapping) + GEMM        src/v2/modules.py:82-100
//   v1: unfold x2 + contiguous().view in the reference's scrambled layout   src/v1/patch_encoder.py:54-73
// plus the CLS-row fill and the backward split (dx -> dtok, dcls, dpos).
#include "common.cuh"

namespace vg {
namespace {

// v2: patches[(b*Np + py*G + px), c*P*P + i*P + j] = img[b, c, py*P+i, px*P+j]; one thread per 4 consecutive j
template <typename T, bool FWD>
__global__ void im2col_kernel(int B, int C, int I, int P, const float* __restrict__ img_in, T* __restrict__ patches,
                              const T* __restrict__ dpatches, float* __restrict__ img_out) {
  const int G = I / P, K = C * P * P;
  const int64_t total = (int64_t)B * C * I * (I / 4);
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    // thread t covers image elements [4t, 4t+4) (contiguous along x; P % 4 == 0 keeps them in one patch row)
    const int64_t e = t * 4;
    const int x = e % I, y = (e / I) % I, c = (e / ((int64_t)I * I)) % C;
    const int64_t b = e / ((int64_t)I * I * C);
    const int px = x / P, j = x % P, py = y / P, i = y % P;
    const int64_t row = b * G * G + (int64_t)py * G + px;
    const int col = c * P * P + i * P + j;
    if (FWD) {
      float v[4];
      Vec4<float>::load(img_in + e, v);
      Vec4<T>::store(patches + row * K + col, v);
    } else {
      float v[4];
      Vec4<T>::load(dpatches + row * K + col, v);
      Vec4<float>::store(img_out + e, v);
    }
  }
}

// v1 scrambled tokens: flat index f over (C, n, n, win, win) of image b  <->  tokens[b, f / 432, f % 432]
template <typename T, bool FWD>
__global__ void v1_tokens_kernel(int B, int C, int I, int win, int stride, int n, const float* __restrict__ img,
                                 T* __restrict__ tokens, const T* __restrict__ dtokens, float* __restrict__ dimg) {
  const int64_t per_img = (int64_t)C * n * n * win * win;
  const int64_t total = (int64_t)B * per_img;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / per_img;
    int64_t f = t % per_img;
    const int wx = f % win; f /= win;
    const int wy = f % win; f /= win;
    const int iw = f % n; f /= n;
    const int ih = f % n; f /= n;
    const int c = (int)f;
    const int64_t src = ((b * C + c) * I + (ih * stride + wy)) * I + (iw * stride + wx);
    if (FWD) tokens[t] = from_f<T>(img[src]);
    else atomicAdd(&dimg[src], to_f<T>(dtokens[t]));
  }
}

template <typename T>
__global__ void fill_rows_kernel(int B, int S, int E, int row, const float* __restrict__ v, const float* __restrict__ v2,
                                 T* __restrict__ x) {
  const int64_t total = (int64_t)B * E;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int e = t % E;
    const int64_t b = t / E;
    float val = v[e];
    if (v2) val += v2[e];
    x[(b * S + row) * E + e] = from_f<T>(val);
  }
}

// dx (B,S,E): each CTA owns a (token s, 128-column slab) and loops over a slice of the batch:
// writes dtok (s>0), accumulates sum over b in registers, one atomic per (s, e) per CTA.
template <typename T>
__global__ void embed_bwd_split_kernel(int B, int S, int E, const T* __restrict__ dx, T* __restrict__ dtok,
                                       float* __restrict__ dcls, float* __restrict__ dpos, int pos_has_cls) {
  const int s = blockIdx.x;
  const int e = blockIdx.y * blockDim.x + threadIdx.x;
  if (e >= E) return;
  float acc = 0.f;
  for (int b = blockIdx.z; b < B; b += gridDim.z) {
    const T val = dx[((int64_t)b * S + s) * E + e];
    acc += to_f<T>(val);
    if (s > 0) dtok[((int64_t)b * (S - 1) + (s - 1)) * E + e] = val;
  }
  if (s == 0) {
    atomicAdd(&dcls[e], acc);
    if (pos_has_cls) atomicAdd(&dpos[e], acc);
  } else {
    atomicAdd(&dpos[(int64_t)(pos_has_cls ? s : s - 1) * E + e], acc);
  }
}

// vectorised variant (E a multiple of 16 B worth of elements, E/V <= 256): thread = (batch lane ty, 16-byte column group);
// CTA = token s x a slice of the batch; 16 B loads/stores, two batch rows in flight per thread, smem fold over the batch lanes.
template <typename T>
__global__ void __launch_bounds__(256)
embed_bwd_split_vec_kernel(int B, int S, int E, const T* __restrict__ dx, T* __restrict__ dtok, float* __restrict__ dcls,
                           float* __restrict__ dpos, int pos_has_cls) {
  constexpr int V = 16 / sizeof(T);
  __shared__ float red[256 * V];
  const int s = blockIdx.x, cgs = E / V, lanes = 256 / cgs;
  const int cg = threadIdx.x % cgs, ty = threadIdx.x / cgs;
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.f;
  if (ty < lanes) {
    const int step = gridDim.z * lanes;
    int b = blockIdx.z * lanes + ty;
    for (; b + step < B; b += 2 * step) {
      const uint4 v0 = *reinterpret_cast<const uint4*>(dx + ((int64_t)b * S + s) * E + cg * V);
      const uint4 v1 = *reinterpret_cast<const uint4*>(dx + ((int64_t)(b + step) * S + s) * E + cg * V);
      const T* e0 = reinterpret_cast<const T*>(&v0);
      const T* e1 = reinterpret_cast<const T*>(&v1);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] += to_f<T>(e0[j]) + to_f<T>(e1[j]);
      if (s > 0) {
        *reinterpret_cast<uint4*>(dtok + ((int64_t)b * (S - 1) + (s - 1)) * E + cg * V) = v0;
        *reinterpret_cast<uint4*>(dtok + ((int64_t)(b + step) * (S - 1) + (s - 1)) * E + cg * V) = v1;
      }
    }
    for (; b < B; b += step) {
      const uint4 v0 = *reinterpret_cast<const uint4*>(dx + ((int64_t)b * S + s) * E + cg * V);
      const T* e0 = reinterpret_cast<const T*>(&v0);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] += to_f<T>(e0[j]);
      if (s > 0) *reinterpret_cast<uint4*>(dtok + ((int64_t)b * (S - 1) + (s - 1)) * E + cg * V) = v0;
    }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) red[threadIdx.x * V + j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < E; c += blockDim.x) {       // column c lives in group c / V, element c % V of every batch lane
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += red[(l * cgs + c / V) * V + (c % V)];
    if (s == 0) {
      atomicAdd(&dcls[c], t);
      if (pos_has_cls) atomicAdd(&dpos[c], t);
    } else {
      atomicAdd(&dpos[(int64_t)(pos_has_cls ? s : s - 1) * E + c], t);
    }
  }
}

int grid1d(int64_t total, int block) {
  const int64_t need = (total + block - 1) / block;
  return (int)max((int64_t)1, min(need, (int64_t)num_sms() * 16));
}

}  // namespace
}  // namespace vg

using namespace vg;

extern "C" int vg_im2col_patches(int dtype, int B, int C, int I, int P, const float* img, void* patches, void* stream) {
  VG_REQUIRE(P % 4 == 0 && I % P == 0, VG_ERR_SHAPE, "im2col: need P %% 4 == 0 and I %% P == 0 (I=%d P=%d)", I, P);
  const int64_t total = (int64_t)B * C * I * (I / 4);
  if (dtype == VG_F32)
    im2col_kernel<float, true><<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(B, C, I, P, img, (float*)patches, nullptr, nullptr);
  else
    im2col_kernel<bf16, true><<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(B, C, I, P, img, (bf16*)patches, nullptr, nullptr);
  return check_launch("im2col_patches");
}

extern "C" int vg_col2im_patches(int dtype, int B, int C, int I, int P, const void* dpatches, float* dimg, void* stream) {
  VG_REQUIRE(P % 4 == 0 && I % P == 0, VG_ERR_SHAPE, "col2im: need P %% 4 == 0 and I %% P == 0 (I=%d P=%d)", I, P);
  const int64_t total = (int64_t)B * C * I * (I / 4);
  if (dtype == VG_F32)
    im2col_kernel<float, false><<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(B, C, I, P, nullptr, nullptr, (const float*)dpatches, dimg);
  else
    im2col_kernel<bf16, false><<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(B, C, I, P, nullptr, nullptr, (const bf16*)dpatches, dimg);
  return check_launch("col2im_patches");
}

extern "C" int vg_v1_tokens_fwd(int dtype, int B, int C, int I, int win, int stride, int n_side, const float* img,
                                void* tokens, void* stream) {
  VG_REQUIRE((n_side - 1) * stride + win <= I, VG_ERR_SHAPE, "v1_tokens: windows exceed the image");
  const int64_t total = (int64_t)B * C * n_side * n_side * win * win;
  if (dtype == VG_F32)
    v1_tokens_kernel<float, true><<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(B, C, I, win, stride, n_side, img, (float*)tokens, nullptr, nullptr);
  else
    v1_tokens_kernel<bf16, true><<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(B, C, I, win, stride, n_side, img, (bf16*)tokens, nullptr, nullptr);
  return check_launch("v1_tokens_fwd");
}

extern "C" int vg_v1_tokens_bwd(int dtype, int B, int C, int I, int win, int stride, int n_side, const void* dtokens,
                                float* dimg, void* stream) {
  VG_REQUIRE((n_side - 1) * stride + win <= I, VG_ERR_SHAPE, "v1_tokens: windows exceed the image");
  const int64_t total = (int64_t)B * C * n_side * n_side * win * win;
  if (dtype == VG_F32)
    v1_tokens_kernel<float, false><<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(B, C, I, win, stride, n_side, nullptr, nullptr, (const float*)dtokens, dimg);
  else
    v1_tokens_kernel<bf16, false><<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(B, C, I, win, stride, n_side, nullptr, nullptr, (const bf16*)dtokens, dimg);
  return check_launch("v1_tokens_bwd");
}

extern "C" int vg_fill_rows(int dtype, int B, int S, int E, int row, const float* v, const float* v2, void* x, void* stream) {
  VG_REQUIRE(row >= 0 && row < S, VG_ERR_ARG, "fill_rows: row %d outside [0,%d)", row, S);
  const int64_t total = (int64_t)B * E;
  if (dtype == VG_F32)
    fill_rows_kernel<float><<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(B, S, E, row, v, v2, (float*)x);
  else
    fill_rows_kernel<bf16><<<grid1d(total, 256), 256, 0, as_stream(stream)>>>(B, S, E, row, v, v2, (bf16*)x);
  return check_launch("fill_rows");
}

extern "C" int vg_embed_bwd_split(int dtype, int B, int S, int E, const void* dx, void* dtok, float* dcls, float* dpos,
                                  int pos_has_cls, void* stream) {
  const int vel = dtype == VG_F32 ? 4 : 8;
  if (E % vel == 0 && E / vel <= 256 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0 && (reinterpret_cast<uintptr_t>(dtok) & 15) == 0) {
    const int lanes = 256 / (E / vel);
    const int zs = max(1, min((B + 2 * lanes - 1) / (2 * lanes), (4 * num_sms() + S - 1) / S));     // >= 2 batch rows per thread, ~4 CTAs per SM
    dim3 g(S, 1, zs);
    if (dtype == VG_F32)
      embed_bwd_split_vec_kernel<float><<<g, 256, 0, as_stream(stream)>>>(B, S, E, (const float*)dx, (float*)dtok, dcls, dpos, pos_has_cls);
    else
      embed_bwd_split_vec_kernel<bf16><<<g, 256, 0, as_stream(stream)>>>(B, S, E, (const bf16*)dx, (bf16*)dtok, dcls, dpos, pos_has_cls);
    return check_launch("embed_bwd_split");
  }
  const int bx = E >= 128 ? 128 : 32;
  int zsplit = max(1, min(B, (4 * num_sms()) / max(1, S * ((E + bx - 1) / bx))));
  dim3 grid(S, (E + bx - 1) / bx, zsplit);
  if (dtype == VG_F32)
    embed_bwd_split_kernel<float><<<grid, bx, 0, as_stream(stream)>>>(B, S, E, (const float*)dx, (float*)dtok, dcls, dpos, pos_has_cls);
  else
    embed_bwd_split_kernel<bf16><<<grid, bx, 0, as_stream(stream)>>>(B, S, E, (const bf16*)dx, (bf16*)dtok, dcls, dpos, pos_has_cls);
  return check_launch("embed_bwd_split");
}
