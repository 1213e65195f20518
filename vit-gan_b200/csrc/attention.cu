// attention.cu -- multi-head self-attention core, flash style (online softmax, no SxS tensor in HBM),
// dot-product and L2-distance score variants, forward and backward.  CUDA-core version (fp32 math):
// serves the fp32 parity path and is the functional baseline the tensor-core attention is checked against.
//
//   forward : src/v2/modules.py:142-155 (scores/sqrt(d), softmax, PV) ; src/v1/attention.py:51,66-70
//   L2 mode : torch.cdist(q,k,p=2) in its matmul form sqrt(clamp_min(|q|^2+|k|^2-2q.k, 0))
//   backward: autograd of the above; P is recomputed from (q,k,lse); for L2:
//             dq_i = sum_j g_ij (q_i-k_j), dk_j = sum_i g_ij (k_j-q_i), g = dS*scale/dist (0 where dist==0)
//
// Layout: element (b,s,h,c) at base + (b*S+s)*ld + h*d + c  -> heads are sliced in place.
// Work split: CTA = (block of RB "own" rows, (b,h)).  The "staged" operand pair (K,V for fwd/dQ; Q,dO for
// dK/dV) is brought to shared memory as fp32 in tiles of KT rows, each tile exactly once per CTA; a warp
// then sweeps its own rows over the tile: lanes span 32 staged rows for the scores, then the head dims for
// the outputs.  Per-row running state (m, l, acc) lives in shared memory between tiles (one tile if S fits).
#include <math.h>

#include "common.cuh"

namespace vg {
namespace {

constexpr int NW = 4;         // warps per CTA
constexpr int MAXD = 256;
constexpr int MAXT = MAXD / 32;
constexpr size_t SMEM_BUDGET = 200 * 1024;

struct Plan { int KT, RB, SW; size_t bytes; };

struct Smem {
  float *X, *Y, *xn, *st0, *st1, *a, *b, *p, *p2, *state;
};

__host__ __device__ inline size_t carve(Smem* s, float* base, int d, int KT, int RB, int SW) {
  const int dp = d + 1, dpad = (d + 3) & ~3;
  size_t off = 0;
  auto take = [&](size_t n) { float* r = base ? base + off : nullptr; off += n; return r; };
  float* X = take((size_t)KT * dp); float* Y = take((size_t)KT * dp);
  float* xn = take(KT); float* st0 = take(KT); float* st1 = take(KT);
  float* a = take((size_t)NW * dpad); float* b = take((size_t)NW * dpad);
  float* p = take(NW * 32); float* p2 = take(NW * 32);
  float* state = take((size_t)RB * SW);
  if (s) { s->X = X; s->Y = Y; s->xn = xn; s->st0 = st0; s->st1 = st1; s->a = a; s->b = b; s->p = p; s->p2 = p2; s->state = state; }
  return off * sizeof(float);
}

// state width per own row: fwd d+2 (acc,m,l); dq d+1 (acc,gsum); dkv 2d+1
Plan make_plan(int S, int d, int state_mult, int state_extra) {
  Plan p;
  const int dpad = (d + 3) & ~3;
  p.RB = S <= 128 ? S : 64;
  p.SW = state_mult * dpad + state_extra;
  int kt = ((S + 31) / 32) * 32;
  while (kt > 32 && carve(nullptr, nullptr, d, kt, p.RB, p.SW) > SMEM_BUDGET) kt -= 32;
  p.KT = kt;
  p.bytes = carve(nullptr, nullptr, d, kt, p.RB, p.SW);
  return p;
}

template <typename T>
__device__ __forceinline__ void stage_rows(float* dst, const T* src, int64_t ld, int row0, int n, int KT, int d) {
  const int dp = d + 1, vec_per_row = d / 4;
  for (int idx = threadIdx.x; idx < KT * vec_per_row; idx += NW * 32) {
    const int r = idx / vec_per_row, c = (idx % vec_per_row) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < n) Vec4<T>::load(src + (int64_t)(row0 + r) * ld + c, v);
    float* o = dst + r * dp + c;
    o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = v[3];
  }
}
template <typename T>
__device__ __forceinline__ void load_own(float* dst, const T* src, int d, int lane) {
  for (int c = lane * 4; c < d; c += 128) {
    float v[4];
    Vec4<T>::load(src + c, v);
    dst[c] = v[0]; dst[c + 1] = v[1]; dst[c + 2] = v[2]; dst[c + 3] = v[3];
  }
}
__device__ __forceinline__ float dot_row(const float* own, const float* staged_row, int d) {
  float s = 0.f;
#pragma unroll 4
  for (int c = 0; c < d; ++c) s = fmaf(own[c], staged_row[c], s);
  return s;
}
__device__ __forceinline__ float sqnorm_own(const float* own, int d, int lane) {
  float s = 0.f;
  for (int c = lane; c < d; c += 32) s = fmaf(own[c], own[c], s);
  return warp_sum(s);
}
__device__ __forceinline__ void staged_sqnorm(const Smem& sm, int n, int d) {
  for (int r = threadIdx.x; r < n; r += NW * 32) {
    const float* row = sm.X + r * (d + 1);
    float s = 0.f;
    for (int c = 0; c < d; ++c) s = fmaf(row[c], row[c], s);
    sm.xn[r] = s;
  }
}

struct Geo {
  int H, S, d, KT, RB, SW;
  int64_t ld, ldo, ldd;
  float scale;
};

// ------------------------------------------------------------------------------------------------ forward
template <typename T, int MODE>
__global__ void __launch_bounds__(NW * 32)
attn_fwd_kernel(Geo g, const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                T* __restrict__ o, float* __restrict__ lse) {
  extern __shared__ __align__(16) float smem_raw[];
  Smem sm;
  carve(&sm, smem_raw, g.d, g.KT, g.RB, g.SW);
  const int d = g.d, S = g.S, dp = d + 1, dpad = (d + 3) & ~3;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int bh = blockIdx.x, b = bh / g.H, h = bh % g.H;
  const int r0 = blockIdx.y * g.RB;
  const int nrows = min(g.RB, S - r0);
  const int64_t base = (int64_t)b * S * g.ld + (int64_t)h * d;
  const T* qb = q + base; const T* kb = k + base; const T* vb = v + base;
  T* ob = o + (int64_t)b * S * g.ldo + (int64_t)h * d;
  float* own = sm.a + wid * dpad;
  float* pw = sm.p + wid * 32;
  const int nt = (d + 31) / 32;

  for (int t0 = 0; t0 < S; t0 += g.KT) {
    const int tn = min(g.KT, S - t0);
    const bool first = t0 == 0, last = t0 + g.KT >= S;
    __syncthreads();
    stage_rows<T>(sm.X, kb, g.ld, t0, tn, g.KT, d);
    stage_rows<T>(sm.Y, vb, g.ld, t0, tn, g.KT, d);
    __syncthreads();
    if (MODE == VG_ATTN_L2) { staged_sqnorm(sm, tn, d); __syncthreads(); }

    for (int i = wid; i < nrows; i += NW) {
      load_own<T>(own, qb + (int64_t)(r0 + i) * g.ld, d, lane);
      __syncwarp();
      const float qq = (MODE == VG_ATTN_L2) ? sqnorm_own(own, d, lane) : 0.f;
      float* st = sm.state + (size_t)i * g.SW;
      float m = -INFINITY, l = 0.f, acc[MAXT];
#pragma unroll
      for (int t = 0; t < MAXT; ++t) acc[t] = 0.f;
      if (!first) {
        m = st[dpad]; l = st[dpad + 1];
#pragma unroll
        for (int t = 0; t < MAXT; ++t) { const int c = t * 32 + lane; if (t < nt && c < d) acc[t] = st[c]; }
      }
      for (int j0 = 0; j0 < tn; j0 += 32) {
        const int n = min(32, tn - j0);
        float s = -INFINITY;
        if (lane < n) {
          s = dot_row(own, sm.X + (j0 + lane) * dp, d);
          if (MODE == VG_ATTN_L2) s = sqrtf(fmaxf(qq + sm.xn[j0 + lane] - 2.f * s, 0.f));
          s *= g.scale;
        }
        const float m_new = fmaxf(m, warp_max(s));
        const float p = (lane < n) ? __expf(s - m_new) : 0.f;
        const float corr = __expf(m - m_new);      // m = -inf on the first chunk -> 0
        l = l * corr + warp_sum(p);
        m = m_new;
        pw[lane] = p;
        __syncwarp();
#pragma unroll
        for (int t = 0; t < MAXT; ++t) {
          if (t < nt) {
            const int c = t * 32 + lane;
            float a = acc[t] * corr;
            if (c < d) {
              const float* yc = sm.Y + (size_t)j0 * dp + c;
              for (int j = 0; j < n; ++j) a = fmaf(pw[j], yc[j * dp], a);
            }
            acc[t] = a;
          }
        }
        __syncwarp();
      }
      if (last) {
        const float inv = 1.0f / l;
        T* orow = ob + (int64_t)(r0 + i) * g.ldo;
#pragma unroll
        for (int t = 0; t < MAXT; ++t) { const int c = t * 32 + lane; if (t < nt && c < d) orow[c] = from_f<T>(acc[t] * inv); }
        if (lane == 0) lse[(int64_t)bh * S + r0 + i] = m + __logf(l);
      } else {
#pragma unroll
        for (int t = 0; t < MAXT; ++t) { const int c = t * 32 + lane; if (t < nt && c < d) st[c] = acc[t]; }
        if (lane == 0) { st[dpad] = m; st[dpad + 1] = l; }
      }
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------------ delta = rowsum(dO * O)
template <typename T>
__global__ void attn_delta_kernel(int B, int H, int S, int d, const T* __restrict__ o, const T* __restrict__ d_o,
                                  int64_t ldo, float* __restrict__ delta) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // over B*H*S
  if (row >= (int64_t)B * H * S) return;
  const int s = row % S, bh = row / S, b = bh / H, h = bh % H;
  const int64_t off = ((int64_t)b * S + s) * ldo + (int64_t)h * d;
  float acc = 0.f;
  for (int c = lane; c < d; c += 32) acc = fmaf(to_f<T>(o[off + c]), to_f<T>(d_o[off + c]), acc);
  acc = warp_sum(acc);
  if (lane == 0) delta[row] = acc;
}

// ------------------------------------------------------------------------------------------------ backward: dQ
// own rows = queries (a: q_i, b: dO_i), staged = K (X), V (Y)
template <typename T, int MODE>
__global__ void __launch_bounds__(NW * 32)
attn_bwd_dq_kernel(Geo g, const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                   const T* __restrict__ d_o, const float* __restrict__ lse, const float* __restrict__ delta,
                   T* __restrict__ dq) {
  extern __shared__ __align__(16) float smem_raw[];
  Smem sm;
  carve(&sm, smem_raw, g.d, g.KT, g.RB, g.SW);
  const int d = g.d, S = g.S, dp = d + 1, dpad = (d + 3) & ~3;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int bh = blockIdx.x, b = bh / g.H, h = bh % g.H;
  const int r0 = blockIdx.y * g.RB;
  const int nrows = min(g.RB, S - r0);
  const int64_t base = (int64_t)b * S * g.ld + (int64_t)h * d;
  const T* qb = q + base; const T* kb = k + base; const T* vb = v + base;
  const T* dob = d_o + (int64_t)b * S * g.ldo + (int64_t)h * d;
  T* dqb = dq + (int64_t)b * S * g.ldd + (int64_t)h * d;
  float* own = sm.a + wid * dpad;
  float* own_do = sm.b + wid * dpad;
  float* pw = sm.p + wid * 32;
  const int nt = (d + 31) / 32;

  for (int t0 = 0; t0 < S; t0 += g.KT) {
    const int tn = min(g.KT, S - t0);
    const bool first = t0 == 0, last = t0 + g.KT >= S;
    __syncthreads();
    stage_rows<T>(sm.X, kb, g.ld, t0, tn, g.KT, d);
    stage_rows<T>(sm.Y, vb, g.ld, t0, tn, g.KT, d);
    __syncthreads();
    if (MODE == VG_ATTN_L2) { staged_sqnorm(sm, tn, d); __syncthreads(); }

    for (int i = wid; i < nrows; i += NW) {
      load_own<T>(own, qb + (int64_t)(r0 + i) * g.ld, d, lane);
      load_own<T>(own_do, dob + (int64_t)(r0 + i) * g.ldo, d, lane);
      const float lse_i = lse[(int64_t)bh * S + r0 + i];
      const float del_i = delta[(int64_t)bh * S + r0 + i];
      __syncwarp();
      const float qq = (MODE == VG_ATTN_L2) ? sqnorm_own(own, d, lane) : 0.f;
      float* st = sm.state + (size_t)i * g.SW;
      float gsum = 0.f, acc[MAXT];
#pragma unroll
      for (int t = 0; t < MAXT; ++t) acc[t] = 0.f;
      if (!first) {
        gsum = st[dpad];
#pragma unroll
        for (int t = 0; t < MAXT; ++t) { const int c = t * 32 + lane; if (t < nt && c < d) acc[t] = st[c]; }
      }
      for (int j0 = 0; j0 < tn; j0 += 32) {
        const int n = min(32, tn - j0);
        float gg = 0.f;
        if (lane < n) {
          float s = dot_row(own, sm.X + (j0 + lane) * dp, d);
          float dist = 1.f;
          if (MODE == VG_ATTN_L2) { dist = sqrtf(fmaxf(qq + sm.xn[j0 + lane] - 2.f * s, 0.f)); s = dist; }
          const float p = __expf(s * g.scale - lse_i);
          const float dpv = dot_row(own_do, sm.Y + (j0 + lane) * dp, d);
          gg = p * (dpv - del_i) * g.scale;                     // dL/d(raw score)
          if (MODE == VG_ATTN_L2) gg = dist > 0.f ? gg / dist : 0.f;
        }
        if (MODE == VG_ATTN_L2) gsum += warp_sum(gg);
        pw[lane] = gg;
        __syncwarp();
#pragma unroll
        for (int t = 0; t < MAXT; ++t) {
          const int c = t * 32 + lane;
          if (t < nt && c < d) {
            float a = acc[t];
            const float* xc = sm.X + (size_t)j0 * dp + c;
            for (int j = 0; j < n; ++j) a = fmaf(pw[j], xc[j * dp], a);
            acc[t] = a;
          }
        }
        __syncwarp();
      }
      if (last) {
        T* drow = dqb + (int64_t)(r0 + i) * g.ldd;
#pragma unroll
        for (int t = 0; t < MAXT; ++t) {
          const int c = t * 32 + lane;
          if (t < nt && c < d) drow[c] = from_f<T>((MODE == VG_ATTN_L2) ? own[c] * gsum - acc[t] : acc[t]);
        }
      } else {
#pragma unroll
        for (int t = 0; t < MAXT; ++t) { const int c = t * 32 + lane; if (t < nt && c < d) st[c] = acc[t]; }
        if (lane == 0) st[dpad] = gsum;
      }
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward: dK, dV
// own rows = keys (a: k_j, b: v_j), staged = Q (X), dO (Y), lse (st0), delta (st1)
template <typename T, int MODE>
__global__ void __launch_bounds__(NW * 32)
attn_bwd_dkv_kernel(Geo g, const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                    const T* __restrict__ d_o, const float* __restrict__ lse, const float* __restrict__ delta,
                    T* __restrict__ dk, T* __restrict__ dv) {
  extern __shared__ __align__(16) float smem_raw[];
  Smem sm;
  carve(&sm, smem_raw, g.d, g.KT, g.RB, g.SW);
  const int d = g.d, S = g.S, dp = d + 1, dpad = (d + 3) & ~3;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int bh = blockIdx.x, b = bh / g.H, h = bh % g.H;
  const int r0 = blockIdx.y * g.RB;
  const int nrows = min(g.RB, S - r0);
  const int64_t base = (int64_t)b * S * g.ld + (int64_t)h * d;
  const T* qb = q + base; const T* kb = k + base; const T* vb = v + base;
  const T* dob = d_o + (int64_t)b * S * g.ldo + (int64_t)h * d;
  T* dkb = dk + (int64_t)b * S * g.ldd + (int64_t)h * d;
  T* dvb = dv + (int64_t)b * S * g.ldd + (int64_t)h * d;
  float* own_k = sm.a + wid * dpad;
  float* own_v = sm.b + wid * dpad;
  float* pw = sm.p + wid * 32;
  float* gw = sm.p2 + wid * 32;
  const int nt = (d + 31) / 32;

  for (int t0 = 0; t0 < S; t0 += g.KT) {
    const int tn = min(g.KT, S - t0);
    const bool first = t0 == 0, last = t0 + g.KT >= S;
    __syncthreads();
    stage_rows<T>(sm.X, qb, g.ld, t0, tn, g.KT, d);
    stage_rows<T>(sm.Y, dob, g.ldo, t0, tn, g.KT, d);
    for (int r = threadIdx.x; r < tn; r += NW * 32) {
      sm.st0[r] = lse[(int64_t)bh * S + t0 + r];
      sm.st1[r] = delta[(int64_t)bh * S + t0 + r];
    }
    __syncthreads();
    if (MODE == VG_ATTN_L2) { staged_sqnorm(sm, tn, d); __syncthreads(); }

    for (int j = wid; j < nrows; j += NW) {
      load_own<T>(own_k, kb + (int64_t)(r0 + j) * g.ld, d, lane);
      load_own<T>(own_v, vb + (int64_t)(r0 + j) * g.ld, d, lane);
      __syncwarp();
      const float kk = (MODE == VG_ATTN_L2) ? sqnorm_own(own_k, d, lane) : 0.f;
      float* st = sm.state + (size_t)j * g.SW;
      float gsum = 0.f, acck[MAXT], accv[MAXT];
#pragma unroll
      for (int t = 0; t < MAXT; ++t) { acck[t] = 0.f; accv[t] = 0.f; }
      if (!first) {
        gsum = st[2 * dpad];
#pragma unroll
        for (int t = 0; t < MAXT; ++t) {
          const int c = t * 32 + lane;
          if (t < nt && c < d) { acck[t] = st[c]; accv[t] = st[dpad + c]; }
        }
      }
      for (int i0 = 0; i0 < tn; i0 += 32) {
        const int n = min(32, tn - i0);
        float p = 0.f, gg = 0.f;
        if (lane < n) {
          float s = dot_row(own_k, sm.X + (i0 + lane) * dp, d);
          float dist = 1.f;
          if (MODE == VG_ATTN_L2) { dist = sqrtf(fmaxf(kk + sm.xn[i0 + lane] - 2.f * s, 0.f)); s = dist; }
          p = __expf(s * g.scale - sm.st0[i0 + lane]);
          const float dpv = dot_row(own_v, sm.Y + (i0 + lane) * dp, d);
          gg = p * (dpv - sm.st1[i0 + lane]) * g.scale;
          if (MODE == VG_ATTN_L2) gg = dist > 0.f ? gg / dist : 0.f;
        }
        if (MODE == VG_ATTN_L2) gsum += warp_sum(gg);
        pw[lane] = p;
        gw[lane] = gg;
        __syncwarp();
#pragma unroll
        for (int t = 0; t < MAXT; ++t) {
          const int c = t * 32 + lane;
          if (t < nt && c < d) {
            float ak = acck[t], av = accv[t];
            const float* xc = sm.X + (size_t)i0 * dp + c;
            const float* yc = sm.Y + (size_t)i0 * dp + c;
            for (int i = 0; i < n; ++i) {
              av = fmaf(pw[i], yc[i * dp], av);
              ak = fmaf(gw[i], xc[i * dp], ak);
            }
            acck[t] = ak; accv[t] = av;
          }
        }
        __syncwarp();
      }
      if (last) {
        T* dkrow = dkb + (int64_t)(r0 + j) * g.ldd;
        T* dvrow = dvb + (int64_t)(r0 + j) * g.ldd;
#pragma unroll
        for (int t = 0; t < MAXT; ++t) {
          const int c = t * 32 + lane;
          if (t < nt && c < d) {
            dkrow[c] = from_f<T>((MODE == VG_ATTN_L2) ? own_k[c] * gsum - acck[t] : acck[t]);
            dvrow[c] = from_f<T>(accv[t]);
          }
        }
      } else {
#pragma unroll
        for (int t = 0; t < MAXT; ++t) {
          const int c = t * 32 + lane;
          if (t < nt && c < d) { st[c] = acck[t]; st[dpad + c] = accv[t]; }
        }
        if (lane == 0) st[2 * dpad] = gsum;
      }
      __syncwarp();
    }
  }
}

int check_shape(int B, int H, int S, int d, int64_t ld, int64_t ldo) {
  VG_REQUIRE(B > 0 && H > 0 && S > 0, VG_ERR_SHAPE, "attention: empty problem B=%d H=%d S=%d", B, H, S);
  VG_REQUIRE(d > 0 && d % 4 == 0 && d <= MAXD, VG_ERR_SHAPE, "attention: head dim %d must be a multiple of 4 and <= %d", d, MAXD);
  VG_REQUIRE(ld % 4 == 0 && ldo % 4 == 0, VG_ERR_ALIGN, "attention: leading dims must be multiples of 4 elements");
  return VG_OK;
}

template <typename K>
int set_smem(K kern, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    VG_REQUIRE(e == cudaSuccess, VG_ERR_LAUNCH, "attention: cudaFuncSetAttribute(%zu B): %s", bytes, cudaGetErrorString(e));
  }
  return VG_OK;
}

Geo make_geo(int H, int S, int d, const Plan& p, int64_t ld, int64_t ldo, int64_t ldd, float scale) {
  Geo g; g.H = H; g.S = S; g.d = d; g.KT = p.KT; g.RB = p.RB; g.SW = p.SW; g.ld = ld; g.ldo = ldo; g.ldd = ldd; g.scale = scale;
  return g;
}

template <typename T, int MODE>
int fwd_t(int B, int H, int S, int d, const void* q, const void* k, const void* v, int64_t ld, void* o, int64_t ldo,
          float* lse, float scale, cudaStream_t st) {
  const Plan p = make_plan(S, d, 1, 2);
  int rc = set_smem(attn_fwd_kernel<T, MODE>, p.bytes);
  if (rc) return rc;
  dim3 grid(B * H, (S + p.RB - 1) / p.RB);
  attn_fwd_kernel<T, MODE><<<grid, NW * 32, p.bytes, st>>>(make_geo(H, S, d, p, ld, ldo, 0, scale), (const T*)q, (const T*)k,
                                                           (const T*)v, (T*)o, lse);
  return check_launch("attention_fwd");
}

template <typename T, int MODE>
int bwd_t(int B, int H, int S, int d, const void* q, const void* k, const void* v, int64_t ld, const void* o,
          const void* d_o, int64_t ldo, const float* lse, void* dq, void* dk, void* dv, int64_t ldd, float scale,
          float* delta, cudaStream_t st) {
  const Plan pq = make_plan(S, d, 1, 1), pkv = make_plan(S, d, 2, 1);
  int rc = set_smem(attn_bwd_dq_kernel<T, MODE>, pq.bytes);
  if (rc) return rc;
  rc = set_smem(attn_bwd_dkv_kernel<T, MODE>, pkv.bytes);
  if (rc) return rc;
  const int64_t rows = (int64_t)B * H * S;
  attn_delta_kernel<T><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(B, H, S, d, (const T*)o, (const T*)d_o, ldo, delta);
  attn_bwd_dq_kernel<T, MODE><<<dim3(B * H, (S + pq.RB - 1) / pq.RB), NW * 32, pq.bytes, st>>>(
      make_geo(H, S, d, pq, ld, ldo, ldd, scale), (const T*)q, (const T*)k, (const T*)v, (const T*)d_o, lse, delta, (T*)dq);
  attn_bwd_dkv_kernel<T, MODE><<<dim3(B * H, (S + pkv.RB - 1) / pkv.RB), NW * 32, pkv.bytes, st>>>(
      make_geo(H, S, d, pkv, ld, ldo, ldd, scale), (const T*)q, (const T*)k, (const T*)v, (const T*)d_o, lse, delta, (T*)dk, (T*)dv);
  return check_launch("attention_bwd");
}

}  // namespace
}  // namespace vg

namespace vg {
bool attention_tc_supported(int dtype, int mode, int B, int H, int S, int d, const void* q, const void* k, const void* v,
                            int64_t ld_qkv, const void* o, int64_t ld_o);
int attention_fwd_tc(int B, int H, int S, int d, const void* q, const void* k, const void* v, int64_t ld, void* o, int64_t ldo,
                     float* lse, float scale, cudaStream_t st);
int attention_bwd_tc(int B, int H, int S, int d, const void* q, const void* k, const void* v, int64_t ld, const void* d_o,
                     int64_t ldo, const float* lse, void* dq, void* dk, void* dv, int64_t ldd, float scale, cudaStream_t st);
// multi-tile tcgen05 kernels (attention_mt.cu): d = 96 / 112 / 192, dot or L2 scores, S <= 272
bool attention_mt_supported(int dtype, int mode, int B, int H, int S, int d, const void* q, const void* k, const void* v,
                            int64_t ld_qkv, const void* o, int64_t ld_o);
int attention_fwd_mt(int mode, int B, int H, int S, int d, const void* q, const void* k, const void* v, int64_t ld, void* o, int64_t ldo,
                     float* lse, float scale, cudaStream_t st);
int attention_bwd_mt(int mode, int B, int H, int S, int d, const void* q, const void* k, const void* v, int64_t ld, const void* o,
                     const void* d_o, int64_t ldo, const float* lse, void* dq, void* dk, void* dv, int64_t ldd, float scale, float* delta,
                     cudaStream_t st);
}  // namespace vg

using namespace vg;

extern "C" int vg_attention_fwd(int dtype, int mode, int B, int H, int S, int d, const void* q, const void* k,
                                const void* v, int64_t ld_qkv, void* o, int64_t ld_o, float* lse, float scale,
                                void* stream) {
  int rc = check_shape(B, H, S, d, ld_qkv, ld_o);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  // tensor-core path (attention_tc.cu) for the single-tile bf16 dot-product case; CUDA-core flash kernel otherwise
  if (attention_tc_supported(dtype, mode, B, H, S, d, q, k, v, ld_qkv, o, ld_o))
    return attention_fwd_tc(B, H, S, d, q, k, v, ld_qkv, o, ld_o, lse, scale, st);
  if (attention_mt_supported(dtype, mode, B, H, S, d, q, k, v, ld_qkv, o, ld_o))
    return attention_fwd_mt(mode, B, H, S, d, q, k, v, ld_qkv, o, ld_o, lse, scale, st);
  if (dtype == VG_F32)
    return mode == VG_ATTN_L2 ? fwd_t<float, VG_ATTN_L2>(B, H, S, d, q, k, v, ld_qkv, o, ld_o, lse, scale, st)
                              : fwd_t<float, VG_ATTN_DOT>(B, H, S, d, q, k, v, ld_qkv, o, ld_o, lse, scale, st);
  return mode == VG_ATTN_L2 ? fwd_t<bf16, VG_ATTN_L2>(B, H, S, d, q, k, v, ld_qkv, o, ld_o, lse, scale, st)
                            : fwd_t<bf16, VG_ATTN_DOT>(B, H, S, d, q, k, v, ld_qkv, o, ld_o, lse, scale, st);
}

extern "C" int vg_attention_bwd(int dtype, int mode, int B, int H, int S, int d, const void* q, const void* k,
                                const void* v, int64_t ld_qkv, const void* o, const void* d_o, int64_t ld_o,
                                const float* lse, void* dq, void* dk, void* dv, int64_t ld_dqkv, float scale,
                                float* delta_ws, void* stream) {
  int rc = check_shape(B, H, S, d, ld_qkv, ld_o);
  if (rc) return rc;
  VG_REQUIRE(ld_dqkv % 4 == 0, VG_ERR_ALIGN, "attention_bwd: ld_dqkv must be a multiple of 4");
  cudaStream_t st = as_stream(stream);
  if (attention_tc_supported(dtype, mode, B, H, S, d, q, k, v, ld_qkv, d_o, ld_o) && ld_dqkv % 8 == 0 &&
      ((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) & 15) == 0)
    return attention_bwd_tc(B, H, S, d, q, k, v, ld_qkv, d_o, ld_o, lse, dq, dk, dv, ld_dqkv, scale, st);
  if (attention_mt_supported(dtype, mode, B, H, S, d, q, k, v, ld_qkv, d_o, ld_o) && ld_dqkv % 8 == 0 &&
      ((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv) | reinterpret_cast<uintptr_t>(o)) & 15) == 0)
    return attention_bwd_mt(mode, B, H, S, d, q, k, v, ld_qkv, o, d_o, ld_o, lse, dq, dk, dv, ld_dqkv, scale, delta_ws, st);
  if (dtype == VG_F32)
    return mode == VG_ATTN_L2
               ? bwd_t<float, VG_ATTN_L2>(B, H, S, d, q, k, v, ld_qkv, o, d_o, ld_o, lse, dq, dk, dv, ld_dqkv, scale, delta_ws, st)
               : bwd_t<float, VG_ATTN_DOT>(B, H, S, d, q, k, v, ld_qkv, o, d_o, ld_o, lse, dq, dk, dv, ld_dqkv, scale, delta_ws, st);
  return mode == VG_ATTN_L2
             ? bwd_t<bf16, VG_ATTN_L2>(B, H, S, d, q, k, v, ld_qkv, o, d_o, ld_o, lse, dq, dk, dv, ld_dqkv, scale, delta_ws, st)
             : bwd_t<bf16, VG_ATTN_DOT>(B, H, S, d, q, k, v, ld_qkv, o, d_o, ld_o, lse, dq, dk, dv, ld_dqkv, scale, delta_ws, st);
}

/* Which kernel family vg_attention_fwd / _bwd take for this problem (pointers and pitches assumed 16-byte friendly):
 * 0 = CUDA-core flash kernel, 1 = single-tile tcgen05 (attention_tc.cu), 2 = multi-tile tcgen05 (attention_mt.cu). */
extern "C" int vg_attention_path(int dtype, int mode, int B, int H, int S, int d) {
  const void* al = reinterpret_cast<const void*>(uintptr_t(256));
  const int64_t ld = (int64_t)3 * H * d, ldo = (int64_t)H * d;
  if (ld % 8 == 0 && attention_tc_supported(dtype, mode, B, H, S, d, al, al, al, ld, al, ldo)) return 1;
  if (ld % 8 == 0 && attention_mt_supported(dtype, mode, B, H, S, d, al, al, al, ld, al, ldo)) return 2;
  return 0;
}
