// gemm_simt.cu -- CUDA-core (FFMA, fp32 accumulate) GEMM with the full vg_gemm epilogue.
//
// Role: (1) the fp32 PARITY path (tcgen05 has no true-fp32 MMA) -- per-block
// outputs/gradients within 1e-4 of the CPU reference; (2) shapes the tensor-core kernel rejects
// (N = 10 classifier head, N = 1 D head, unaligned leading dimensions).  The bf16 fast path runs
// gemm_tc.cu instead.  Replaces ATen addmm/mm/bmm under F.linear and its backward.
#include "common.cuh"

namespace vg {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4, NTHREADS = 256;

template <typename TC, int ACT>
__device__ __noinline__ void simt_epilogue(const vg_gemm_args& g, const float (&acc)[TM][TN], int mbase, int nbase) {
  TC* __restrict__ C = static_cast<TC*>(g.C);
  const TC* __restrict__ aux = static_cast<const TC*>(g.aux);
  const TC* __restrict__ res = static_cast<const TC*>(g.residual);
  TC* __restrict__ cpre = static_cast<TC*>(g.c_pre);
  const bool first_split = blockIdx.z == 0;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = mbase + i;
    if (m >= g.M) continue;
    const int64_t orow = out_row(m, g.c_row_group);
    const int64_t rrow = res_row(m, g.res_row_mod, g.res_row_off);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = nbase + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.accumulate) {   // split-K partial sums: linear epilogue only (bias/residual added by split 0)
        if (first_split) {
          if (g.bias) v += g.bias[n];
          if (res) v += to_f<TC>(res[rrow * g.ldres + n]);
        }
        atomicAdd(reinterpret_cast<float*>(C) + orow * g.ldc + n, v);
      } else {
        if (g.bias) v += g.bias[n];
        if (cpre) cpre[(int64_t)m * g.ldpre + n] = from_f<TC>(v);
        const float a = (aux != nullptr) ? to_f<TC>(aux[(int64_t)m * g.ldaux + n]) : 0.f;
        v = act_t<ACT, false>(v, a, g.act_param);
        if (res) v += to_f<TC>(res[rrow * g.ldres + n]);
        C[orow * g.ldc + n] = from_f<TC>(v);
      }
    }
  }
}

template <typename TAB, typename TC>
__global__ void __launch_bounds__(NTHREADS)
gemm_simt_kernel(vg_gemm_args g, int k_per_split) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];

  const TAB* __restrict__ A = static_cast<const TAB*>(g.A);
  const TAB* __restrict__ B = static_cast<const TAB*>(g.B);
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;   // M tiles on x: neighbours share the weight tile in L2
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(g.K, k_begin + k_per_split);

  // element strides of the logical A(m,k), B(k,n)
  const int64_t a_sm = g.trans_a ? 1 : g.lda, a_sk = g.trans_a ? g.lda : 1;
  const int64_t b_sk = g.trans_b ? 1 : g.ldb, b_sn = g.trans_b ? g.ldb : 1;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    // ---- stage A tile (BM x BK) and B tile (BK x BN) as fp32; thread->element map follows the contiguous dim
#pragma unroll
    for (int i = 0; i < (BM * BK) / NTHREADS; ++i) {
      const int idx = tid + i * NTHREADS;
      int m, k;
      if (g.trans_a) { m = idx % BM; k = idx / BM; } else { k = idx % BK; m = idx / BK; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < g.M && gk < k_end) v = to_f<TAB>(A[gm * a_sm + gk * a_sk]);
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / NTHREADS; ++i) {
      const int idx = tid + i * NTHREADS;
      int n, k;
      if (g.trans_b) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < g.N && gk < k_end) v = to_f<TAB>(B[gk * b_sk + gn * b_sn]);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * TN]);
      a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
      b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue (activation dispatched once per thread, element loops are branch-free)
  VG_ACT_SWITCH(g.act, (simt_epilogue<TC, ACT>(g, acc, m0 + ty * TM, n0 + tx * TN)))
}

}  // namespace

int gemm_simt_launch(const vg_gemm_args& a, cudaStream_t st) {
  VG_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, VG_ERR_SHAPE, "vg_gemm: empty problem M=%d N=%d K=%d", a.M, a.N, a.K);
  VG_REQUIRE(!(a.accumulate && a.c_dtype != VG_F32), VG_ERR_ARG, "vg_gemm: accumulate needs fp32 C");
  VG_REQUIRE(!(a.accumulate && (a.act != VG_ACT_NONE || a.c_pre)), VG_ERR_ARG,
             "vg_gemm: accumulate supports a linear epilogue only");
  VG_REQUIRE(!act_needs_aux(a.act) || a.aux, VG_ERR_ARG, "vg_gemm: activation %d needs aux", a.act);
  int splits = 1;
  const int tiles = ((a.M + BM - 1) / BM) * ((a.N + BN - 1) / BN);
  if (a.accumulate) {   // weight gradients: few output tiles, very long K -> split K over the SMs
    const int want = (2 * num_sms() + tiles - 1) / tiles;
    const int max_splits = (a.K + 4 * BK - 1) / (4 * BK);
    splits = max(1, min(want, max_splits));
  }
  int k_per_split = (a.K + splits - 1) / splits;
  k_per_split = ((k_per_split + BK - 1) / BK) * BK;
  splits = (a.K + k_per_split - 1) / k_per_split;
  dim3 grid((a.M + BM - 1) / BM, (a.N + BN - 1) / BN, splits);
  VG_REQUIRE(grid.y <= 65535 && grid.z <= 65535, VG_ERR_SHAPE, "vg_gemm(simt): grid too large");
  if (a.ab_dtype == VG_F32 && a.c_dtype == VG_F32)
    gemm_simt_kernel<float, float><<<grid, NTHREADS, 0, st>>>(a, k_per_split);
  else if (a.ab_dtype == VG_BF16 && a.c_dtype == VG_BF16)
    gemm_simt_kernel<bf16, bf16><<<grid, NTHREADS, 0, st>>>(a, k_per_split);
  else if (a.ab_dtype == VG_BF16 && a.c_dtype == VG_F32)
    gemm_simt_kernel<bf16, float><<<grid, NTHREADS, 0, st>>>(a, k_per_split);
  else if (a.ab_dtype == VG_F32 && a.c_dtype == VG_BF16)
    gemm_simt_kernel<float, bf16><<<grid, NTHREADS, 0, st>>>(a, k_per_split);
  else
    VG_REQUIRE(false, VG_ERR_ARG, "vg_gemm: bad dtypes %d/%d", a.ab_dtype, a.c_dtype);
  return check_launch("gemm_simt");
}

}  // namespace vg
