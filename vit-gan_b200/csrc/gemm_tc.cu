// gemm_tc.cu -- bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM),
// operands streamed by TMA (cp.async.bulk.tensor, 128B swizzle) through an mbarrier ring, persistent
// warp-specialised CTAs.  Hand-written PTX; no CUTLASS/cuBLAS.
//
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocator + MMA issuer (one elected lane issues tcgen05.mma / tcgen05.commit)
//   warps 2..9  : epilogue: tcgen05.ld accumulator -> bias / activation / residual -> global stores
//
// Tile 128 x 128 x 64 (UMMA 128x128x16, 4 per k-block), NSTAGES-deep smem ring, 2 TMEM accumulator
// stages (256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
// All three products of a Linear layer run here without transposed copies in HBM:
//   forward  C = A W^T   : A K-major , B K-major         (trans_a=0, trans_b=1)
//   dgrad    dX = dY W   : A K-major , B MN-major        (trans_a=0, trans_b=0)
//   wgrad    dW = dY^T X : A MN-major, B MN-major, split-K + fp32 atomic accumulate (trans_a=1, trans_b=0)
// Replaces ATen addmm/mm under F.linear (src/v2/modules.py:128-139,161,173-175; src/v1/attention.py:46-48,102;
// muilti_layer_perceptron.py:39; siren.py:45) and their autograd backward.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace vg {
namespace {

constexpr int BM = 128, BN = 128, BK = 64, UK = 16;
constexpr int NSTAGES = 4;
constexpr int NACC = 2;
constexpr int EPI_WARPS = 8;
constexpr int NTHREADS = (2 + EPI_WARPS) * 32;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int STG_BYTES = 8192;         // per epilogue warp: two 4 KB swizzled staging tiles (32 rows x 128 B each)
constexpr int BIAS_BYTES = EPI_WARPS * 64 * 4;   // per epilogue warp: the 64 bias values of its column slab
constexpr int ONES_BYTES = 2048;        // [16 n x 64 k] bf16 tile of 1.0: B operand of the row-sum MMA (layout-agnostic)
constexpr int LN_BYTES = EPI_WARPS * 128 * 4 + 2 * EPI_WARPS * 32 * 4;   // fused LayerNorm: per warp gamma|beta of its 64 columns + 2 exchange slots
constexpr int SMEM_BYTES = NSTAGES * STAGE_BYTES + EPI_WARPS * STG_BYTES + BIAS_BYTES + ONES_BYTES + LN_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
constexpr int TMEM_COLS = NACC * BN;   // 256: power of two >= 32
constexpr int PAIR_STAGES = 5;          // CTA-pair kernel: operand ring depth (32 KB per stage per CTA)
constexpr int RS_COL = TMEM_COLS, RS_N = 16;   // row-sum accumulators (a_rowsum): 16 columns per stage behind the tile accumulators

struct Params {
  int M, N, K;
  int trans_a, trans_b;            // see file header
  int m_tiles, n_tiles, splits, kb_per_split, kb_total;
  int c_is_f32;
  void* C; int64_t ldc;
  const float* bias;
  int act; float act_param;
  const void* aux; int64_t ldaux;
  const void* residual; int64_t ldres;
  void* c_pre; int64_t ldpre;
  int c_row_group, res_row_mod, res_row_off;
  int accumulate;
  uint32_t mn_lbo, mn_sbo, mn_kstep;   // MN-major smem descriptor geometry (debug-overridable, see gemm_tc_launch)
  int dbg;                             // bring-up switches (VG_TC_DBG): 1 = epilogue skips global memory, 2 = force direct epilogue
  int epi_tma;                         // 1: smem-staged epilogue with TMA loads (residual/aux) and TMA stores / reduce-add
  int na_stages;                       // weight-stationary kernel: depth of the A k-block ring
  float* rowsum;                       // a_rowsum: rowsum[m] += sum_k opA(A)[m,k] (bias gradient of a wgrad GEMM) or NULL
  const float* ln_gamma; const float* ln_beta; float* ln_mean; float* ln_rstd; float ln_eps;   // fused LayerNorm of the output rows (N == 128) or NULL
  unsigned long long* trace;           // per-CTA timeline buffer or NULL
};
// per-warp scratch of the fused LayerNorm epilogue
// optional per-CTA timeline (vg_gemm_set_trace): trace[blockIdx.x * 16 + k] = %globaltimer at milestone k (see TR_* below)
enum { TR_ENTRY = 0, TR_PROLOGUE, TR_PDL, TR_TMA0, TR_FULL0, TR_MMA0, TR_TFULL0, TR_STORE0, TR_DRAINED, TR_EXIT, TR_TILES };
// Compiled in only with -DVG_TC_TRACE (make TRACE=1): even a never-taken branch per milestone cost 0.2-0.3 us per skinny GEMM.
__device__ __forceinline__ void trace_mark(unsigned long long* tr, int k) {
#ifdef VG_TC_TRACE
  if (tr != nullptr) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); tr[blockIdx.x * 16 + k] = t; }
#endif
}
struct LnScratch { float* gb; float* slots; };   // gb: gamma[64] | beta[64] of this warp's columns; slots: [2][EPI_WARPS][32] row partials

template <typename TC>
__device__ __forceinline__ void epilogue_chunk(const Params& p, const uint32_t (&r)[32], int m, int n0) {
  // one thread = one output row m, 32 consecutive columns n0..n0+31
  const int64_t orow = out_row(m, p.c_row_group);
  const int64_t rrow = res_row(m, p.res_row_mod, p.res_row_off);
  TC* C = static_cast<TC*>(p.C) + orow * p.ldc + n0;
  const TC* aux = p.aux ? static_cast<const TC*>(p.aux) + (int64_t)m * p.ldaux + n0 : nullptr;
  const TC* res = p.residual ? static_cast<const TC*>(p.residual) + rrow * p.ldres + n0 : nullptr;
  TC* cpre = p.c_pre ? static_cast<TC*>(p.c_pre) + (int64_t)m * p.ldpre + n0 : nullptr;
  const int nvalid = min(32, p.N - n0);
  constexpr int ALIGN_ELEMS = 16 / sizeof(TC) * 1;   // elements per 16 B
  const bool vec_ok = nvalid == 32 && (p.ldc % ALIGN_ELEMS == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                      (!aux || (p.ldaux % ALIGN_ELEMS == 0 && (reinterpret_cast<uintptr_t>(p.aux) & 15) == 0)) &&
                      (!res || (p.ldres % ALIGN_ELEMS == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0)) &&
                      (!cpre || (p.ldpre % ALIGN_ELEMS == 0 && (reinterpret_cast<uintptr_t>(p.c_pre) & 15) == 0));
#pragma unroll
  for (int g = 0; g < 8; ++g) {       // 8 groups of 4 columns
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = __uint_as_float(r[g * 4 + j]);
    const int c = g * 4;
    if (p.accumulate) {               // split-K partial: fp32 atomics; bias/residual contributed by split 0 only (folded by caller)
      float* Cf = reinterpret_cast<float*>(C);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < nvalid) atomicAdd(Cf + c + j, v[j]);
      continue;
    }
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < nvalid) v[j] += __ldg(p.bias + n0 + c + j);
    }
    if (vec_ok) {
      if (cpre) Vec4<TC>::store(cpre + c, v);
      if (p.act != VG_ACT_NONE) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        if (aux) Vec4<TC>::load(aux + c, a);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = apply_act(p.act, v[j], a[j], p.act_param);
      }
      if (res) {
        float t[4];
        Vec4<TC>::load(res + c, t);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += t[j];
      }
      Vec4<TC>::store(C + c, v);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (c + j < nvalid) {
          float x = v[j];
          if (cpre) cpre[c + j] = from_f<TC>(x);
          const float a = aux ? to_f<TC>(aux[c + j]) : 0.f;
          x = apply_act(p.act, x, a, p.act_param);
          if (res) x += to_f<TC>(res[c + j]);
          C[c + j] = from_f<TC>(x);
        }
      }
    }
  }
}

// Staged epilogue, one 32-column chunk of one accumulator row per thread.
//   bf16: the warp's 64 columns live in ONE 4 KB tile (32 rows x 128 B); chunk cc covers 16 B chunks 4cc..4cc+3
//   fp32: chunk cc has its own 4 KB tile (32 rows x 32 fp32)
// Tiles use the TMA 128B swizzle: 16 B chunk L of row r is stored at r*128 + ((L ^ (r & 7)) * 16)  (conflict-free).
// bufC holds the residual tile on entry (if any) and the output on exit; bufX holds aux on entry or c_pre on exit.
template <bool F32, int ACT, bool KEEP = false>
__device__ __forceinline__ void staged_chunk(const Params& p, uint32_t (&r)[32], int lane, int cc, int n0,
                                             uint32_t bufC, uint32_t bufX, const float* __restrict__ bias_s) {
  // KEEP (bf16 only): the final, bf16-ROUNDED output values are written back into r[] as floats (fused LayerNorm input:
  // the statistics are then taken over exactly the values a separate LayerNorm kernel would read from memory)
  const uint32_t row_off = (uint32_t)lane * 128u, sw = (uint32_t)(lane & 7);
  const bool has_res = p.residual != nullptr, has_aux = p.aux != nullptr, has_pre = p.c_pre != nullptr;
  if (F32) {
    const uint32_t base = (cc == 0 ? bufC : bufX) + row_off;      // fp32: second chunk uses the second tile
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = __uint_as_float(r[g * 4 + j]);
      const uint32_t addr = base + (((uint32_t)g ^ sw) << 4);
      if (!p.accumulate) {
        if (p.bias) {      // bias_s: this warp's 64 bias values, staged in smem at the start of the tile (broadcast reads)
          const float4 bv = *reinterpret_cast<const float4*>(bias_s + cc * 32 + g * 4);
          v[0] += bv.x; v[1] += bv.y; v[2] += bv.z; v[3] += bv.w;
        }
        if (ACT != VG_ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = act_t<ACT, true>(v[j], 0.f, p.act_param);
        }
        if (has_res) {
          uint32_t a, b, c, d;
          lds128(addr, a, b, c, d);
          v[0] += __uint_as_float(a); v[1] += __uint_as_float(b); v[2] += __uint_as_float(c); v[3] += __uint_as_float(d);
        }
      }
      sts128(addr, __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
    }
  } else {
#pragma unroll
    for (int h = 0; h < 4; ++h) {        // 8 columns -> one 16 B chunk
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[h * 8 + j]);
      const uint32_t off = row_off + ((((uint32_t)(cc * 4 + h)) ^ sw) << 4);
      if (p.bias) {
        const float4 b0 = *reinterpret_cast<const float4*>(bias_s + cc * 32 + h * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(bias_s + cc * 32 + h * 8 + 4);
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (has_aux) {
        uint32_t x0, x1, x2, x3;
        lds128(bufX + off, x0, x1, x2, x3);
        a[0] = bf16_lo(x0); a[1] = bf16_hi(x0); a[2] = bf16_lo(x1); a[3] = bf16_hi(x1);
        a[4] = bf16_lo(x2); a[5] = bf16_hi(x2); a[6] = bf16_lo(x3); a[7] = bf16_hi(x3);
      }
      if (has_pre) sts128(bufX + off, pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      if (ACT != VG_ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = act_t<ACT, true>(v[j], a[j], p.act_param);
      }
      if (has_res) {
        uint32_t x0, x1, x2, x3;
        lds128(bufC + off, x0, x1, x2, x3);
        v[0] += bf16_lo(x0); v[1] += bf16_hi(x0); v[2] += bf16_lo(x1); v[3] += bf16_hi(x1);
        v[4] += bf16_lo(x2); v[5] += bf16_hi(x2); v[6] += bf16_lo(x3); v[7] += bf16_hi(x3);
      }
      const uint32_t q0 = pack_bf16(v[0], v[1]), q1 = pack_bf16(v[2], v[3]), q2 = pack_bf16(v[4], v[5]), q3 = pack_bf16(v[6], v[7]);
      sts128(bufC + off, q0, q1, q2, q3);
      if (KEEP) {
        r[h * 8 + 0] = q0 << 16; r[h * 8 + 1] = q0 & 0xFFFF0000u; r[h * 8 + 2] = q1 << 16; r[h * 8 + 3] = q1 & 0xFFFF0000u;
        r[h * 8 + 4] = q2 << 16; r[h * 8 + 5] = q2 & 0xFFFF0000u; r[h * 8 + 6] = q3 << 16; r[h * 8 + 7] = q3 & 0xFFFF0000u;
      }
    }
  }
}

// ---- register-first variants of the bf16 staged epilogue (output-only bufC): the math runs while the previous group's bulk
// store may still be reading the staging tile; the tile is touched only by put_chunk(), after the caller's wait.
// bias (+ pre-activation copy to bufX) of one 32-column chunk, in place in r[] as floats
__device__ __forceinline__ void bias_pre_chunk(const Params& p, uint32_t (&r)[32], int lane, int cc, uint32_t bufX, const float* __restrict__ bias_s,
                                               bool write_pre) {
  const uint32_t row_off = (uint32_t)lane * 128u, sw = (uint32_t)(lane & 7);
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[h * 8 + j]);
    if (p.bias) {
      const float4 b0 = *reinterpret_cast<const float4*>(bias_s + cc * 32 + h * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(bias_s + cc * 32 + h * 8 + 4);
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
#pragma unroll
      for (int j = 0; j < 8; ++j) r[h * 8 + j] = __float_as_uint(v[j]);
    }
    if (write_pre)
      sts128(bufX + row_off + ((((uint32_t)(cc * 4 + h)) ^ sw) << 4), pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}
// activation (aux from bufX) / residual (from bufX) of one 32-column chunk whose bias is already added -> 16 packed bf16 pairs
template <int ACT>
__device__ __forceinline__ void act_chunk_packed(const Params& p, const uint32_t (&r)[32], uint32_t (&q)[16], int lane, int cc, uint32_t bufX,
                                                 bool aux_in_x, bool res_in_x) {
  const uint32_t row_off = (uint32_t)lane * 128u, sw = (uint32_t)(lane & 7);
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    float v[8], a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[h * 8 + j]);
    uint32_t x0 = 0u, x1 = 0u, x2 = 0u, x3 = 0u;
    if (aux_in_x || res_in_x) lds128(bufX + row_off + ((((uint32_t)(cc * 4 + h)) ^ sw) << 4), x0, x1, x2, x3);
    if (aux_in_x) {
      a[0] = bf16_lo(x0); a[1] = bf16_hi(x0); a[2] = bf16_lo(x1); a[3] = bf16_hi(x1);
      a[4] = bf16_lo(x2); a[5] = bf16_hi(x2); a[6] = bf16_lo(x3); a[7] = bf16_hi(x3);
    }
    if (ACT != VG_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = act_t<ACT, true>(v[j], a[j], p.act_param);
    }
    if (res_in_x) {
      v[0] += bf16_lo(x0); v[1] += bf16_hi(x0); v[2] += bf16_lo(x1); v[3] += bf16_hi(x1);
      v[4] += bf16_lo(x2); v[5] += bf16_hi(x2); v[6] += bf16_lo(x3); v[7] += bf16_hi(x3);
    }
    q[h * 4 + 0] = pack_bf16(v[0], v[1]); q[h * 4 + 1] = pack_bf16(v[2], v[3]); q[h * 4 + 2] = pack_bf16(v[4], v[5]); q[h * 4 + 3] = pack_bf16(v[6], v[7]);
  }
}
__device__ __forceinline__ void put_chunk(const uint32_t (&q)[16], int lane, int cc, uint32_t bufC) {
  const uint32_t row_off = (uint32_t)lane * 128u, sw = (uint32_t)(lane & 7);
#pragma unroll
  for (int h = 0; h < 4; ++h)
    sts128(bufC + row_off + ((((uint32_t)(cc * 4 + h)) ^ sw) << 4), q[h * 4 + 0], q[h * 4 + 1], q[h * 4 + 2], q[h * 4 + 3]);
}

// Fused LayerNorm of the 128-wide output row (N == BN == 128): this thread holds 64 of the row's values (r0 | r1, already
// rounded to bf16), its partner warp (same TMEM lane quadrant, other column half) the other 64.  Two-pass statistics with one
// smem exchange + 64-thread named barrier per pass; the normalised row goes to bufX (-> TMA store through the c_pre map).
__device__ __forceinline__ void ln_epilogue(const Params& p, const uint32_t (&r0)[32], const uint32_t (&r1)[32], int lane, int ew, int m,
                                            uint32_t bufX, const LnScratch& ln) {
  const int quad = ew & 3, half = ew >> 2, partner = ew ^ 4;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += __uint_as_float(r0[i]) + __uint_as_float(r1[i]);
  ln.slots[ew * 32 + lane] = s;
  asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
  const float mean = (s + ln.slots[partner * 32 + lane]) * (1.0f / 128.0f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float a = __uint_as_float(r0[i]) - mean, b = __uint_as_float(r1[i]) - mean;
    q = fmaf(a, a, q); q = fmaf(b, b, q);
  }
  ln.slots[EPI_WARPS * 32 + ew * 32 + lane] = q;
  asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
  const float rstd = rsqrtf((q + ln.slots[EPI_WARPS * 32 + partner * 32 + lane]) * (1.0f / 128.0f) + p.ln_eps);
  const uint32_t row_off = (uint32_t)lane * 128u, sw = (uint32_t)(lane & 7);
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float y[8];
      const int c0 = cc * 32 + h * 8;       // gamma | beta of these 8 columns: four 16-byte broadcast loads
      const float4 g0 = *reinterpret_cast<const float4*>(ln.gb + c0), g1 = *reinterpret_cast<const float4*>(ln.gb + c0 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(ln.gb + 64 + c0), b1 = *reinterpret_cast<const float4*>(ln.gb + 64 + c0 + 4);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = __uint_as_float(cc == 0 ? r0[h * 8 + j] : r1[h * 8 + j]);
        y[j] = fmaf((v - mean) * rstd, gg[j], bb[j]);
      }
      sts128(bufX + row_off + ((((uint32_t)(cc * 4 + h)) ^ sw) << 4), pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
    }
  }
  if (half == 0 && m < p.M) { p.ln_mean[m] = mean; p.ln_rstd[m] = rstd; }
}

// One epilogue warp's share ([32 rows x 64 columns]) of one accumulator tile, staged (TMA) epilogue:
//   wait for the staging tiles -> stage bias -> prefetch residual/aux by TMA -> wait accumulator -> TMEM -> registers ->
//   release TMEM -> math -> swizzled smem -> TMA store / reduce-add.
template <int MODE, int ACT, bool LN = false, bool REMAP = false>
__device__ __forceinline__ void staged_tile(const Params& p, const CUtensorMap* tmap_c, const CUtensorMap* tmap_pre,
                                            const CUtensorMap* tmap_res, const CUtensorMap* tmap_aux, uint32_t taddr,
                                            uint32_t tfull, uint32_t tfull_phase, uint32_t tempty, int m0, int n0, uint32_t bufC,
                                            uint32_t bufX, uint32_t wbar, uint32_t& wphase, float* bias_s, int lane,
                                            uint32_t rs_taddr = 0u, bool release = true, int ew = 0, LnScratch ln = LnScratch{nullptr, nullptr},
                                            bool alt = false, bool* aux_ready = nullptr, bool has_next = false, int next_m0 = 0, int next_n0 = 0,
                                            uint32_t tempty_cluster = 0u) {
  // tempty_cluster != 0 (CTA-pair kernel): the accumulator-drained arrival goes to the pair leader's barrier (shared::cluster address)
  // aux_ready / has_next / next_*: aux-only epilogues (dgrad + activation derivative) prefetch the NEXT group's aux tile as soon
  // as this group's math has consumed bufX -- the tile is never stored from, so it need not wait for this group's bulk store;
  // otherwise the ~1 us TMA load of every group sat exposed between the accumulator being ready and the math (fc2 dgrad+GELU'
  // at the C4 shape: 243 us vs 178 us for the plain GEMM of the same FLOPs)
  // alt (bf16 epilogues without residual / aux / c_pre / LayerNorm only): consecutive calls alternate between the warp's two
  // 4 KB staging tiles, so this tile is staged while the previous tile's bulk store is still reading the other one
  // release == false: more 64-column groups of the same accumulator follow (256-wide tiles); the last group frees TMEM
  // rs_taddr != 0: this warp also drains one row-sum column (a_rowsum) of its 32 rows and reduce-adds it into p.rowsum
  constexpr bool f32 = MODE == 2;
  const bool has_res = p.residual != nullptr, has_aux = p.aux != nullptr, has_pre = p.c_pre != nullptr;
  // staging tiles are free once the previous tile's bulk stores have READ them
  const bool dbl = MODE == 1 && !LN && !has_res && !has_aux && !has_pre;
  // register-first flows (bf16, bufC output-only): `pre2` = bias -> pre-activation tile stored as its own bulk group, activation
  // computed into registers while that store and the previous group's C store drain; `late` = side-tile chain (aux or residual in
  // bufX, never stored from): the wait for bufC moves from the top of the group to just before the tile is written.
  // Measured (profiles/bench_gemm_big.py, C4 shapes): the side-tile prefetch took fc2 dgrad x GELU' from 243 to 184 us and
  // fc2 + residual from 144 to 132 us; moving the store waits off the critical path changed fc1 + GELU (+pre) by < 1 % (167.6 ->
  // 166.2 us) -- at 128 KB of stores per 128 x 256 tile on top of 384 KB of operand fill that epilogue is bound by the SM's
  // memory interface, not by the wait.
  const bool side_chain = aux_ready != nullptr && (has_aux != has_res) && !has_pre && MODE == 1 && !(LN && MODE == 1) && !REMAP && !(p.dbg & 1);
  const bool pre2 = MODE == 1 && !LN && !REMAP && has_pre && !has_res && !has_aux && !p.accumulate;
  const bool late = side_chain && !p.accumulate;
  if (dbl && alt) { const uint32_t t = bufC; bufC = bufX; bufX = t; }
  float bv0 = 0.f, bv1 = 0.f;                 // this group's bias values: the loads are in flight while lane 0 waits below
  if (p.bias) {
    const int c0 = n0 + lane, c1 = n0 + 32 + lane;
    bv0 = c0 < p.N ? __ldg(p.bias + c0) : 0.f;
    bv1 = c1 < p.N ? __ldg(p.bias + c1) : 0.f;
  }
  if (lane == 0) {
    if (dbl || pre2) tma_wait_read1();        // pre2: the previous group's pre-activation store (its C store may still be pending)
    else if (!late) tma_wait_read();
  }
  __syncwarp();
  if (p.bias) {                               // stage this warp's 64 bias values (previous tile's readers are past them)
    bias_s[lane] = bv0;
    bias_s[32 + lane] = bv1;
    __syncwarp();
  }
  constexpr bool do_ln = LN && MODE == 1;     // compile-time: the LayerNorm code exists only in the one instantiation that needs it
  if (do_ln) {                                // ... and its 64 LayerNorm gammas / betas (N == 128: always in range)
    ln.gb[lane] = __ldg(p.ln_gamma + n0 + lane); ln.gb[32 + lane] = __ldg(p.ln_gamma + n0 + 32 + lane);
    ln.gb[64 + lane] = __ldg(p.ln_beta + n0 + lane); ln.gb[96 + lane] = __ldg(p.ln_beta + n0 + 32 + lane);
    __syncwarp();
  }
  // side-tile chain: exactly one side input (aux or residual), no second output -> it lives in bufX and is prefetched one group ahead
  const bool aux_chain = side_chain;
  const bool aux_here = !(aux_chain && *aux_ready);   // false: the previous group already issued this group's side-tile load
  if (aux_chain) {
    if (aux_here && lane == 0) {
      mbar_expect_tx(wbar, 4096u);
      tma_load_2d(bufX, has_aux ? tmap_aux : tmap_res, wbar, n0, m0);
    }
  } else if ((has_res || has_aux) && lane == 0) {    // prefetch residual / aux tiles while the MMAs run
    mbar_expect_tx(wbar, (has_res ? (f32 ? 8192u : 4096u) : 0u) + (has_aux ? 4096u : 0u));
    if (has_res) {
      // REMAP: the residual is a [mod (+off), N] table broadcast over the row groups (positional embedding)
      tma_load_2d(bufC, tmap_res, wbar, n0, REMAP ? (p.res_row_mod > 0 ? m0 % p.res_row_mod : m0) + p.res_row_off : m0);
      if (f32) tma_load_2d(bufX, tmap_res, wbar, n0 + 32, m0);
    }
    if (has_aux) tma_load_2d(bufX, tmap_aux, wbar, n0, m0);
  }
  mbar_wait(tfull, tfull_phase);
  tc_fence_after();
  uint32_t r0[32], r1[32];
  tmem_ld32_nowait(taddr, r0);
  tmem_ld32_nowait(taddr + 32u, r1);
  uint32_t rs = 0u;
  if (MODE == 2 && rs_taddr != 0u) rs = tmem_ld1_nowait(rs_taddr);
  tmem_ld_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0 && release) {                      // accumulator is in registers: release TMEM to the MMA warp
    if (tempty_cluster != 0u) mbar_arrive_cluster(tempty_cluster); else mbar_arrive(tempty);
  }
  if (MODE == 2 && rs_taddr != 0u) bias_s[lane] = __uint_as_float(rs);   // bias staging is idle in accumulate mode
  if (has_res || has_aux) { mbar_wait(wbar, wphase); wphase ^= 1u; }
  if (MODE == 1 && !do_ln && !REMAP && (pre2 || late) && !(p.dbg & 1)) {      // (late implies !(dbg & 1))
    bias_pre_chunk(p, r0, lane, 0, bufX, bias_s, pre2);
    bias_pre_chunk(p, r1, lane, 1, bufX, bias_s, pre2);
    const bool live = m0 < p.M && n0 < p.N && !(p.dbg & 4);
    if (pre2) {                                // the pre-activation tile leaves as its own bulk group
      fence_async_smem();
      __syncwarp();
      if (lane == 0 && live) { tma_store_2d(tmap_pre, bufX, n0, m0); tma_commit(); }
    }
    uint32_t q0[16], q1[16];
    act_chunk_packed<ACT>(p, r0, q0, lane, 0, bufX, late && has_aux, late && has_res);
    act_chunk_packed<ACT>(p, r1, q1, lane, 1, bufX, late && has_aux, late && has_res);
    if (late) {                                // bufX consumed by every lane: fetch the next group's side tile now
      fence_async_smem();
      __syncwarp();
      *aux_ready = has_next;
      if (has_next && lane == 0) {
        mbar_expect_tx(wbar, 4096u);
        tma_load_2d(bufX, has_aux ? tmap_aux : tmap_res, wbar, next_n0, next_m0);
      }
    }
    // bufC: the previous group's C store must have read it (pre2: all but the pre store just committed; late: everything)
    if (lane == 0) { if (pre2 && live) tma_wait_read1(); else tma_wait_read(); }
    __syncwarp();
    put_chunk(q0, lane, 0, bufC);
    put_chunk(q1, lane, 1, bufC);
    fence_async_smem();
    __syncwarp();
    if (lane == 0 && live) { tma_store_2d(tmap_c, bufC, n0, m0); tma_commit(); }
  } else if (!(p.dbg & 1)) {
    if (do_ln) {
      staged_chunk<f32, ACT, do_ln>(p, r0, lane, 0, n0, bufC, bufX, bias_s);
      staged_chunk<f32, ACT, do_ln>(p, r1, lane, 1, n0 + 32, bufC, bufX, bias_s);
      ln_epilogue(p, r0, r1, lane, ew, m0 + lane, bufX, ln);
    } else {
      staged_chunk<f32, ACT>(p, r0, lane, 0, n0, bufC, bufX, bias_s);      // (the side-tile chain always takes the register-first flow above)
      staged_chunk<f32, ACT>(p, r1, lane, 1, n0 + 32, bufC, bufX, bias_s);
    }
    fence_async_smem();                       // generic-proxy smem writes -> visible to the async (TMA) proxy
    __syncwarp();
    if (lane == 0 && m0 < p.M && n0 < p.N && !(p.dbg & 4)) {
      if (p.accumulate) {
        tma_reduce_add_2d(tmap_c, bufC, n0, m0);
        if (n0 + 32 < p.N) tma_reduce_add_2d(tmap_c, bufX, n0 + 32, m0);
        if (MODE == 2 && rs_taddr != 0u)
          bulk_reduce_add_f32(p.rowsum + m0, smem_u32(bias_s), (uint32_t)min(32, p.M - m0) * 4u);
      } else if (f32) {
        tma_store_2d(tmap_c, bufC, n0, m0);
        if (n0 + 32 < p.N) tma_store_2d(tmap_c, bufX, n0 + 32, m0);
      } else if (REMAP) {
        // C is [M / G, G + 1, N]: row m of the product lands in slot 1 + m % G of group m / G (slot 0 = CLS token, left alone);
        // G % 32 == 0, so this warp's 32 rows stay inside one group
        tma_store_3d(tmap_c, bufC, n0, 1 + m0 % p.c_row_group, m0 / p.c_row_group);
      } else {
        tma_store_2d(tmap_c, bufC, n0, m0);
        if (has_pre || do_ln) tma_store_2d(tmap_pre, bufX, n0, m0);     // c_pre, or the fused LayerNorm output (same map slot)
      }
      tma_commit();
    }
  }
}

// MODE 0: direct epilogue (row remaps / unaligned outputs; activation switched at run time)
// MODE 1: staged TMA epilogue, bf16 output, activation ACT fixed at compile time
// MODE 2: staged TMA epilogue, fp32 output (plain or split-K reduce-add), no activation
// BNT: tile width.  128 (4-stage ring) for the skinny / memory-bound shapes; 256 (3 stages, both accumulator stages = all 512
// TMEM columns) for the compute-bound ones: a 128x128 tile reads 32 KB of smem per 2.1 MFLOP (256 clk of smem bandwidth for 256
// clk of tensor pipe: smem-bound), a 128x256 tile 48 KB per 4.2 MFLOP (384 vs 512 clk).
template <int MODE, int ACT, int BNT, bool LN = false, bool REMAP = false>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_pre,
               const __grid_constant__ CUtensorMap tmap_res, const __grid_constant__ CUtensorMap tmap_aux, const Params p) {
  constexpr int BN = BNT, NSTAGES = BNT == 256 ? 3 : 4;                   // shadow the file-scope 128-wide constants
  constexpr int B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_dyn[];
  // SWIZZLE_128B tiles need 1024 B alignment
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t stg_base = smem_base + NSTAGES * STAGE_BYTES;          // 1024-aligned (stage bytes are multiples of 1024)
  const uint32_t bias_base = stg_base + EPI_WARPS * STG_BYTES;
  const uint32_t ones_base = bias_base + BIAS_BYTES;                    // 1024-aligned (all regions above are multiples of 1024)
  const uint32_t ln_base = ones_base + ONES_BYTES;
  const uint32_t bar_base = ln_base + LN_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (NSTAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * NSTAGES + NACC + a); };
  auto warp_bar = [&](int w) { return bar_base + 8u * (2 * NSTAGES + 2 * NACC + w); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * NSTAGES + 2 * NACC + EPI_WARPS);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - smem_u32(smem_dyn)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) trace_mark(p.trace, TR_ENTRY);
  const bool do_rs = MODE == 2 && BNT == 128 && p.rowsum != nullptr;
  const uint32_t tmem_cols = (do_rs || BNT == 256) ? 512u : (uint32_t)TMEM_COLS;

  if (do_rs && warp == 2) {        // B operand of the row-sum MMA: bf16 ones (any swizzle / major reads ones)
#pragma unroll
    for (int i = 0; i < ONES_BYTES / 512; ++i) sts128(ones_base + (uint32_t)(i * 32 + lane) * 16u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_async_smem();
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    if (MODE != 0) {     // the epilogue's maps too: a cold descriptor fetch costs ~0.5-0.9 us on the first tile (profiles/trace_gemm.py)
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_c) : "memory");
      if (p.c_pre != nullptr || LN) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_pre) : "memory");
      if (p.residual != nullptr) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_res) : "memory");
      if (p.aux != nullptr) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_aux) : "memory");
    }
    for (int s = 0; s < NSTAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < NACC; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), EPI_WARPS); }
    for (int w = 0; w < EPI_WARPS; ++w) mbar_init(warp_bar(w), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // whole warp: allocate TMEM columns, publish base address through smem
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (threadIdx.x == 0) trace_mark(p.trace, TR_PROLOGUE);
  pdl_trigger();   // dependents may start their prologue; they wait for our completion before touching memory
  pdl_wait();      // everything above overlapped the previous kernel's tail; global memory is touched only below
  if (threadIdx.x == 0) trace_mark(p.trace, TR_PDL);

  const int total_work = p.m_tiles * p.n_tiles * p.splits;
  // smem tile geometry per operand
  //   K-major : [rows][64 k] 128 B per row                      -> SBO = 1024 (8 rows), k-step = 32 B
  //   MN-major: [mn/64 blocks][64 k rows][64 mn] 128 B per row  -> LBO = 64*128 = 8192 (next 64-mn block),
  //             SBO = 1024 (next 8 k rows), k-step (16 rows) = 2048 B
  const uint32_t a_lbo = p.trans_a ? p.mn_lbo : 16u, a_sbo = p.trans_a ? p.mn_sbo : 1024u, a_kstep = p.trans_a ? p.mn_kstep : 32u;
  const uint32_t b_lbo = p.trans_b ? 16u : p.mn_lbo, b_sbo = p.trans_b ? 1024u : p.mn_sbo, b_kstep = p.trans_b ? 32u : p.mn_kstep;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int n_blk = w % p.n_tiles, m_blk = (w / p.n_tiles) % p.m_tiles, split = w / (p.n_tiles * p.m_tiles);
        const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          if (!p.trans_a) {
            tma_load_2d(sa, &tmap_a, full_bar(stage), kb * BK, m_blk * BM);                 // box {64 k, 128 m}
          } else {
            tma_load_2d(sa, &tmap_a, full_bar(stage), m_blk * BM, kb * BK);                 // box {64 m, 64 k} x2
            tma_load_2d(sa + 8192, &tmap_a, full_bar(stage), m_blk * BM + 64, kb * BK);
          }
          if (p.trans_b) {
            tma_load_2d(sb, &tmap_b, full_bar(stage), kb * BK, n_blk * BN);                 // box {64 k, 128 n}
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)                                               // box {64 n, 64 k} x BN/64
              tma_load_2d(sb + j * 8192, &tmap_b, full_bar(stage), n_blk * BN + 64 * j, kb * BK);
          }
          if (++stage == NSTAGES) { stage = 0; phase ^= 1u; }
        }
        if (w == (int)blockIdx.x) trace_mark(p.trace, TR_TMA0);
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=BF16, majors, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((p.trans_a ? 1u : 0u) << 15) |
                             ((p.trans_b ? 0u : 1u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      // row-sum MMA (a_rowsum): same A, B = ones [16 n x 16 k] K-major -> every column of D2 is sum_k opA(A)[m, k]
      const uint32_t idesc_rs = (1u << 4) | (1u << 7) | (1u << 10) | ((p.trans_a ? 1u : 0u) << 15) | ((uint32_t)(RS_N >> 3) << 17) |
                                ((uint32_t)(BM >> 4) << 24);
      const uint64_t ones_desc = make_desc(ones_base, 16u, 1024u);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int split = w / (p.n_tiles * p.m_tiles);
        const bool rs_tile = do_rs && (w % p.n_tiles) == 0;
        const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);          // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);                 // TMA bytes have landed
          tc_fence_after();
          if (w == (int)blockIdx.x && kb == kb0) trace_mark(p.trace, TR_FULL0);
          const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t ad = make_desc(sa + k * a_kstep, a_lbo, a_sbo);
            const uint64_t bd = make_desc(sb + k * b_kstep, b_lbo, b_sbo);
            tc_mma(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            if (rs_tile) tc_mma(tmem_base + (uint32_t)(RS_COL + acc * RS_N), ad, ones_desc, idesc_rs, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit(empty_bar(stage));                       // frees the smem slot when these MMAs retire
          if (++stage == NSTAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(acc));                           // accumulator complete -> epilogue
        if (w == (int)blockIdx.x) trace_mark(p.trace, TR_MMA0);
        if (++acc == NACC) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ============================== epilogue ==============================
    const int ew = warp - 2;               // 0..7
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access (warp id % 4)
    const int half = ew >> 2;              // which 64-column half of the tile
    int acc = 0; uint32_t acc_phase = 0;
    if (MODE != 0) {
      const uint32_t bufC = stg_base + (uint32_t)ew * STG_BYTES, bufX = bufC + 4096u;
      const uint32_t wbar = warp_bar(ew);
      uint32_t wphase = 0;
      float* bias_s = reinterpret_cast<float*>(smem_dyn + (bias_base - smem_u32(smem_dyn))) + ew * 64;
      float* ln_f = reinterpret_cast<float*>(smem_dyn + (ln_base - smem_u32(smem_dyn)));
      const LnScratch ln = LN ? LnScratch{ln_f + ew * 128, ln_f + EPI_WARPS * 128} : LnScratch{nullptr, nullptr};
      int n_staged = 0;
      bool aux_ready = false;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int n_blk = w % p.n_tiles, m_blk = (w / p.n_tiles) % p.m_tiles;
        const int m0 = m_blk * BM + quad * 32;
        const int wn = w + (int)gridDim.x;             // this CTA's next tile (aux prefetch chain)
        const int nn_blk = wn % p.n_tiles, nm0 = ((wn / p.n_tiles) % p.m_tiles) * BM + quad * 32;
        const uint32_t rs_taddr = (do_rs && n_blk == 0 && half == 0 && m0 < p.M) ? tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(RS_COL + acc * RS_N) : 0u;
#pragma unroll
        for (int g = 0; g < BN / 128; ++g) {           // this warp's 64-column groups of the tile (one at BN = 128, two at 256)
          const int col = half * (BN / 2) + g * 64;
          staged_tile<MODE, ACT, LN, REMAP>(p, &tmap_c, &tmap_pre, &tmap_res, &tmap_aux, tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + col),
                                     tfull_bar(acc), acc_phase, tempty_bar(acc), m0, n_blk * BN + col, bufC, bufX, wbar, wphase, bias_s, lane, rs_taddr,
                                     g == BN / 128 - 1, ew, ln, (n_staged++ & 1) != 0, &aux_ready,
                                     g + 1 < BN / 128 || wn < total_work, g + 1 < BN / 128 ? m0 : nm0,
                                     g + 1 < BN / 128 ? n_blk * BN + col + 64 : nn_blk * BN + half * (BN / 2));
        }
        if (ew == 0 && lane == 0 && w == (int)blockIdx.x) trace_mark(p.trace, TR_STORE0);
        if (++acc == NACC) { acc = 0; acc_phase ^= 1u; }
      }
      if (lane == 0) tma_wait_read();   // smem may not be released while bulk stores still read it; visibility comes with grid completion
#ifdef VG_TC_TRACE
      if (ew == 0 && lane == 0) { trace_mark(p.trace, TR_DRAINED); if (p.trace) p.trace[blockIdx.x * 16 + TR_TILES] = (total_work - blockIdx.x + gridDim.x - 1) / gridDim.x; }
#endif
    } else {
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int n_blk = w % p.n_tiles, m_blk = (w / p.n_tiles) % p.m_tiles;
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const int m = m_blk * BM + quad * 32 + lane;
#pragma unroll 1
        for (int cc = 0; cc < BN / 2; cc += 32) {
          const int col = half * (BN / 2) + cc;
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + col), r);
          const int n0 = n_blk * BN + col;
          if (m < p.M && n0 < p.N && !(p.dbg & 1)) {
            if (p.c_is_f32) epilogue_chunk<float>(p, r, m, n0);
            else epilogue_chunk<bf16>(p, r, m, n0);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        if (++acc == NACC) { acc = 0; acc_phase ^= 1u; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace_mark(p.trace, TR_EXIT);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}


// ================================================================================================ CTA-pair variant
// 256 x 256 output tile per 2-CTA cluster (tcgen05.mma.cta_group::2, M = 256): CTA rank r of the pair owns rows [128 r, 128 r + 128)
// of the tile (its own TMEM lanes) and stages, per 64-deep k-block, its own 128 x 64 slab of A plus rows [128 r, 128 r + 128) of the
// B tile -- 32 KB per SM per 512 clk of tensor pipe instead of the 48 KB of the 1-CTA 128 x 256 tile, whose tensor pipe sat at
// 62 % because the operand fill (96 B/clk/SM) exceeds what L2 delivers to 148 SMs at once (profiles/r02b_ncu_full_c4_gemm.txt).
// Roles per CTA as in gemm_tc_kernel (warp 0 TMA, warp 1 MMA, 8 epilogue warps); only the leader's warp 1 issues MMAs.
//   full[s]   leader only: its producer's arrive.expect_tx of 2 x 32 KB, completed by both CTAs' TMA bytes
//   empty[s]  per CTA, signalled in both CTAs by one multicast tcgen05.commit
//   tfull[a]  per CTA, multicast commit;   tempty[a]  leader only, 2 x EPI_WARPS arrivals (the peer's arrive remotely)
template <int MODE, int ACT>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_pre,
                const __grid_constant__ CUtensorMap tmap_res, const __grid_constant__ CUtensorMap tmap_aux, const Params p) {
  // Measured at 8192^3 (profiles/bench_gemm_big.py): MMAs alone (operand loads disabled) 1718 TFLOP/s, with loads 1232 TFLOP/s at
  // 4 and at 5 stages alike -- the ring is deep enough; what remains is the chip's power cap (SM clock ~1.4-1.65 GHz under load).
  // The full barrier takes ONE arrival (the leader's arrive.expect_tx for both CTAs' bytes): a second, remote arrive.expect_tx
  // from the peer's producer per stage cost 1.7x (0.86 us per stage instead of 0.5 us).
  constexpr int BN = 256, NST = PAIR_STAGES;             // pair tile width; stages of A (16 KB) + B half (16 KB)
  constexpr int BH_BYTES = 128 * BK * 2, STAGE = A_BYTES + BH_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  const uint32_t smem_base = smem_u32(smem_dyn);         // no static shared memory in this kernel: the dynamic window starts 1024-aligned
  if (threadIdx.x == 0 && (smem_base & 1023u) != 0u) { printf("vg gemm_tc2: dynamic shared memory not 1024-byte aligned\n"); __trap(); }
  const uint32_t stg_base = smem_base + NST * STAGE;
  const uint32_t bias_base = stg_base + EPI_WARPS * STG_BYTES;
  const uint32_t bar_base = bias_base + BIAS_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (NST + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * NST + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * NST + NACC + a); };
  auto warp_bar = [&](int w) { return bar_base + 8u * (2 * NST + 2 * NACC + w); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * NST + 2 * NACC + EPI_WARPS);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - smem_u32(smem_dyn)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();               // 0 = leader
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_c) : "memory");
    if (p.c_pre != nullptr) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_pre) : "memory");
    if (p.residual != nullptr) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_res) : "memory");
    if (p.aux != nullptr) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_aux) : "memory");
    for (int s = 0; s < NST; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < NACC; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * EPI_WARPS); }
    for (int w = 0; w < EPI_WARPS; ++w) mbar_init(warp_bar(w), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // the same warp of BOTH CTAs: pair-wide TMEM allocation (all 512 columns: two 256-column accumulator stages)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncwarp();
  cluster_sync_all();            // both CTAs' barriers are initialised before any remote arrive / multicast commit / pair TMA load
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_trigger();
  pdl_wait();

  const int m_pairs = (p.M + 2 * BM - 1) / (2 * BM);
  const int total_work = m_pairs * p.n_tiles * p.splits;
  const int pair = (int)blockIdx.x >> 1, npairs = (int)gridDim.x >> 1;
  const uint32_t a_lbo = p.trans_a ? p.mn_lbo : 16u, a_sbo = p.trans_a ? p.mn_sbo : 1024u, a_kstep = p.trans_a ? p.mn_kstep : 32u;
  const uint32_t b_lbo = p.trans_b ? 16u : p.mn_lbo, b_sbo = p.trans_b ? 1024u : p.mn_sbo, b_kstep = p.trans_b ? 32u : p.mn_kstep;

  if (warp == 0) {
    // ============================== TMA producer (both CTAs) ==============================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int w = pair; w < total_work; w += npairs) {
        const int n_blk = w % p.n_tiles, m_pair = (w / p.n_tiles) % m_pairs, split = w / (p.n_tiles * m_pairs);
        const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int mrow = m_pair * 2 * BM + (int)rank * BM;          // this CTA's 128 rows of A (and of the output tile)
        const int nrow = n_blk * BN + (int)rank * 128;              // this CTA's 128 rows of the B tile
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * STAGE, sb = sa + A_BYTES;
          const uint32_t fb = mapa_u32(full_bar(stage), 0u);        // the leader's barrier counts both CTAs' bytes
          if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * STAGE);   // one local arrival; the peer's bytes may land before or after it
          if (!p.trans_a) {
            tma_load_2d_pair(sa, &tmap_a, fb, kb * BK, mrow);                       // box {64 k, 128 m}
          } else {
            tma_load_2d_pair(sa, &tmap_a, fb, mrow, kb * BK);                       // box {64 m, 64 k} x2
            tma_load_2d_pair(sa + 8192, &tmap_a, fb, mrow + 64, kb * BK);
          }
          if (p.trans_b) {
            tma_load_2d_pair(sb, &tmap_b, fb, kb * BK, nrow);                       // box {64 k, 128 n}
          } else {
            tma_load_2d_pair(sb, &tmap_b, fb, nrow, kb * BK);                       // box {64 n, 64 k} x2
            tma_load_2d_pair(sb + 8192, &tmap_b, fb, nrow + 64, kb * BK);
          }
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer (leader CTA only) ==============================
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((p.trans_a ? 1u : 0u) << 15) |
                             ((p.trans_b ? 0u : 1u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int w = pair; w < total_work; w += npairs) {
        const int split = w / (p.n_tiles * m_pairs);
        const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);          // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);                 // both CTAs' TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE, sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t ad = make_desc(sa + k * a_kstep, a_lbo, a_sbo);
            const uint64_t bd = make_desc(sb + k * b_kstep, b_lbo, b_sbo);
            tc_mma_pair(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit_pair(empty_bar(stage), 3);               // frees the stage in BOTH CTAs when these MMAs retire
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
        tc_commit_pair(tfull_bar(acc), 3);                   // accumulator complete -> both CTAs' epilogues
        if (++acc == NACC) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ============================== epilogue (both CTAs, own 128 rows x 256 columns) ==============================
    const int ew = warp - 2, quad = warp & 3, half = ew >> 2;
    int acc = 0; uint32_t acc_phase = 0;
    const uint32_t bufC = stg_base + (uint32_t)ew * STG_BYTES, bufX = bufC + 4096u;
    const uint32_t wbar = warp_bar(ew);
    uint32_t wphase = 0;
    float* bias_s = reinterpret_cast<float*>(smem_dyn + (bias_base - smem_u32(smem_dyn))) + ew * 64;
    int n_staged = 0;
    bool aux_ready = false;
    for (int w = pair; w < total_work; w += npairs) {
      const int n_blk = w % p.n_tiles, m_pair = (w / p.n_tiles) % m_pairs;
      const int m0 = m_pair * 2 * BM + (int)rank * BM + quad * 32;
      const int wn = w + npairs;
      const int nn_blk = wn % p.n_tiles, nm0 = ((wn / p.n_tiles) % m_pairs) * 2 * BM + (int)rank * BM + quad * 32;
      const uint32_t te = mapa_u32(tempty_bar(acc), 0u);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int col = half * 128 + g * 64;
        staged_tile<MODE, ACT, false, false>(p, &tmap_c, &tmap_pre, &tmap_res, &tmap_aux, tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + col),
                                             tfull_bar(acc), acc_phase, tempty_bar(acc), m0, n_blk * BN + col, bufC, bufX, wbar, wphase, bias_s, lane, 0u,
                                             g == 1, ew, LnScratch{nullptr, nullptr}, (n_staged++ & 1) != 0, &aux_ready,
                                             g == 0 || wn < total_work, g == 0 ? m0 : nm0, g == 0 ? n_blk * BN + col + 64 : nn_blk * BN + half * 128, te);
      }
      if (++acc == NACC) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) tma_wait_read();
  }

  tc_fence_before();
  __syncwarp();
  cluster_sync_all();            // neither CTA may free TMEM (or exit, taking its barriers and operand tiles away) while its peer still runs
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D row-major matrix [rows, cols] (leading dim ld elements, bf16 or fp32) -> tensor map with box {box_cols, box_rows}, 128B swizzle
int make_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows, bool f32 = false) {
  EncodeTiledFn enc = get_encode();
  VG_REQUIRE(enc != nullptr, VG_ERR_LAUNCH, "gemm_tc: cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (f32 ? 4 : 2)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VG_REQUIRE(r == CUDA_SUCCESS, VG_ERR_LAUNCH, "gemm_tc: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
             (long long)rows, (long long)cols, (long long)ld);
  return VG_OK;
}

// bf16 tensor [d2, d1, cols] with byte strides s1 (between d1 rows) and s2 (between d2 slabs); box {64 cols, 32 rows, 1}, 128B swizzle
int make_map3(CUtensorMap* map, const void* ptr, int64_t cols, int64_t d1, int64_t d2, int64_t s1, int64_t s2) {
  EncodeTiledFn enc = get_encode();
  VG_REQUIRE(enc != nullptr, VG_ERR_LAUNCH, "gemm_tc: cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)s1, (cuuint64_t)s2};
  cuuint32_t box[3] = {64, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VG_REQUIRE(r == CUDA_SUCCESS, VG_ERR_LAUNCH, "gemm_tc: cuTensorMapEncodeTiled(3d) failed (%d)", (int)r);
  return VG_OK;
}

}  // namespace

static unsigned long long* g_trace = nullptr;
void gemm_tc_set_trace(unsigned long long* p) { g_trace = p; }

bool gemm_tc_supported(const vg_gemm_args& a, const char** why) {
  static const int sm100 = vg_device_is_sm100();   // one-time, thread-safe
  if (!sm100) { *why = "device is not sm_100"; return false; }
  if (a.ab_dtype != VG_BF16) { *why = "A/B must be bf16"; return false; }
  if (a.M < 1 || a.N < 8 || a.K < 1) { *why = "degenerate shape (N < 8)"; return false; }
  if (a.lda % 8 || a.ldb % 8) { *why = "lda/ldb must be multiples of 8 elements (16 B TMA pitch)"; return false; }
  if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.B) & 15)) { *why = "A/B must be 16 B aligned"; return false; }
  if (a.accumulate && a.c_dtype != VG_F32) { *why = "accumulate needs fp32 C"; return false; }
  if (a.accumulate && (a.act != VG_ACT_NONE || a.c_pre || a.bias || a.residual)) { *why = "accumulate supports a plain epilogue only"; return false; }
  if (act_needs_aux(a.act) && !a.aux) { *why = "activation needs aux"; return false; }
  if (a.ln_gamma) {
    if (a.N != 128 || a.c_dtype != VG_BF16 || a.aux || a.c_pre || a.accumulate || a.c_row_group || a.res_row_mod) { *why = "fused LayerNorm needs N == 128, bf16 C, no aux / c_pre / accumulate / row remap"; return false; }
    if (!a.ln_beta || !a.ln_out || !a.ln_mean || !a.ln_rstd || (reinterpret_cast<uintptr_t>(a.ln_out) & 15) || (a.ld_ln * 2) % 16) { *why = "fused LayerNorm: missing or unaligned outputs"; return false; }
    if ((reinterpret_cast<uintptr_t>(a.C) & 15) || (a.ldc * 2) % 16 || (a.residual && ((reinterpret_cast<uintptr_t>(a.residual) & 15) || (a.ldres * 2) % 16))) { *why = "fused LayerNorm needs TMA-addressable C / residual"; return false; }
  }
  if (a.a_rowsum) {
    const bool c_tma = (reinterpret_cast<uintptr_t>(a.C) & 15) == 0 && (a.ldc * 4) % 16 == 0 && a.c_row_group == 0 && a.res_row_mod == 0;
    if (!a.accumulate || !c_tma) { *why = "a_rowsum needs the split-K accumulate mode with a TMA-addressable fp32 C"; return false; }
    if (a.M % 4 || (reinterpret_cast<uintptr_t>(a.a_rowsum) & 15)) { *why = "a_rowsum needs M % 4 == 0 and a 16 B aligned vector"; return false; }
  }
  return true;
}

// Bring-up knobs, read ONCE per process (thread-safe static initialisation; the launch path itself never calls getenv):
//   VG_TC_DBG   bit 0 = epilogue skips math and global memory, bit 1 = force the direct (per-thread) epilogue, bit 2 = epilogue math
//               and staging but no TMA stores (timing experiments), bit 4 = print the pair-cluster capacity
//   VG_TC_BN    128 | 256 overrides the tile-width heuristic
//   VG_TC_PAIR  0 keeps the wide tiles on the 1-CTA kernel (A/B comparison with the CTA-pair kernel)
//   VG_TC_MN_DESC "lbo,sbo,kstep" (bytes) overrides the MN-major descriptor strides
struct TcEnv {
  int dbg = 0, bn = 0, pair = 1;
  unsigned mn_lbo = 8192u, mn_sbo = 1024u, mn_kstep = 2048u;
  TcEnv() {
    if (const char* e = getenv("VG_TC_DBG")) dbg = atoi(e);
    if (const char* e = getenv("VG_TC_BN")) bn = atoi(e);
    if (const char* e = getenv("VG_TC_PAIR")) pair = atoi(e);
    if (const char* e = getenv("VG_TC_MN_DESC")) {
      unsigned l = 0, sb = 0, ks = 0;
      if (sscanf(e, "%u,%u,%u", &l, &sb, &ks) == 3) { mn_lbo = l; mn_sbo = sb; mn_kstep = ks; }
    }
  }
};
const TcEnv& tc_env() { static const TcEnv e; return e; }

// one-time, thread-safe opt-in to the kernel's dynamic shared memory size (C++11 static initialisation)
template <typename K>
cudaError_t set_smem_once(K kernel, int bytes) { return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); }

// 2-CTA cluster launch (+ programmatic dependent launch)
template <typename... KArgs, typename... Args>
cudaError_t launch_pair(void (*kernel)(KArgs...), int grid, int smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
// How many 2-CTA clusters of this kernel the device runs at once.  A pair needs both SMs of one TPC: on parts where yield
// harvesting leaves single-SM TPCs this is LESS than num_sms / 2, and a persistent grid sized beyond it would run its surplus
// clusters as a second wave (measured: 1.7x slower).  Queried once per kernel instantiation.
template <typename K>
int max_active_pairs(K kernel, int smem) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * num_sms()); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  return n;
}
int pair_slots();
constexpr int PAIR_SMEM = PAIR_STAGES * (A_BYTES + 128 * BK * 2) + EPI_WARPS * STG_BYTES + BIAS_BYTES + 512;

// every instantiation has the same footprint (one CTA per SM by shared memory), so one query serves them all
int pair_slots() {
  static const int n = [] {
    if (set_smem_once(gemm_tc2_kernel<1, 0>, PAIR_SMEM) != cudaSuccess) { cudaGetLastError(); return 0; }
    const int m = max_active_pairs(gemm_tc2_kernel<1, 0>, PAIR_SMEM);
    if (tc_env().dbg & 16) fprintf(stderr, "vitgan_b200: gemm_tc CTA-pair kernel: %d active 2-CTA clusters on %d SMs\n", m, num_sms());
    return m;
  }();
  return n;
}

int gemm_tc_launch(const vg_gemm_args& a, cudaStream_t st) {
  const TcEnv& env = tc_env();
  CUtensorMap ma, mb;
  int rc;
  // A: trans_a=0 stored [M,K] -> box {64 k, 128 m};  trans_a=1 stored [K,M] -> box {64 m, 64 k}
  rc = a.trans_a ? make_map(&ma, a.A, a.K, a.M, a.lda, 64, 64) : make_map(&ma, a.A, a.M, a.K, a.lda, 64, BM);
  if (rc) return rc;
  // B: trans_b=1 stored [N,K] -> box {64 k, 128 n};  trans_b=0 stored [K,N] -> box {64 n, 64 k}
  // tile width: 256 for compute-bound problems (deep K, N a multiple of 256, enough tiles to fill the machine) whose
  // epilogue can be staged; 128 otherwise.  VG_TC_BN=128|256 overrides the heuristic (256 still needs the staged epilogue).
  const int sms0 = num_sms();
  const int esz0 = a.c_dtype == VG_F32 ? 4 : 2;
  auto tma_ok0 = [&](const void* ptr, int64_t ld) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * esz0) % 16 == 0; };
  const bool stageable = !(env.dbg & 2) && a.c_row_group == 0 && a.res_row_mod == 0 && !a.a_rowsum && !a.ln_gamma && tma_ok0(a.C, a.ldc) &&
                         (!a.residual || tma_ok0(a.residual, a.ldres)) && (!a.aux || tma_ok0(a.aux, a.ldaux)) &&
                         (!a.c_pre || tma_ok0(a.c_pre, a.ldpre)) && !(a.c_dtype == VG_F32 && (a.aux || a.c_pre || a.act != VG_ACT_NONE));   // == epi_tma below
  const int bn_env = env.bn;
  bool wide = stageable && a.N % 256 == 0 && a.K >= 512 &&
              ((int64_t)((a.M + BM - 1) / BM) * (a.N / 256) >= sms0 ||
               (a.accumulate && a.K >= 4096 && (int64_t)((a.M + BM - 1) / BM) * (a.N / 256) >= 8));   // split-K supplies the parallelism
  if (bn_env == 128) wide = false;
  if (bn_env == 256) wide = stageable && a.N >= 256;
  const int bn = wide ? 256 : BN;
  // wide tiles run on CTA pairs (256 x 256 per 2-CTA cluster, each CTA staging half of the B tile) when M spans at least one pair tile
  const bool pair = wide && env.pair != 0 && a.M >= 2 * BM && pair_slots() >= 1;
  rc = a.trans_b ? make_map(&mb, a.B, a.N, a.K, a.ldb, 64, pair ? 128 : bn) : make_map(&mb, a.B, a.K, a.N, a.ldb, 64, 64);
  if (rc) return rc;

  Params p;
  p.M = a.M; p.N = a.N; p.K = a.K; p.trans_a = a.trans_a; p.trans_b = a.trans_b;
  p.m_tiles = (a.M + BM - 1) / BM; p.n_tiles = (a.N + bn - 1) / bn;
  p.kb_total = (a.K + BK - 1) / BK;
  p.splits = 1;
  const int sms = num_sms();
  if (a.accumulate) {
    // split-K factor: the smallest s (each split keeps >= 4 k-blocks) whose tiles*s work items fill >= 90 % of the
    // CTA rounds they need, else the best-filling one (e.g. 54 wide tiles: s = 8 -> 432 items = 2.92 rounds of 148)
    const int tiles = pair ? ((a.M + 2 * BM - 1) / (2 * BM)) * p.n_tiles : p.m_tiles * p.n_tiles;
    const int slots = pair ? pair_slots() : sms;         // concurrently resident tiles (CTA pairs or CTAs)
    const int smax = max(1, min(slots, (p.kb_total + 3) / 4));
    int best = 1; double best_eff = 0.0;
    for (int sp = 1; sp <= smax; ++sp) {
      const int items = tiles * sp, rounds = (items + slots - 1) / slots;
      const double eff = (double)items / ((double)rounds * slots);
      if (eff > best_eff + 1e-9) { best_eff = eff; best = sp; }
      if (eff >= 0.9) { best = sp; break; }
    }
    p.splits = best;
  }
  p.kb_per_split = (p.kb_total + p.splits - 1) / p.splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.c_is_f32 = a.c_dtype == VG_F32;
  p.C = a.C; p.ldc = a.ldc; p.bias = a.bias; p.act = a.act; p.act_param = a.act_param;
  p.aux = a.aux; p.ldaux = a.ldaux; p.residual = a.residual; p.ldres = a.ldres; p.c_pre = a.c_pre; p.ldpre = a.ldpre;
  p.c_row_group = a.c_row_group; p.res_row_mod = a.res_row_mod; p.res_row_off = a.res_row_off; p.accumulate = a.accumulate;
  p.mn_lbo = env.mn_lbo; p.mn_sbo = env.mn_sbo; p.mn_kstep = env.mn_kstep;
  p.dbg = env.dbg;
  p.rowsum = a.a_rowsum;
  p.trace = g_trace;
  p.ln_gamma = a.ln_gamma; p.ln_beta = a.ln_beta; p.ln_mean = a.ln_mean; p.ln_rstd = a.ln_rstd; p.ln_eps = a.ln_eps;
  // staged (TMA) epilogue whenever the output / side tensors are TMA-addressable and no row remap is requested
  const bool f32 = p.c_is_f32 != 0;
  const int esz = f32 ? 4 : 2;
  auto tma_ok = [&](const void* ptr, int64_t ld) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * esz) % 16 == 0; };
  p.epi_tma = !(p.dbg & 2) && a.c_row_group == 0 && a.res_row_mod == 0 && tma_ok(a.C, a.ldc) &&
              (!a.residual || tma_ok(a.residual, a.ldres)) && (!a.aux || tma_ok(a.aux, a.ldaux)) &&
              (!a.c_pre || tma_ok(a.c_pre, a.ldpre)) && !(f32 && (a.aux || a.c_pre || a.act != VG_ACT_NONE));
  // row-group remap (patch embedding: CLS slot + broadcast positional table) through 3-D TMA stores when the groups are
  // warp-tile aligned; otherwise the direct (MODE 0) epilogue
  const bool remap = !(p.dbg & 2) && a.c_row_group > 0 && a.c_row_group % 32 == 0 && a.M % a.c_row_group == 0 &&
                     (a.res_row_mod == 0 || a.res_row_mod == a.c_row_group) && !f32 && !a.aux && !a.c_pre && !a.ln_gamma && !a.accumulate &&
                     a.act == VG_ACT_NONE && tma_ok(a.C, a.ldc) && (!a.residual || tma_ok(a.residual, a.ldres));
  CUtensorMap mc = ma, mp = ma, mr = ma, mx = ma;     // placeholders when unused
  if (remap) {
    p.epi_tma = 1;
    const int G = a.c_row_group;
    if ((rc = make_map3(&mc, a.C, a.N, G + 1, a.M / G, a.ldc * 2, (int64_t)(G + 1) * a.ldc * 2))) return rc;
    if (a.residual && (rc = make_map(&mr, a.residual, (a.res_row_mod > 0 ? a.res_row_mod : a.M) + a.res_row_off, a.N, a.ldres, 64, 32, false))) return rc;
  } else if (p.epi_tma) {
    const int bc = f32 ? 32 : 64;                      // 128-byte wide boxes, 32 rows
    if ((rc = make_map(&mc, a.C, a.M, a.N, a.ldc, bc, 32, f32))) return rc;
    if (a.c_pre && (rc = make_map(&mp, a.c_pre, a.M, a.N, a.ldpre, bc, 32, f32))) return rc;
    if (a.ln_out && (rc = make_map(&mp, a.ln_out, a.M, a.N, a.ld_ln, bc, 32, f32))) return rc;
    if (a.residual && (rc = make_map(&mr, a.residual, a.M, a.N, a.ldres, bc, 32, f32))) return rc;
    if (a.aux && (rc = make_map(&mx, a.aux, a.M, a.N, a.ldaux, bc, 32, f32))) return rc;
  }
  const int total = p.m_tiles * p.n_tiles * p.splits;
  const int grid = min(total, sms);
  const int mode = !p.epi_tma ? 0 : (f32 ? 2 : 1);
  VG_REQUIRE(!a.a_rowsum || mode == 2, VG_ERR_UNSUPPORTED, "gemm_tc: a_rowsum needs the staged fp32 epilogue");
  VG_REQUIRE(!(bn == 256 && mode == 0), VG_ERR_LAUNCH, "gemm_tc: internal: 256-wide tile without a staged epilogue");
  VG_REQUIRE(!a.ln_gamma || (mode == 1 && bn == 128), VG_ERR_UNSUPPORTED, "gemm_tc: fused LayerNorm needs the staged bf16 epilogue");
  const bool wide_k = bn == 256;
  if (pair) {
    VG_REQUIRE(mode != 0 && !a.ln_gamma && !remap, VG_ERR_LAUNCH, "gemm_tc: internal: CTA-pair tile without a plain staged epilogue");
    const int units = ((a.M + 2 * BM - 1) / (2 * BM)) * p.n_tiles * p.splits;
#define VG_TC2_LAUNCH(MODE_, ACT_)                                                                                             \
  do {                                                                                                                         \
    static const cudaError_t attr_e = set_smem_once(gemm_tc2_kernel<MODE_, ACT_>, PAIR_SMEM);                                  \
    VG_REQUIRE(attr_e == cudaSuccess, VG_ERR_LAUNCH, "gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_e));         \
    const int grid2 = 2 * min(units, pair_slots());                                                                            \
    cudaError_t le = launch_pair(gemm_tc2_kernel<MODE_, ACT_>, grid2, PAIR_SMEM, st, ma, mb, mc, mp, mr, mx, p);               \
    VG_REQUIRE(le == cudaSuccess, VG_ERR_LAUNCH, "gemm_tc: CTA-pair launch: %s", cudaGetErrorString(le));                      \
  } while (0)
    if (mode == 2) VG_TC2_LAUNCH(2, 0);
    else { VG_ACT_SWITCH(a.act, VG_TC2_LAUNCH(1, ACT)) }
#undef VG_TC2_LAUNCH
    return check_launch("gemm_tc2");
  }
#define VG_TC_LAUNCH(MODE_, ACT_, BN_)                                                                                         \
  do {                                                                                                                         \
    constexpr int smem_ = (BN_ == 256 ? 3 * (A_BYTES + 256 * BK * 2) : NSTAGES * STAGE_BYTES) + EPI_WARPS * STG_BYTES + BIAS_BYTES + ONES_BYTES + LN_BYTES + 1024 + 512; \
    static const cudaError_t attr_e = set_smem_once(gemm_tc_kernel<MODE_, ACT_, BN_>, smem_);                                  \
    VG_REQUIRE(attr_e == cudaSuccess, VG_ERR_LAUNCH, "gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_e));         \
    launch_pdl(gemm_tc_kernel<MODE_, ACT_, BN_>, dim3(grid), dim3(NTHREADS), smem_, st, ma, mb, mc, mp, mr, mx, p);             \
  } while (0)
  if (remap) {
    constexpr int smem_ = NSTAGES * STAGE_BYTES + EPI_WARPS * STG_BYTES + BIAS_BYTES + ONES_BYTES + LN_BYTES + 1024 + 512;
    static const cudaError_t attr_e = set_smem_once(gemm_tc_kernel<1, VG_ACT_NONE, 128, false, true>, smem_);
    VG_REQUIRE(attr_e == cudaSuccess, VG_ERR_LAUNCH, "gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_e));
    launch_pdl(gemm_tc_kernel<1, VG_ACT_NONE, 128, false, true>, dim3(grid), dim3(NTHREADS), smem_, st, ma, mb, mc, mp, mr, mx, p);
    return check_launch("gemm_tc");
  }
  if (a.ln_gamma) {
    VG_REQUIRE(a.act == VG_ACT_NONE, VG_ERR_UNSUPPORTED, "gemm_tc: fused LayerNorm is built for the activation-free epilogue");
    constexpr int smem_ = NSTAGES * STAGE_BYTES + EPI_WARPS * STG_BYTES + BIAS_BYTES + ONES_BYTES + LN_BYTES + 1024 + 512;
    static const cudaError_t attr_e = set_smem_once(gemm_tc_kernel<1, VG_ACT_NONE, 128, true>, smem_);
    VG_REQUIRE(attr_e == cudaSuccess, VG_ERR_LAUNCH, "gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_e));
    launch_pdl(gemm_tc_kernel<1, VG_ACT_NONE, 128, true>, dim3(grid), dim3(NTHREADS), smem_, st, ma, mb, mc, mp, mr, mx, p);
    return check_launch("gemm_tc");
  }
  if (mode == 0) VG_TC_LAUNCH(0, 0, 128);
  else if (wide_k) {
    if (mode == 2) VG_TC_LAUNCH(2, 0, 256);
    else { VG_ACT_SWITCH(a.act, VG_TC_LAUNCH(1, ACT, 256)) }
  } else if (mode == 2) VG_TC_LAUNCH(2, 0, 128);
  else { VG_ACT_SWITCH(a.act, VG_TC_LAUNCH(1, ACT, 128)) }
#undef VG_TC_LAUNCH
  return check_launch("gemm_tc");
}

}  // namespace vg
