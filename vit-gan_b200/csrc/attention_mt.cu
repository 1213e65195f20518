// attention_mt.cu -- multi-tile flash attention on tcgen05 / TMEM / TMA for the head sizes the single-tile kernels of
// attention_tc.cu cannot take: d = 192 with S = 257 (the scaled v2 config, src/v2/modules.py:142-159), d = 96 (v1 generator,
// dot scores, src/v1/attention.py:69-70) and d = 112 (v1 discriminator head of 108 zero-padded to 112 by the caller,
// L2-distance scores, src/v1/attention.py:66-67).  bf16 operands, fp32 accumulation / statistics, S <= 272 keys.
//
// Three kernels, one CTA per SM, persistent over (batch, head) problems; warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM
// alloc), warps 2..5 = 128 compute threads (thread = TMEM lane = one row of the accumulator tile):
//   forward          item = 128-query tile.  S[128 x NK] = Q K^T is accumulated over 64-key blocks streamed through a TMA ring
//                    (NK <= 272 fp32 columns: the whole key range fits TMEM, so the softmax is exact two-pass -- no online
//                    rescaling); P goes to smem as bf16 in 64-key chunks and O += P_j V_j starts while later chunks are still
//                    being exponentiated; O (D columns) sits beside S in TMEM.
//   backward, dQ     item = 128-query tile, q-major: per 64-key block S = Q K_j^T and dP = dO V_j^T (double-buffered TMEM),
//                    dS = P (dP - delta) scale -> smem -> dQ += dS K_j.  Also produces delta = rowsum(dO * O) for the dK/dV
//                    kernel.
//   backward, dK/dV  item = 128-key tile, key-major (transposed scores): per 64-query block S^T = K Q_i^T, P^T -> smem,
//                    dP^T = V dO_i^T over the consumed S^T columns, dV += P^T dO_i, dS^T -> smem, dK += dS^T Q_i.
//                    dK | dV accumulators take 2 D <= 384 TMEM columns, the two S^T/dP^T buffers the other 128.
// The two backward kernels recompute S (7 GEMMs instead of 5) but need no cross-CTA reduction and no atomics: with d = 192
// the dQ, dK and dV accumulators of one problem (3 x 192 columns x 3 row tiles) cannot live in one SM's 512 TMEM columns.
//
// L2-distance mode (torch.cdist matmul-path semantics): score = sqrt(max(0, |q|^2 + |k|^2 - 2 q.k)); the row norms
// come from the bf16 operands in fp32.  Backward: G = P (dP - delta) scale / dist (0 where dist = 0),
// dQ = rowsum(G) q - G K, dK = colsum(G) k - G^T Q  -- the same MMAs with a rank-1 correction in the drain.
//
// All global <-> shared traffic is TMA: operands arrive as [rows x 64 col] SWIZZLE_128B chunks straight from the fused QKV
// projection output (3-D maps [B, S, cols]; rows >= S zero-filled on load, clipped on store), results leave through swizzled
// staging sub-tiles of 64 / 32 / 16 columns (SWIZZLE_128B / 64B / 32B) so that d = 96 and 112 store exactly their own columns.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace vg {
namespace {

constexpr int NTHREADS = 192;
constexpr int NTHREADS_BWD = 320, NCOMP_BWD = 256;       // dK / dV kernel: 8 compute warps (two per TMEM lane quadrant)
constexpr int QCH = 128 * 128;            // one [128 rows x 64 bf16] swizzled chunk
constexpr int BCH = 64 * 128;             // one [64 rows x 64 bf16] swizzled chunk
constexpr int MAXNK = 272;                // S rounded up to 16 must fit the S region of TMEM
constexpr int MAXKB = 5;                  // ceil(272 / 64)
constexpr float LOG2E = 1.4426950408889634f;
#ifndef VG_MT_P_TMEM
#define VG_MT_P_TMEM 1
#endif
constexpr bool P_TMEM = VG_MT_P_TMEM != 0;   // forward: P as the TMEM-resident A operand of O = P V (tcgen05.mma .ts form)

struct OutMaps { CUtensorMap m64, m32, m16; };

struct MtGeo {
  int B, H, S, NK;          // NK = S rounded up to 16
  int n_t;                  // 128-row tiles per problem
  int n_b;                  // 64-row blocks per problem
  int64_t ld, ldo;          // element pitch of q/k/v rows and of o/dO rows (direct global reads: norms, delta)
  float scale;
  float* lse;               // [B, H, S]
  float* delta;             // [B, H, S]
  const bf16 *q, *k, *o;
  bf16 *dq, *dk, *dv;       // backward outputs written straight from registers (row pitch ldd)
  int64_t ldd;
  int skew_ns, no_prefetch;    // experiment knobs (VG_ATTN_SKEW, VG_ATTN_NOPF)
  unsigned long long* trace;   // bring-up timeline of CTA 0 (vg_attention_set_trace) or NULL
};

// timeline of CTA 0: per role (0 producer, 1 MMA issuer, 2 compute leader) up to 512 (code, %globaltimer) pairs
struct Tracer {
  unsigned long long* p; int n;
  __device__ __forceinline__ Tracer(unsigned long long* base, int role) : p(base && blockIdx.x == 0 ? base + role * 1024 : nullptr), n(0) {}
  __device__ __forceinline__ void operator()(int code) {
    if (p && n < 512) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      p[2 * n] = (unsigned long long)code; p[2 * n + 1] = t; ++n;
    }
  }
};

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void tmem_ld16p(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
// n = 32 or 16 accumulator columns of this thread's lane -> v[0..n)
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, uint32_t (&v)[32], int n) {
  if (n >= 32) tmem_ld32_nowait(taddr, v); else tmem_ld16p(taddr, v);
  tmem_ld_wait();
}
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// ---- staging tile of a [128 x D] bf16 result: D/64 SW128 chunks, then a 32-column SW64 sub-tile, then a 16-column SW32 one
template <int D> struct Stg {
  static constexpr int FULL = D / 64;
  static constexpr bool R32 = (D % 64) >= 32;
  static constexpr bool R16 = (D % 32) >= 16;
};
// ROWS = rows of the staging tile (128: a whole accumulator tile; 64: one half, staged in a ring stage)
template <int D, int ROWS = 128>
__device__ __forceinline__ uint32_t stg_addr(uint32_t base, int row, int c8) {   // 16-byte chunk c8 = column / 8 of `row`
  constexpr int FULL = Stg<D>::FULL;
  if (c8 < FULL * 8) return base + (uint32_t)(c8 >> 3) * (ROWS * 128) + (uint32_t)row * 128u + ((((uint32_t)c8 & 7u) ^ ((uint32_t)row & 7u)) << 4);
  uint32_t b2 = base + FULL * (ROWS * 128);
  int c = c8 - FULL * 8;
  if (Stg<D>::R32) {
    if (c < 4) return b2 + (uint32_t)row * 64u + (((uint32_t)c ^ (((uint32_t)row >> 1) & 3u)) << 4);
    b2 += ROWS * 64; c -= 4;
  }
  return b2 + (uint32_t)row * 32u + (((uint32_t)c ^ (((uint32_t)row >> 2) & 1u)) << 4);
}
template <int D, int ROWS = 128>
__device__ __forceinline__ void stg_store(const OutMaps& m, uint32_t base, int col0, int row0, int b) {
  constexpr int FULL = Stg<D>::FULL;
#pragma unroll
  for (int i = 0; i < FULL; ++i) tma_store_3d(&m.m64, base + i * (ROWS * 128), col0 + 64 * i, row0, b);
  if (Stg<D>::R32) tma_store_3d(&m.m32, base + FULL * (ROWS * 128), col0 + 64 * FULL, row0, b);
  if (Stg<D>::R16) tma_store_3d(&m.m16, base + FULL * (ROWS * 128) + (Stg<D>::R32 ? ROWS * 64 : 0), col0 + 64 * FULL + (Stg<D>::R32 ? 32 : 0), row0, b);
}
// 32 (or 16) fp32 accumulator values -> bf16 -> staging columns [c, c + n) of `row`
template <int D, int ROWS = 128>
__device__ __forceinline__ void stg_write(uint32_t base, int row, int c, const float* o, int n) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (8 * i < n)
      sts128(stg_addr<D, ROWS>(base, row, (c >> 3) + i), pack_bf16(o[8 * i], o[8 * i + 1]), pack_bf16(o[8 * i + 2], o[8 * i + 3]),
             pack_bf16(o[8 * i + 4], o[8 * i + 5]), pack_bf16(o[8 * i + 6], o[8 * i + 7]));
}
// squared norm of D contiguous bf16 (16-byte aligned), fp32
template <int D>
__device__ __forceinline__ float row_sqnorm(const bf16* p) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < D; c += 8) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p + c));
    const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float a = bf16_lo(u[j]), b = bf16_hi(u[j]); s = fmaf(a, a, s); s = fmaf(b, b, s); }
  }
  return s;
}
// this thread's row of a K-major operand tile ([rows x 64 col] SW128 chunks, `chunk_bytes` apart) -> fp32
template <int D>
__device__ __forceinline__ void read_tile_row(uint32_t tile, uint32_t chunk_bytes, int row, float* out) {
#pragma unroll
  for (int c8 = 0; c8 < D / 8; ++c8) {
    uint32_t a0, a1, a2, a3;
    lds128(swz(tile + (uint32_t)(c8 >> 3) * chunk_bytes, row, c8 & 7), a0, a1, a2, a3);
    out[8 * c8] = bf16_lo(a0); out[8 * c8 + 1] = bf16_hi(a0); out[8 * c8 + 2] = bf16_lo(a1); out[8 * c8 + 3] = bf16_hi(a1);
    out[8 * c8 + 4] = bf16_lo(a2); out[8 * c8 + 5] = bf16_hi(a2); out[8 * c8 + 6] = bf16_lo(a3); out[8 * c8 + 7] = bf16_hi(a3);
  }
}
// K-major descriptor of k-step `k` (16 bf16) inside a tile made of 64-column chunks `chunk_bytes` apart
__device__ __forceinline__ uint64_t kdesc(uint32_t tile, uint32_t chunk_bytes, int k) {
  return desc_k(tile + (uint32_t)(k >> 2) * chunk_bytes + (uint32_t)(k & 3) * 32u);
}

__device__ __forceinline__ float sqrta(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// raw accumulator value -> score: the dot product itself, or the distance sqrt(max(0, |q|^2 + |k|^2 - 2 q.k)) (nn = |q|^2 + |k|^2)
template <int MODE> __device__ __forceinline__ float score_of(float dot, float nn) {
  return MODE == VG_ATTN_L2 ? sqrta(fmaxf(fmaf(-2.f, dot, nn), 0.f)) : dot;
}

// The compute warps run ONE warp per scheduler, so nothing hides instruction latency: the per-element loops below are
// straight-line for a compile-time column count N (no per-element branches; MASK only for the group that crosses S) and keep
// four independent accumulator chains.
// forward pass 1: running row maximum over N accumulator columns starting at key c
template <int N, bool MASK, int MODE>
__device__ __forceinline__ void fwd_max(const uint32_t (&v)[32], int c, int S, float qq, const float* kn, float (&m)[4]) {
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const float s = score_of<MODE>(__uint_as_float(v[j]), MODE == VG_ATTN_L2 ? qq + kn[c + j] : 0.f);
    if (!MASK || c + j < S) m[j & 3] = fmaxf(m[j & 3], s);
  }
}
// forward pass 2: p = exp2(s * sc2 - mb) for N columns, row sum, packed bf16
template <int N, bool MASK, int MODE>
__device__ __forceinline__ void fwd_exp(const uint32_t (&v)[32], int c, int S, float qq, const float* kn, float sc2, float mb, float (&l)[4],
                                        uint32_t (&pk)[16]) {
  float p[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const float s = score_of<MODE>(__uint_as_float(v[j]), MODE == VG_ATTN_L2 ? qq + kn[c + j] : 0.f);
    p[j] = ex2a(fmaf(s, sc2, -mb));
    if (MASK && c + j >= S) p[j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < N; ++j) l[j & 3] += p[j];
#pragma unroll
  for (int j = 0; j < N / 2; ++j) pk[j] = pack_bf16(p[2 * j], p[2 * j + 1]);
}
// backward (q-major): dS (dot) or G (L2) = P (dP - delta) scale [/ dist] for N columns, packed bf16; gsum += row sum (L2)
template <int N, bool MASK, int MODE>
__device__ __forceinline__ void dq_group(const uint32_t (&sv)[32], const uint32_t (&dv)[32], int c, int S, float qq, const float* kn, float sc2,
                                         float lse2, float delta, float scale, float& gsum, uint32_t (&pk)[16]) {
  float d[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const float s = score_of<MODE>(__uint_as_float(sv[j]), MODE == VG_ATTN_L2 ? qq + kn[c + j] : 0.f);
    float f = scale;
    if (MODE == VG_ATTN_L2) f = s > 0.f ? scale * rcpa(s) : 0.f;
    d[j] = ex2a(fmaf(s, sc2, -lse2)) * (__uint_as_float(dv[j]) - delta) * f;
    if (MASK && c + j >= S) d[j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < N / 2; ++j) pk[j] = pack_bf16(d[2 * j], d[2 * j + 1]);
  if (MODE == VG_ATTN_L2) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < N / 2; ++j) { a[j & 3] += bf16_lo(pk[j]); a[(j + 2) & 3] += bf16_hi(pk[j]); }
    gsum += (a[0] + a[1]) + (a[2] + a[3]);
  }
}
// backward (key-major) stage A: P^T (packed bf16 -> pk) and the factor f = P scale [/ dist] (packed bf16 -> f) for N query columns
template <int N, int MODE>
__device__ __forceinline__ void dkv_stage_a(const uint32_t (&sv)[32], int q0, float kk2, const float* qn_s, const float* lse_s, float sc2, float scale,
                                            uint32_t (&pk)[16], uint32_t* f) {
  float p[N], ff[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const float s = score_of<MODE>(__uint_as_float(sv[j]), MODE == VG_ATTN_L2 ? kk2 + qn_s[q0 + j] : 0.f);
    float fs = scale;
    if (MODE == VG_ATTN_L2) fs = s > 0.f ? scale * rcpa(s) : 0.f;
    p[j] = ex2a(fmaf(s, sc2, -lse_s[q0 + j]));
    ff[j] = p[j] * fs;
  }
#pragma unroll
  for (int j = 0; j < N / 2; ++j) { pk[j] = pack_bf16(p[2 * j], p[2 * j + 1]); f[j] = pack_bf16(ff[2 * j], ff[2 * j + 1]); }
}
// stage B: dS^T = f (dP^T - delta[q]) for N query columns, packed bf16; gsum += row sum (L2)
template <int N, int MODE>
__device__ __forceinline__ void dkv_stage_b(const uint32_t (&dv)[32], int q0, const float* del_s, const uint32_t* f, float& gsum, uint32_t (&pk)[16]) {
#pragma unroll
  for (int j = 0; j < N / 2; ++j) {
    const float d0 = bf16_lo(f[j]) * (__uint_as_float(dv[2 * j]) - del_s[q0 + 2 * j]);
    const float d1 = bf16_hi(f[j]) * (__uint_as_float(dv[2 * j + 1]) - del_s[q0 + 2 * j + 1]);
    pk[j] = pack_bf16(d0, d1);
  }
  if (MODE == VG_ATTN_L2) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < N / 2; ++j) { a[j & 3] += bf16_lo(pk[j]); a[(j + 2) & 3] += bf16_hi(pk[j]); }
    gsum += (a[0] + a[1]) + (a[2] + a[3]);
  }
}
// n / 8 16-byte chunks of packed bf16 -> row `row` of a K-major [128 x 64] SW128 tile, starting at column c
template <int N>
__device__ __forceinline__ void put_row(uint32_t tile, int row, int c, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int i = 0; i < N / 8; ++i) sts128(swz(tile, row, (c >> 3) + i), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
}

__device__ __forceinline__ void tmem_alloc512(uint32_t slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free512(uint32_t tmem) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ================================================================================================ forward
// K of the (batch, head) problem stays resident in smem for all of its query tiles (loaded once: [NK rows x 64 col] SW128 chunks),
// so S = Q K^T is ONE N = 256 MMA (+ one N = NK - 256) per k-step -- an SS-MMA pays 64 B/clk of operand fetch, i.e. the
// 128 x 16 A tile costs as much as 128 columns of B, and N = 64 blocks were 3x off the tensor pipe's rate.  V streams through a
// ring of 64-key blocks.  P is written back to TMEM as bf16 over the consumed S columns and feeds O += P V as the TMEM-resident
// A operand (tcgen05.mma .ts): no P tile in smem, no A fetch.  The output staging tile aliases the V ring; the producer warp
// issues the TMA store (it is idle then), so no compute thread ever waits for a store.
template <int D> struct FwdCfg {
  static constexpr int DC = (D + 63) / 64, KS = D / 16;
  static constexpr int NSTV = 3;                             // ring stages of one 64-key V block
  static constexpr int KCH = MAXNK * 128;                    // one resident K chunk: [272 rows x 64 col]
  static constexpr int O_COL = 288;                          // S (fp32) / P (bf16, in place) in TMEM columns [0, 272), O in [288, 288 + D)
  static constexpr int Q_OFF = 0, K_OFF = DC * QCH, V_OFF = K_OFF + DC * KCH, KN_OFF = V_OFF + NSTV * DC * BCH;
  static constexpr int BAR_OFF = KN_OFF + MAXNK * 4, NBAR = 8 + 2 * NSTV + MAXKB;
  static constexpr int SMEM = BAR_OFF + NBAR * 8 + 16;
  static_assert(NSTV * DC * BCH >= 128 * D * 2, "the output staging tile aliases the V ring");
};

template <int D, int MODE>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_fwd_mt_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k64,
                   const __grid_constant__ CUtensorMap map_k16, const __grid_constant__ CUtensorMap map_v,
                   const __grid_constant__ OutMaps map_o, const MtGeo g) {
  using C = FwdCfg<D>;
  constexpr int DC = C::DC, KS = C::KS, NSTV = C::NSTV, O_COL = C::O_COL, KCH = C::KCH;
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  const uint32_t base = smem_u32(smem_dyn);
  const uint32_t q_t = base + C::Q_OFF, k_t = base + C::K_OFF, vring = base + C::V_OFF, stg = vring, bar = base + C::BAR_OFF;
  float* kn = reinterpret_cast<float*>(smem_dyn + C::KN_OFF);
  const uint32_t q_full = bar, q_empty = bar + 8, k_full = bar + 16, k_empty = bar + 24, s_full = bar + 32, o_full = bar + 40,
                 o_free = bar + 48, stg_full = bar + 56;
  auto v_full = [&](int i) { return bar + 8u * (8 + i); };
  auto v_empty = [&](int i) { return bar + 8u * (8 + NSTV + i); };
  auto p_full = [&](int i) { return bar + 8u * (8 + 2 * NSTV + i); };
  const uint32_t tmem_slot = bar + 8u * C::NBAR;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_dyn + C::BAR_OFF + 8 * C::NBAR);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if (base & 1023u) { printf("attention_mt: dynamic smem base not 1024-aligned\n"); __trap(); }
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k64) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    mbar_init(q_full, 1); mbar_init(q_empty, 1); mbar_init(k_full, 1); mbar_init(k_empty, 1); mbar_init(s_full, 1);
    mbar_init(o_full, 1); mbar_init(o_free, 128); mbar_init(stg_full, 128);
    for (int i = 0; i < NSTV; ++i) { mbar_init(v_full(i), 1); mbar_init(v_empty(i), 1); }
    for (int i = 0; i < MAXKB; ++i) mbar_init(p_full(i), 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc512(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  pdl_trigger();
  pdl_wait();
  const int total = g.B * g.H;
  const int n64 = g.NK / 64, n16 = (g.NK % 64) / 16;              // K is loaded as n64 boxes of 64 rows + n16 boxes of 16 rows per chunk

  if (warp == 0) {
    if (lane == 0) {
      uint32_t vc = 0, tc = 0, ic = 0;              // V-block, tile and item counters
      auto load_k = [&](int w) {
        const int b = w / g.H, col0 = (w % g.H) * D;
        mbar_expect_tx(k_full, (uint32_t)(DC * g.NK * 128));
        for (int c = 0; c < DC; ++c) {
          for (int i = 0; i < n64; ++i) tma_load_3d(k_t + c * KCH + i * BCH, &map_k64, k_full, col0 + 64 * c, 64 * i, b);
          for (int i = 0; i < n16; ++i) tma_load_3d(k_t + c * KCH + n64 * BCH + i * 2048, &map_k16, k_full, col0 + 64 * c, 64 * n64 + 16 * i, b);
        }
      };
      int pend_w = -1, pend_t = 0;                  // tile whose output the compute threads are about to stage
      auto store_pending = [&]() {                  // the staging tile (over the V ring) is written: store it, wait for the read
        if (pend_w < 0) return;
        mbar_wait(stg_full, (tc - 1) & 1u);
        stg_store<D>(map_o, stg, (pend_w % g.H) * D, pend_t * 128, pend_w / g.H);
        tma_commit();
        tma_wait_read();
        pend_w = -1;
      };
      if ((int)blockIdx.x < total) load_k(blockIdx.x);
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++ic) {
        const int b = w / g.H, col0 = (w % g.H) * D;
        for (int t = 0; t < g.n_t; ++t, ++tc) {
          mbar_wait(q_empty, (tc & 1u) ^ 1u);
          mbar_expect_tx(q_full, DC * QCH);
#pragma unroll
          for (int c = 0; c < DC; ++c) tma_load_3d(q_t + c * QCH, &map_q, q_full, col0 + 64 * c, t * 128, b);
          store_pending();                           // previous tile's O: its V blocks are all consumed, none of this tile's is issued yet
          const bool last = t + 1 == g.n_t;
          for (int kb = 0; kb < g.n_b; ++kb, ++vc) {
            if (last && kb == min(g.n_b, NSTV) && w + (int)gridDim.x < total) {   // next problem's K, as soon as this one's last S is done
              mbar_wait(k_empty, ic & 1u);
              load_k(w + gridDim.x);
            }
            const int st = vc % NSTV;
            mbar_wait(v_empty(st), ((vc / NSTV) & 1u) ^ 1u);
            mbar_expect_tx(v_full(st), DC * BCH);
#pragma unroll
            for (int c = 0; c < DC; ++c) tma_load_3d(vring + (st * DC + c) * BCH, &map_v, v_full(st), col0 + 64 * c, kb * 64, b);
          }
          if (last && g.n_b <= NSTV && w + (int)gridDim.x < total) { mbar_wait(k_empty, ic & 1u); load_k(w + gridDim.x); }
          pend_w = w; pend_t = t;
        }
      }
      store_pending();
      tma_wait_all();
    }
  } else if (warp == 1) {
    // MMA issue: the whole warp runs the loop (warp-uniform control flow and operands), one elected lane issues
    {
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t idesc_pv = make_idesc(128, D, 0, 1);            // O = P V : A from TMEM, B MN-major
      const SDesc qd = sdesc_k(q_t), kd0 = sdesc_k(k_t);
      uint32_t vc = 0, tc = 0, ic = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++ic) {
        mbar_wait(k_full, ic & 1u);
        for (int t = 0; t < g.n_t; ++t, ++tc) {
          const uint32_t par = tc & 1u;
          mbar_wait(q_full, par);
          tc_fence_after();
          for (int n0 = 0; n0 < g.NK; n0 += 256) {                    // S[:, n0 ..] = Q K[n0 ..]^T, whole key range, in program order
            const uint32_t idesc_s = make_idesc(128, min(256, g.NK - n0), 0, 0);     // behind the previous tile's P V (same TMEM columns)
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < KS; ++k)
                tc_mma_d(tm + (uint32_t)n0, qd, kstep_off(k, QCH), kd0, kstep_off(k, KCH) + (uint32_t)n0 * 8u, idesc_s, k > 0);
            }
            __syncwarp();
          }
          if (elect_one()) {
            tc_commit(s_full);
            tc_commit(q_empty);
            if (t + 1 == g.n_t) tc_commit(k_empty);
          }
          __syncwarp();
          mbar_wait(o_free, par ^ 1u);
          for (int kb = 0; kb < g.n_b; ++kb, ++vc) {                  // O += P_kb V_kb
            const int st = vc % NSTV, nk = min(64, g.NK - 64 * kb);
            mbar_wait(v_full(st), (vc / NSTV) & 1u);
            mbar_wait(p_full(kb), par);
            tc_fence_after();
            const SDesc vd = sdesc_mn(vring + st * DC * BCH, BCH);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                if (kk < nk / 16) tc_mma_ts_d(tm + O_COL, tm + (uint32_t)(32 * kb + 8 * kk), vd, kk * 128u, idesc_pv, (kb > 0 || kk > 0) ? 1u : 0u);
              tc_commit(v_empty(st));
            }
            __syncwarp();
          }
          if (elect_one()) tc_commit(o_full);
          __syncwarp();
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax + output drain: thread = query row
    const int quad = warp & 3, row = quad * 32 + lane, tid = threadIdx.x - 64;
    const uint32_t t_lane = tmem + ((uint32_t)(quad * 32) << 16);
    const float sc2 = g.scale * LOG2E;
    uint32_t tc = 0;
    Tracer tr(threadIdx.x == 128 ? g.trace : nullptr, 2);
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int b = w / g.H, h = w % g.H, col0 = h * D;
      if (MODE == VG_ATTN_L2) {                       // |k_j|^2 of this problem's keys -> smem (broadcast reads below)
        named_bar(2, 128);
        for (int j = tid; j < g.NK; j += 128)
          kn[j] = j < g.S ? row_sqnorm<D>(g.k + ((int64_t)b * g.S + j) * g.ld + col0) : 0.f;
        named_bar(2, 128);
      }
      for (int t = 0; t < g.n_t; ++t, ++tc) {
        const uint32_t par = tc & 1u;
        const int row_g = t * 128 + row;
        const bool warp_on = t * 128 + quad * 32 < g.S;               // warp-uniform: this warp owns real query rows
        float qq = 0.f;
        if (MODE == VG_ATTN_L2 && row_g < g.S) qq = row_sqnorm<D>(g.q + ((int64_t)b * g.S + row_g) * g.ld + col0);
        mbar_wait(s_full, par);
        tr(52);
        tc_fence_after();
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, l4[4] = {0.f, 0.f, 0.f, 0.f};
        // pass 1: row maximum of the raw scores (dot) / distances (L2) over the real keys
        if (warp_on) {
          for (int c = 0; c < g.NK; c += 32) {
            uint32_t v[32];
            if (g.NK - c >= 32) {
              tmem_ld32(t_lane + (uint32_t)c, v);
              if (c + 32 <= g.S) fwd_max<32, false, MODE>(v, c, g.S, qq, kn, m4); else fwd_max<32, true, MODE>(v, c, g.S, qq, kn, m4);
            } else {
              tmem_ld16p(t_lane + (uint32_t)c, v);
              tmem_ld_wait();
              if (c + 16 <= g.S) fwd_max<16, false, MODE>(v, c, g.S, qq, kn, m4); else fwd_max<16, true, MODE>(v, c, g.S, qq, kn, m4);
            }
          }
        }
        const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        tr(53);
        // pass 2: P = exp2((s - m) scale log2e) -> bf16 -> TMEM in place (P of keys [2c, 2c + 2) lands in column c, already read),
        // one 64-key chunk at a time: the P V MMAs trail by one chunk
        const float mb = m * sc2;
        for (int kb = 0; kb < g.n_b; ++kb) {
          if (warp_on) {
            const int nk = min(64, g.NK - 64 * kb), c0 = 64 * kb;
            uint32_t va[32], vb2[32], pk[16];
            if (nk == 64) {                          // both halves requested before either is exponentiated
              tmem_ld32_nowait(t_lane + (uint32_t)c0, va);
              tmem_ld32_nowait(t_lane + (uint32_t)(c0 + 32), vb2);
              tmem_ld_wait();
              if (c0 + 64 <= g.S) {
                fwd_exp<32, false, MODE>(va, c0, g.S, qq, kn, sc2, mb, l4, pk);
                tmem_st16(t_lane + (uint32_t)(c0 >> 1), pk);
                fwd_exp<32, false, MODE>(vb2, c0 + 32, g.S, qq, kn, sc2, mb, l4, pk);
                tmem_st16(t_lane + (uint32_t)((c0 >> 1) + 16), pk);
              } else {
                fwd_exp<32, true, MODE>(va, c0, g.S, qq, kn, sc2, mb, l4, pk);
                tmem_st16(t_lane + (uint32_t)(c0 >> 1), pk);
                fwd_exp<32, true, MODE>(vb2, c0 + 32, g.S, qq, kn, sc2, mb, l4, pk);
                tmem_st16(t_lane + (uint32_t)((c0 >> 1) + 16), pk);
              }
            } else {
#pragma unroll
              for (int c = 0; c < 64; c += 32) {
                if (c >= nk) break;
                if (nk - c >= 32) {
                  tmem_ld32(t_lane + (uint32_t)(c0 + c), va);
                  fwd_exp<32, true, MODE>(va, c0 + c, g.S, qq, kn, sc2, mb, l4, pk);
                  tmem_st16(t_lane + (uint32_t)((c0 + c) >> 1), pk);
                } else {
                  tmem_ld16p(t_lane + (uint32_t)(c0 + c), va);
                  tmem_ld_wait();
                  fwd_exp<16, true, MODE>(va, c0 + c, g.S, qq, kn, sc2, mb, l4, pk);
                  tmem_st8(t_lane + (uint32_t)((c0 + c) >> 1), pk);
                }
              }
            }
            tmem_st_wait();
          }
          tc_fence_before();                           // P chunk written (and its S columns read) before the tensor core touches them
          mbar_arrive(p_full(kb));
          tr(60 + kb);
        }
        const float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        const float inv_l = 1.0f / l;
        if (row_g < g.S) g.lse[(int64_t)w * g.S + row_g] = m * g.scale + __logf(l);
        // ---- drain O: TMEM -> * 1/l -> bf16 -> staging over the V ring (every V block of the tile is consumed once o_full fires,
        //      and the producer issues no V load of the next tile before it has stored this one)
        mbar_wait(o_full, par);
        tr(55);
        tc_fence_after();
        if (warp_on) {
#pragma unroll
          for (int c = 0; c < D; c += 32) {
            uint32_t v[32];
            const int n = (D - c) >= 32 ? 32 : 16;
            tmem_ldn(t_lane + (uint32_t)(O_COL + c), v, n);
            float o[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]) * inv_l;
            stg_write<D>(stg, row, c, o, n);
          }
        }
        tc_fence_before();
        mbar_arrive(o_free);
        fence_async_smem();
        mbar_arrive(stg_full);
        tr(56);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_free512(tmem); }
}

// ================================================================================================ backward: dQ (+ delta)
// The Q and dO tiles are A operands of 2 x n_b MMAs each.  From smem every MMA pays 64 clk of A fetch on top of N / 2 clk of B
// (operand fetch runs at 64 B/clk), which at N = 64 is three times the tensor pipe's rate; so the compute threads -- which read
// their dO row anyway for delta = rowsum(dO * O) -- copy both rows into TMEM once per tile (tcgen05.st, two bf16 per column) and
// S = Q K_j^T, dP = dO V_j^T run as TS-MMAs.  TMEM: Q | dO as operands (D columns), one [S | dP] pair of 64-key blocks (128 columns;
// released as soon as the block sits in registers, so the next pair is computed while this one is exponentiated), dQ (D columns).
// dS goes to smem (double buffered) as the A operand of dQ += dS K_j.
//
// Shared memory is ONE in-order pool of [64 rows x D] stages.  A TMA load takes ~2 us here while a 64-key block lasts ~1 us, and a
// K block lives from its S MMA to its dQ MMA, so separate two-stage K / V rings left the tensor core waiting for K in every
// block.  Through the pool flow, in a fixed order per tile: the two halves of dO and of Q (released once copied to TMEM), then
// V_j, K_j for every key block, then the two halves of the dQ staging tile.  Every role derives a request's stage and phase from
// the running request number, so the producer runs up to a whole pool (~3 key blocks, and the next tile's dO / Q) ahead; it also
// issues the dQ stores (it polls for the staged tile between its loads), so no compute thread waits for a store.
template <int D> struct DqCfg {
  static constexpr int DC = (D + 63) / 64, KS = D / 16;
  static constexpr int STG_BYTES = DC * BCH;                 // one [64 rows x D] stage
  static constexpr int NP = D > 128 ? 8 : 11;                // pool stages
  static constexpr int QA_COL = 0, DOA_COL = D / 2, S_COL = D, DP_COL = D + 64, DQ_COL = D + 128;
  static_assert(DQ_COL + D <= 512, "TMEM layout");
  static constexpr int POOL_OFF = 0, DS_OFF = NP * STG_BYTES, KN_OFF = DS_OFF + 2 * QCH, XCH_OFF = KN_OFF + MAXNK * 4,
                        BAR_OFF = XCH_OFF + (D > 128 ? 2 : 4) * 128 * 4;   // XCH: per-row delta, |q|^2 and (L2, D <= 128) rowsum(G) halves, exchanged between a row's two threads
  static constexpr int NBAR = 2 * NP + 10;
  static constexpr int SMEM = BAR_OFF + NBAR * 8 + 16;
  static_assert(SMEM <= 232448, "shared memory");
  static_assert(STG_BYTES >= 64 * D * 2, "a staging half fits a stage");
};

// this thread's row of a K-major [rows x D] SW128 operand tile (64-column chunks `chunk_bytes` apart) -> TMEM columns
// [col, col + D / 2) of its lane (bf16 A operand of a TS-MMA: element k of the row in column k / 2, low half first), 32 elements
// (16 columns) at a time; `f(c, r, n)` sees every chunk (c = first TMEM column of the chunk, r = its n = 16 or 8 packed registers)
// n (32 or 16) fp32 values -> bf16 -> 16-byte global stores (p 16-byte aligned)
__device__ __forceinline__ void store_row_bf16(bf16* p, const float (&o)[32], int n) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (8 * i < n) {
      uint4 u;
      u.x = pack_bf16(o[8 * i], o[8 * i + 1]); u.y = pack_bf16(o[8 * i + 2], o[8 * i + 3]);
      u.z = pack_bf16(o[8 * i + 4], o[8 * i + 5]); u.w = pack_bf16(o[8 * i + 6], o[8 * i + 7]);
      *reinterpret_cast<uint4*>(p + 8 * i) = u;
    }
}

template <int D, typename F>
__device__ __forceinline__ void row_to_tmem(uint32_t tile, uint32_t chunk_bytes, int row, uint32_t taddr, F&& f) {
#pragma unroll
  for (int c = 0; c < D / 2; c += 16) {
    uint32_t r[16];
    const int n = (D / 2 - c) >= 16 ? 16 : 8;            // D / 2 is a multiple of 8
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (4 * i < n) {
        const int c8 = (c >> 2) + i;                       // 16-byte chunk = 8 elements = 4 columns
        lds128(swz(tile + (uint32_t)(c8 >> 3) * chunk_bytes, row, c8 & 7), r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
      }
    if (n == 16) tmem_st16(taddr + (uint32_t)c, r); else tmem_st8(taddr + (uint32_t)c, r);
    f(c, r, n);
  }
}

template <int D, int MODE>
__global__ void __launch_bounds__(NTHREADS_BWD, 1)      // 10 warps: 3 on two of the four sub-partitions -> at most 168 registers
attn_bwd_dq_mt_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                      const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                      const __grid_constant__ OutMaps map_dq, const MtGeo g) {
  using C = DqCfg<D>;
  constexpr int DC = C::DC, KS = C::KS, NP = C::NP, STG = C::STG_BYTES;
  constexpr int QA_COL = C::QA_COL, DOA_COL = C::DOA_COL, S_COL = C::S_COL, DP_COL = C::DP_COL, DQ_COL = C::DQ_COL;
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  const uint32_t base = smem_u32(smem_dyn);
  const uint32_t pool = base + C::POOL_OFF, ds_t = base + C::DS_OFF, bar = base + C::BAR_OFF;
  float* kn = reinterpret_cast<float*>(smem_dyn + C::KN_OFF);
  float* xd = reinterpret_cast<float*>(smem_dyn + C::XCH_OFF);     // delta of a row (written by the row's first thread)
  float* xq = xd + 128;                                            // L2: |q|^2 (second thread), then [128, 384): rowsum(G) halves
  auto full = [&](uint32_t st) { return bar + 8u * st; };
  auto empty = [&](uint32_t st) { return bar + 8u * (NP + st); };
  const uint32_t a_ready = bar + 8u * (2 * NP), stg_full = a_ready + 16, dq_full = a_ready + 24, dq_free = a_ready + 32,
                 sdp_full = a_ready + 40, sdp_free = a_ready + 48;
  auto ds_full = [&](int i) { return a_ready + 8u * (7 + i); };
  auto ds_empty = [&](int i) { return a_ready + 8u * (9 + i); };
  const uint32_t tmem_slot = bar + 8u * C::NBAR;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_dyn + C::BAR_OFF + 8 * C::NBAR);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // request numbering inside a tile: 0,1 dO halves; 2,3 Q halves; 4 + 2 j V_j; 5 + 2 j K_j; R - 2, R - 1 staging halves
  const uint32_t R = 6u + 2u * (uint32_t)g.n_b;

  if (threadIdx.x == 0) {
    if (base & 1023u) { printf("attention_mt: dynamic smem base not 1024-aligned\n"); __trap(); }
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_do) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    for (int i = 0; i < NP; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), 1); }
    mbar_init(a_ready, NCOMP_BWD); mbar_init(stg_full, NCOMP_BWD); mbar_init(dq_full, 1); mbar_init(dq_free, NCOMP_BWD);
    mbar_init(sdp_full, 1); mbar_init(sdp_free, NCOMP_BWD);
    for (int i = 0; i < 2; ++i) { mbar_init(ds_full(i), NCOMP_BWD); mbar_init(ds_empty(i), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc512(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  pdl_trigger();
  pdl_wait();
  if (g.skew_ns > 0) { const unsigned ns = (blockIdx.x % 4u) * (unsigned)g.skew_ns; if (ns) __nanosleep(ns); }
  const int total = g.B * g.H;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t tc = 0;                               // tile counter
      Tracer tr(g.trace, 3);
      // the previous tile's staged dQ: stored as soon as the compute threads have written it (polled between the loads below)
      int pend_w = -1, pend_t = 0;
      uint32_t pend_rq = 0, pend_par = 0;
      auto service_store = [&](bool block) {
        if (pend_w < 0) return;
        if (!block && !mbar_test(stg_full, pend_par)) return;
        if (block) mbar_wait(stg_full, pend_par);
        const int pb = pend_w / g.H, pc = (pend_w % g.H) * D;
        stg_store<D, 64>(map_dq, pool + (pend_rq % NP) * STG, pc, pend_t * 128, pb);
        stg_store<D, 64>(map_dq, pool + ((pend_rq + 1) % NP) * STG, pc, pend_t * 128 + 64, pb);
        tma_commit();
        tma_wait_read();
        mbar_arrive(empty(pend_rq % NP));
        mbar_arrive(empty((pend_rq + 1) % NP));
        pend_w = -1;
      };
      auto acquire = [&](uint32_t rq) {              // the stage of request rq is free again (its previous user released it)
        const uint32_t st = rq % NP;
        if (pend_w >= 0 && (st == pend_rq % NP || st == (pend_rq + 1) % NP)) service_store(true);
        mbar_wait(empty(st), ((rq / NP) & 1u) ^ 1u);
        return st;
      };
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int b = w / g.H, col0 = (w % g.H) * D;
        for (int t = 0; t < g.n_t; ++t, ++tc) {
          const uint32_t rq0 = tc * R;
          for (int hf = 0; hf < 2; ++hf) {           // dO halves, then Q halves
            const uint32_t st = acquire(rq0 + hf);
            mbar_expect_tx(full(st), STG);
#pragma unroll
            for (int c = 0; c < DC; ++c) tma_load_3d(pool + st * STG + c * BCH, &map_do, full(st), col0 + 64 * c, t * 128 + 64 * hf, b);
          }
          for (int hf = 0; hf < 2; ++hf) {
            const uint32_t st = acquire(rq0 + 2 + hf);
            mbar_expect_tx(full(st), STG);
#pragma unroll
            for (int c = 0; c < DC; ++c) tma_load_3d(pool + st * STG + c * BCH, &map_q, full(st), col0 + 64 * c, t * 128 + 64 * hf, b);
          }
          tr(1);
          {
            // ask for the NEXT tile's operands in L2 now, a whole tile ahead: their TMA loads then only pay the L2 latency
            const bool last_t = t + 1 == g.n_t;
            const int wn = last_t ? w + (int)gridDim.x : w, tn = last_t ? 0 : t + 1;
            if (wn < total && !g.no_prefetch) {
              const int bn = wn / g.H, cn = (wn % g.H) * D;
              for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                for (int c = 0; c < DC; ++c) {
                  tma_prefetch_3d(&map_do, cn + 64 * c, tn * 128 + 64 * hf, bn);
                  tma_prefetch_3d(&map_q, cn + 64 * c, tn * 128 + 64 * hf, bn);
                }
            }
          }
          for (int kb = 0; kb < g.n_b; ++kb) {
            service_store(false);
            uint32_t st = acquire(rq0 + 4 + 2 * kb);
            tr(110 + kb);
            mbar_expect_tx(full(st), STG);
#pragma unroll
            for (int c = 0; c < DC; ++c) tma_load_3d(pool + st * STG + c * BCH, &map_v, full(st), col0 + 64 * c, kb * 64, b);
            st = acquire(rq0 + 5 + 2 * kb);
            tr(100 + kb);
            mbar_expect_tx(full(st), STG);
#pragma unroll
            for (int c = 0; c < DC; ++c) tma_load_3d(pool + st * STG + c * BCH, &map_k, full(st), col0 + 64 * c, kb * 64, b);
          }
          service_store(true);                       // at the latest here: one staged tile outstanding at a time
          // this tile's staging halves: wait until their previous users are done, then hand them to the compute threads by
          // completing the stages' `full` phase (every request completes exactly one phase of its stage's barrier pair)
          mbar_arrive(full(acquire(rq0 + R - 2)));
          mbar_arrive(full(acquire(rq0 + R - 1)));
          pend_w = w; pend_t = t; pend_rq = rq0 + R - 2; pend_par = tc & 1u;
        }
      }
      service_store(true);
      tma_wait_all();
    }
  } else if (warp == 1) {
    // MMA issue: the whole warp runs the loop (warp-uniform control flow and operands), one elected lane issues
    {
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t idesc_dq = make_idesc(128, D, 0, 1);            // dQ = dS K : A K-major (dS), B MN-major (K block)
      uint32_t bc0 = 0, tc = 0;
      Tracer tr(lane == 0 ? g.trace : nullptr, 4);
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        for (int t = 0; t < g.n_t; ++t, ++tc) {
          const uint32_t tpar = tc & 1u, rq0 = tc * R;
          mbar_wait(a_ready, tpar);                                   // Q and dO rows of this tile are in TMEM
          tr(10);
          tc_fence_after();
          // event loop: dQ(j) as soon as its dS block is written (it releases the K stage), otherwise the next S / dP pair as
          // soon as its K, V blocks have landed and the previous pair sits in registers
          int js = 0, jq = 0;                                         // next block to issue S/dP for, next block to issue dQ for
          while (jq < g.n_b) {
            if (jq < js) {
              const uint32_t c = bc0 + jq, rk = rq0 + 5 + 2 * jq;
              const int buf = c & 1, nk = min(64, g.NK - 64 * jq);
              if (mbar_test(ds_full(buf), (c >> 1) & 1u) && (jq > 0 || mbar_test(dq_free, tpar ^ 1u))) {
                tr(40 + jq);
                tc_fence_after();
                const SDesc kd = sdesc_mn(pool + (rk % NP) * STG, BCH), dsd = sdesc_k(ds_t + buf * QCH);
                if (elect_one()) {
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk)
                    if (kk < nk / 16) tc_mma_d(tm + DQ_COL, dsd, kk * 2u, kd, kk * 128u, idesc_dq, (jq > 0 || kk > 0) ? 1u : 0u);
                  tc_commit(ds_empty(buf));
                  tc_commit(empty(rk % NP));
                }
                __syncwarp();
                ++jq;
                continue;
              }
            }
            if (js < g.n_b) {
              const uint32_t c = bc0 + js, rv = rq0 + 4 + 2 * js, rk = rv + 1;
              const int nk = min(64, g.NK - 64 * js);
              if (mbar_test(full(rk % NP), (rk / NP) & 1u) && mbar_test(full(rv % NP), (rv / NP) & 1u) && mbar_test(sdp_free, (c & 1u) ^ 1u)) {
                tr(20 + js);
                tc_fence_after();
                const uint32_t idesc_s = make_idesc(128, nk, 0, 0);
                const SDesc kd = sdesc_k(pool + (rk % NP) * STG), vd = sdesc_k(pool + (rv % NP) * STG);
                if (elect_one()) {
#pragma unroll
                  for (int k = 0; k < KS; ++k) tc_mma_ts_d(tm + S_COL, tm + (uint32_t)(QA_COL + 8 * k), kd, kstep_off(k, BCH), idesc_s, k > 0);
#pragma unroll
                  for (int k = 0; k < KS; ++k) tc_mma_ts_d(tm + DP_COL, tm + (uint32_t)(DOA_COL + 8 * k), vd, kstep_off(k, BCH), idesc_s, k > 0);
                  tc_commit(sdp_full);
                  tc_commit(empty(rv % NP));
                }
                __syncwarp();
                ++js;
              }
            }
          }
          tr(14);
          if (elect_one()) tc_commit(dq_full);
          __syncwarp();
          bc0 += g.n_b;
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- two threads per query row (two warps per TMEM lane
    // quadrant, so that each scheduler has a second warp to run while the first waits on TMEM, MUFU or shared memory): the first
    // (half 0) moves the dO row to TMEM and computes delta, the second the Q row; in the block loop each owns 32 of the 64 keys;
    // at the end each drains half of the dQ columns
    const int quad = warp & 3, half = (warp - 2) >> 2, row = quad * 32 + lane, tid = threadIdx.x - 64;
    const uint32_t t_lane = tmem + ((uint32_t)(quad * 32) << 16);
    const float sc2 = g.scale * LOG2E;
    const int hf = row >> 6, lrow = row & 63;          // this row's half of the Q / dO / staging tiles, and its row inside the stage
    constexpr int NCH = (D + 31) / 32, CH0 = (NCH + 1) / 2;                 // 32-column drain chunks; the first thread takes CH0 of them
    uint32_t bc0 = 0, tc = 0;
    Tracer tr(threadIdx.x == 128 ? g.trace : nullptr, 5);                  // warp 4: quadrant 0, half 0
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int b = w / g.H, h = w % g.H, col0 = h * D;
      if (MODE == VG_ATTN_L2) {
        named_bar(2, NCOMP_BWD);
        for (int j = tid; j < g.NK; j += NCOMP_BWD)
          kn[j] = j < g.S ? row_sqnorm<D>(g.k + ((int64_t)b * g.S + j) * g.ld + col0) : 0.f;
      }
      for (int t = 0; t < g.n_t; ++t, ++tc) {
        const uint32_t tpar = tc & 1u, rq0 = tc * R;
        const int row_g = t * 128 + row;
        const bool row_on = row_g < g.S;
        const bool warp_on = t * 128 + quad * 32 < g.S;
        float delta = 0.f, lse2 = 1e30f, qq = 0.f;
        const uint32_t r_do = rq0 + hf, r_q = rq0 + 2 + hf;
        // delta = rowsum(dO * O).  O comes from global memory; eight lanes share a row so that every request is a whole 128-byte
        // line (one row per thread costs an L1 wavefront per row and instruction: ~2 us per tile).  This warp takes 16 rows of its
        // quadrant, four at a time; the loads are issued BEFORE the wait for the dO tile so that the latencies overlap.
        const int g8 = lane >> 3, p8 = lane & 7;
        const int dr0 = quad * 32 + half * 16 + g8;      // + 4 s: the rows this lane helps with
        uint4 ov[4][DC];
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
          const int rg = t * 128 + dr0 + 4 * s4;
          const bf16* op = g.o + ((int64_t)b * g.S + min(rg, g.S - 1)) * g.ldo + col0 + 8 * p8;
#pragma unroll
          for (int c = 0; c < DC; ++c)
            ov[s4][c] = (rg < g.S && 64 * c + 8 * p8 < D) ? __ldg(reinterpret_cast<const uint4*>(op + 64 * c)) : make_uint4(0u, 0u, 0u, 0u);
        }
        mbar_wait(full(r_do % NP), (r_do / NP) & 1u);
        tr(50);
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
          const int rr = dr0 + 4 * s4;                   // row of the tile (same 64-row half, i.e. same stage, as this thread's row)
          float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int c = 0; c < DC; ++c)
            if (64 * c + 8 * p8 < D) {
              uint32_t x, y, z, wv;
              lds128(swz(pool + (r_do % NP) * STG + c * BCH, rr & 63, p8), x, y, z, wv);
              const uint4 o4 = ov[s4][c];
              d4[0] = fmaf(bf16_lo(x), bf16_lo(o4.x), d4[0]); d4[1] = fmaf(bf16_hi(x), bf16_hi(o4.x), d4[1]);
              d4[2] = fmaf(bf16_lo(y), bf16_lo(o4.y), d4[2]); d4[3] = fmaf(bf16_hi(y), bf16_hi(o4.y), d4[3]);
              d4[0] = fmaf(bf16_lo(z), bf16_lo(o4.z), d4[0]); d4[1] = fmaf(bf16_hi(z), bf16_hi(o4.z), d4[1]);
              d4[2] = fmaf(bf16_lo(wv), bf16_lo(o4.w), d4[2]); d4[3] = fmaf(bf16_hi(wv), bf16_hi(o4.w), d4[3]);
            }
          float dsum = (d4[0] + d4[1]) + (d4[2] + d4[3]);
          dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
          dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
          dsum += __shfl_xor_sync(0xffffffffu, dsum, 4);
          if (p8 == 0) {
            xd[rr] = dsum;
            if (t * 128 + rr < g.S) g.delta[(int64_t)w * g.S + t * 128 + rr] = dsum;
          }
        }
        if (half == 0) {                                 // dO row -> TMEM (rows beyond S were zero-filled by TMA)
          row_to_tmem<D>(pool + (r_do % NP) * STG, BCH, lrow, t_lane + DOA_COL, [&](int, const uint32_t (&)[16], int) {});
        } else {
          mbar_wait(full(r_q % NP), (r_q / NP) & 1u);
          row_to_tmem<D>(pool + (r_q % NP) * STG, BCH, lrow, t_lane + QA_COL, [&](int c, const uint32_t (&r)[16], int n) {
            if (MODE == VG_ATTN_L2) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (i < n) { const float a = bf16_lo(r[i]), bb = bf16_hi(r[i]); qq = fmaf(a, a, qq); qq = fmaf(bb, bb, qq); }
            }
          });
          if (MODE == VG_ATTN_L2) xq[row] = qq;
        }
        if (row_on) lse2 = __ldg(g.lse + (int64_t)w * g.S + row_g) * LOG2E;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(a_ready);                            // this thread's operand row is in TMEM
        named_bar(3 + hf, 128);                          // the 128 threads that read this half's dO and Q stages are done with them ...
        if (lrow == 0 && half == 0) { mbar_arrive(empty(r_do % NP)); mbar_arrive(empty(r_q % NP)); }   // ... which go back to the pool
        {                                                // next tile's O row and lse entry: into L2 (see the producer's prefetch)
          const bool last_t = t + 1 == g.n_t;
          const int wn = last_t ? w + (int)gridDim.x : w, rn = (last_t ? 0 : t + 1) * 128 + row;
          if (half == 0 && wn < total && rn < g.S && !g.no_prefetch) {
            const bf16* op = g.o + ((int64_t)(wn / g.H) * g.S + rn) * g.ldo + (wn % g.H) * D;
#pragma unroll
            for (int c = 0; c < D; c += 64) prefetch_l2(op + c);
            if ((row & 31) == 0) prefetch_l2(g.lse + (int64_t)wn * g.S + rn);
          }
        }
        named_bar(2, NCOMP_BWD);                         // delta and |q|^2 of every row are in shared memory (and, L2, the key norms)
        delta = xd[row];
        if (MODE == VG_ATTN_L2) qq = xq[row];
        tr(51);
        float gsum = 0.f;
        for (int j = 0; j < g.n_b; ++j) {
          const uint32_t c = bc0 + j;
          const int buf = c & 1, nk = min(64, g.NK - 64 * j), c0 = 64 * j + 32 * half;
          const int mine = min(32, nk - 32 * half);      // this thread's keys of the block: 32, 16 or none
          uint32_t s0[32], p0[32], pk[16];
          mbar_wait(sdp_full, c & 1u);
          tr(60 + j);
          tc_fence_after();
          if (warp_on && mine > 0) {                     // the thread's part into registers, then the TMEM pair is released at once
            if (mine == 32) { tmem_ld32_nowait(t_lane + S_COL + 32 * half, s0); tmem_ld32_nowait(t_lane + DP_COL + 32 * half, p0); }
            else { tmem_ld16p(t_lane + S_COL + 32 * half, s0); tmem_ld16p(t_lane + DP_COL + 32 * half, p0); }
            tmem_ld_wait();
          }
          tc_fence_before();
          mbar_arrive(sdp_free);                         // the next S / dP pair is computed while this one is exponentiated
          mbar_wait(ds_empty(buf), ((c >> 1) & 1u) ^ 1u);           // dQ MMA of block j-2 has consumed this dS buffer
          tr(70 + j);
          if (warp_on && mine > 0) {
            const uint32_t tile = ds_t + buf * QCH;
            if (mine == 32) {
              if (c0 + 32 <= g.S) dq_group<32, false, MODE>(s0, p0, c0, g.S, qq, kn, sc2, lse2, delta, g.scale, gsum, pk);
              else dq_group<32, true, MODE>(s0, p0, c0, g.S, qq, kn, sc2, lse2, delta, g.scale, gsum, pk);
              put_row<32>(tile, row, 32 * half, pk);
            } else {
              dq_group<16, true, MODE>(s0, p0, c0, g.S, qq, kn, sc2, lse2, delta, g.scale, gsum, pk);
              put_row<16>(tile, row, 32 * half, pk);
            }
          }
          fence_async_smem();
          mbar_arrive(ds_full(buf));
          tr(80 + j);
        }
        bc0 += g.n_b;
        if (MODE == VG_ATTN_L2) {                        // rowsum(G) of a row = the sum of its two threads' halves
          xq[128 + 128 * half + row] = gsum;
          named_bar(2, NCOMP_BWD);
          gsum += xq[128 + 128 * (half ^ 1) + row];
        }
        // ---- drain dQ: TMEM -> (L2: rowsum(G) q - G K) -> bf16 -> this row's half of the staging tile (two pool stages) -> the
        //      producer warp stores it
        mbar_wait(dq_full, tpar);                       // every MMA of the tile is complete (the TMEM operand rows may be replaced)
        tr(55);
        tc_fence_after();
        const uint32_t r_st = rq0 + R - 2 + hf;
        mbar_wait(full(r_st % NP), (r_st / NP) & 1u);   // this half's staging stage has left its previous user
        const uint32_t stg = pool + (r_st % NP) * STG;
        if (warp_on) {
#pragma unroll
          for (int k = 0; k < NCH; ++k) {
            if ((k < CH0) != (half == 0)) continue;
            const int c = 32 * k;
            uint32_t v[32], qv[16];
            const int n = (D - c) >= 32 ? 32 : 16;
            if (MODE == VG_ATTN_L2) {                    // the q row is still in TMEM (operand columns, two bf16 per column)
              if (n == 32) tmem_ld16p(t_lane + (uint32_t)(QA_COL + (c >> 1)), qv);
              else asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                : "=r"(qv[0]), "=r"(qv[1]), "=r"(qv[2]), "=r"(qv[3]), "=r"(qv[4]), "=r"(qv[5]), "=r"(qv[6]), "=r"(qv[7])
                                : "r"(t_lane + (uint32_t)(QA_COL + (c >> 1))));
            }
            tmem_ldn(t_lane + (uint32_t)(DQ_COL + c), v, n);
            float o[32];
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              o[jj] = __uint_as_float(v[jj]);
              if (MODE == VG_ATTN_L2 && jj < n) o[jj] = fmaf(gsum, (jj & 1) ? bf16_hi(qv[jj >> 1]) : bf16_lo(qv[jj >> 1]), -o[jj]);
            }
            stg_write<D, 64>(stg, lrow, c, o, n);
          }
        }
        tc_fence_before();
        mbar_arrive(dq_free);
        fence_async_smem();
        mbar_arrive(stg_full);
        tr(56);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_free512(tmem); }
}

// ================================================================================================ backward: dK, dV
template <int D> struct DkvCfg {
  static constexpr int DC = (D + 63) / 64, KS = D / 16;
  static constexpr int NS = D > 128 ? 2 : 3;                 // ring stages of one (Q block, dO block) pair
  static constexpr int DK_COL = 128, DV_COL = 128 + D;       // S^T/dP^T buffers in columns [0, 64) and [64, 128)
  static constexpr int K_OFF = 0, V_OFF = DC * QCH, QR_OFF = 2 * DC * QCH, DOR_OFF = QR_OFF + NS * DC * BCH;
  static constexpr int PT_OFF = DOR_OFF + NS * DC * BCH, DST_OFF = PT_OFF + QCH;
  static constexpr int LSE_OFF = DST_OFF + QCH, DEL_OFF = LSE_OFF + MAXNK * 4, QN_OFF = DEL_OFF + MAXNK * 4;
  static constexpr int NBAR = 8 + 2 * NS + 7;
  static constexpr int smem(int mode) { return (mode == VG_ATTN_L2 ? QN_OFF + MAXNK * 4 + 512 : QN_OFF) + NBAR * 8 + 16; }   // L2: |q|^2 table + 128 floats
  static_assert(NS * DC * BCH >= 128 * D * 2, "the dK / dV staging tiles alias the Q and dO rings");
};

template <int D, int MODE>
__global__ void __launch_bounds__(NTHREADS_BWD, 1)
attn_bwd_dkv_mt_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                       const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                       const __grid_constant__ OutMaps map_dk, const __grid_constant__ OutMaps map_dv, const MtGeo g) {
  using C = DkvCfg<D>;
  constexpr int DC = C::DC, KS = C::KS, NS = C::NS, DK_COL = C::DK_COL, DV_COL = C::DV_COL;
  constexpr int BAR_OFF = MODE == VG_ATTN_L2 ? C::QN_OFF + MAXNK * 4 + 512 : C::QN_OFF;
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  const uint32_t base = smem_u32(smem_dyn);
  const uint32_t k_t = base + C::K_OFF, v_t = base + C::V_OFF, qr = base + C::QR_OFF, dor = base + C::DOR_OFF;
  const uint32_t pt_t = base + C::PT_OFF, dst_t = base + C::DST_OFF, bar = base + BAR_OFF;
  float* lse_s = reinterpret_cast<float*>(smem_dyn + C::LSE_OFF);
  float* del_s = reinterpret_cast<float*>(smem_dyn + C::DEL_OFF);
  float* qn_s = reinterpret_cast<float*>(smem_dyn + C::QN_OFF);      // L2 mode only
  float* gs_s = qn_s + MAXNK;                                         // L2 mode only: colsum(G) exchange between a row's two threads
  const uint32_t kvt_full = bar, kvt_empty = bar + 8, out_full = bar + 16, out_free = bar + 24;
  const uint32_t pt_full = bar + 32, pt_empty = bar + 40, dst_full = bar + 48, dst_empty = bar + 56;
  auto qdo_full = [&](int i) { return bar + 8u * (8 + i); };
  auto qdo_empty = [&](int i) { return bar + 8u * (8 + NS + i); };
  auto st_full = [&](int i) { return bar + 8u * (8 + 2 * NS + i); };
  auto st_free = [&](int i) { return bar + 8u * (10 + 2 * NS + i); };
  auto dpt_full = [&](int i) { return bar + 8u * (12 + 2 * NS + i); };
  const uint32_t stg_full = bar + 8u * (14 + 2 * NS);
  const uint32_t tmem_slot = bar + 8u * C::NBAR;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_dyn + BAR_OFF + 8 * C::NBAR);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if (base & 1023u) { printf("attention_mt: dynamic smem base not 1024-aligned\n"); __trap(); }
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_do) : "memory");
    mbar_init(kvt_full, 1); mbar_init(kvt_empty, 1); mbar_init(out_full, 1); mbar_init(out_free, NCOMP_BWD); mbar_init(stg_full, NCOMP_BWD);
    mbar_init(pt_full, NCOMP_BWD); mbar_init(pt_empty, 1); mbar_init(dst_full, NCOMP_BWD); mbar_init(dst_empty, 1);
    for (int i = 0; i < NS; ++i) { mbar_init(qdo_full(i), 1); mbar_init(qdo_empty(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(st_full(i), 1); mbar_init(st_free(i), NCOMP_BWD); mbar_init(dpt_full(i), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc512(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  pdl_trigger();
  pdl_wait();
  const int total = g.B * g.H;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t bc = 0, ic = 0;                       // query-block counter, item (key tile) counter
      Tracer tr(g.trace, 6);
      int pend_w = -1, pend_t = 0;                   // item whose dK / dV the compute threads stage over the (then idle) Q / dO ring
      auto store_pending = [&]() {
        if (pend_w < 0) return;
        mbar_wait(stg_full, (ic - 1) & 1u);
        stg_store<D>(map_dk, qr, (pend_w % g.H) * D, pend_t * 128, pend_w / g.H);
        stg_store<D>(map_dv, dor, (pend_w % g.H) * D, pend_t * 128, pend_w / g.H);
        tma_commit();
        tma_wait_read();
        pend_w = -1;
      };
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int b = w / g.H, col0 = (w % g.H) * D;
        for (int t = 0; t < g.n_t; ++t, ++ic) {
          mbar_wait(kvt_empty, (ic & 1u) ^ 1u);        // every MMA of the previous item is complete: its K, V tiles are dead ...
          tr(1);
          mbar_expect_tx(kvt_full, 2 * DC * QCH);      // ... and the next ones are fetched while its dK / dV are drained and stored
#pragma unroll
          for (int c = 0; c < DC; ++c) {
            tma_load_3d(k_t + c * QCH, &map_k, kvt_full, col0 + 64 * c, t * 128, b);
            tma_load_3d(v_t + c * QCH, &map_v, kvt_full, col0 + 64 * c, t * 128, b);
          }
          store_pending();                             // the ring is refilled only after the staged outputs have been read from it
          {
            // the CTAs run in lock step, so first-touch loads reach HBM as bursts: ask for the next item's K / V tiles (and, on the
            // last key tile of a (b, h), for the next (b, h)'s Q / dO blocks) in L2 one item ahead
            const bool last_t = t + 1 == g.n_t;
            const int wn = last_t ? w + (int)gridDim.x : w, tn = last_t ? 0 : t + 1;
            if (wn < total) {
              const int bn = wn / g.H, cn = (wn % g.H) * D;
#pragma unroll
              for (int c = 0; c < DC; ++c) {
                tma_prefetch_3d(&map_k, cn + 64 * c, tn * 128, bn);
                tma_prefetch_3d(&map_v, cn + 64 * c, tn * 128, bn);
              }
              if (last_t)
                for (int i = 0; i < g.n_b; ++i)
#pragma unroll
                  for (int c = 0; c < DC; ++c) {
                    tma_prefetch_3d(&map_q, cn + 64 * c, i * 64, bn);
                    tma_prefetch_3d(&map_do, cn + 64 * c, i * 64, bn);
                  }
            }
          }
          for (int i = 0; i < g.n_b; ++i, ++bc) {
            const int st = bc % NS;
            mbar_wait(qdo_empty(st), ((bc / NS) & 1u) ^ 1u);
            tr(100 + i);
            mbar_expect_tx(qdo_full(st), 2 * DC * BCH);
#pragma unroll
            for (int c = 0; c < DC; ++c) {
              tma_load_3d(qr + (st * DC + c) * BCH, &map_q, qdo_full(st), col0 + 64 * c, i * 64, b);
              tma_load_3d(dor + (st * DC + c) * BCH, &map_do, qdo_full(st), col0 + 64 * c, i * 64, b);
            }
          }
          pend_w = w; pend_t = t;
        }
      }
      store_pending();
      tma_wait_all();
    }
  } else if (warp == 1) {
    // MMA issue: the whole warp runs the loop (warp-uniform control flow and operands), one elected lane issues
    {
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t idesc_kv = make_idesc(128, D, 0, 1);            // dV = P^T dO, dK = dS^T Q : A K-major, B MN-major
      const SDesc ktd = sdesc_k(k_t), vtd = sdesc_k(v_t), ptd = sdesc_k(pt_t), dstd = sdesc_k(dst_t);
      uint32_t bc0 = 0, ic = 0;
      Tracer tr(lane == 0 ? g.trace : nullptr, 7);
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        for (int t = 0; t < g.n_t; ++t, ++ic) {
          const uint32_t ipar = ic & 1u;
          mbar_wait(kvt_full, ipar);
          tr(10);
          tc_fence_after();
          // event loop over three kinds of work, oldest dependency first: dK(i) once dS^T(i) is written (it releases the ring
          // stage), dP^T(i) + dV(i) once P^T(i) is written, S^T(i) once its Q / dO block has landed and its TMEM buffer is free
          int is = 0, ip = 0, ik = 0;                                 // next block for S^T, for dP^T + dV, for dK
          while (ik < g.n_b) {
            if (ik < ip) {
              const uint32_t c = bc0 + ik;
              const int st = c % NS, ni = min(64, g.NK - 64 * ik);
              if (mbar_test(dst_full, c & 1u)) {
                tr(45 + ik);
                tc_fence_after();
                const SDesc qd = sdesc_mn(qr + st * DC * BCH, BCH);
                if (elect_one()) {
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk)
                    if (kk < ni / 16) tc_mma_d(tm + DK_COL, dstd, kk * 2u, qd, kk * 128u, idesc_kv, (ik > 0 || kk > 0) ? 1u : 0u);
                  tc_commit(dst_empty);
                  tc_commit(qdo_empty(st));
                }
                __syncwarp();
                ++ik;
                continue;
              }
            }
            if (ip < is) {
              const uint32_t c = bc0 + ip;
              const int buf = c & 1, st = c % NS, ni = min(64, g.NK - 64 * ip);
              if (mbar_test(pt_full, c & 1u) && (ip > 0 || mbar_test(out_free, ipar ^ 1u))) {   // P^T in smem; S^T fully read
                tr(40 + ip);
                tc_fence_after();
                const uint32_t idesc_s = make_idesc(128, ni, 0, 0);
                const SDesc dok = sdesc_k(dor + st * DC * BCH), dom = sdesc_mn(dor + st * DC * BCH, BCH);
                if (elect_one()) {
#pragma unroll
                  for (int k = 0; k < KS; ++k) tc_mma_d(tm + (uint32_t)(64 * buf), vtd, kstep_off(k, QCH), dok, kstep_off(k, BCH), idesc_s, k > 0);
                  tc_commit(dpt_full(buf));
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk)
                    if (kk < ni / 16) tc_mma_d(tm + DV_COL, ptd, kk * 2u, dom, kk * 128u, idesc_kv, (ip > 0 || kk > 0) ? 1u : 0u);
                  tc_commit(pt_empty);
                }
                __syncwarp();
                ++ip;
                continue;
              }
            }
            if (is < g.n_b && is < ip + 2) {
              const uint32_t c = bc0 + is;
              const int buf = c & 1, st = c % NS, ni = min(64, g.NK - 64 * is);
              if (mbar_test(qdo_full(st), (c / NS) & 1u) && mbar_test(st_free(buf), ((c >> 1) & 1u) ^ 1u)) {
                tr(20 + is);
                tc_fence_after();
                const uint32_t idesc_s = make_idesc(128, ni, 0, 0);
                const SDesc qk = sdesc_k(qr + st * DC * BCH);
                if (elect_one()) {
#pragma unroll
                  for (int k = 0; k < KS; ++k) tc_mma_d(tm + (uint32_t)(64 * buf), ktd, kstep_off(k, QCH), qk, kstep_off(k, BCH), idesc_s, k > 0);
                  tc_commit(st_full(buf));
                }
                __syncwarp();
                ++is;
              }
            }
          }
          if (elect_one()) { tc_commit(out_full); tc_commit(kvt_empty); }
          __syncwarp();
          tr(14);
          bc0 += g.n_b;
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- two threads per key row: each owns one 32-query half of
    // every 64-query block (two warps per scheduler hide each other's TMEM / MUFU / shared-memory latencies)
    const int quad = warp & 3, half = (warp - 2) >> 2, row = quad * 32 + lane, tid = threadIdx.x - 64;
    const uint32_t t_lane = tmem + ((uint32_t)(quad * 32) << 16);
    const float sc2 = g.scale * LOG2E;
    uint32_t bc0 = 0, ic = 0;
    Tracer tr(threadIdx.x == 128 ? g.trace : nullptr, 8);        // warp 4: quadrant 0, half 0
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int b = w / g.H, h = w % g.H, col0 = h * D;
      tr(50);
      // per-query statistics of this problem -> smem: lse * log2e (+huge for padding queries: P = 0), delta, (L2) |q|^2
      named_bar(2, NCOMP_BWD);
      for (int q = tid; q < g.n_b * 64 && q < MAXNK; q += NCOMP_BWD) {
        const bool on = q < g.S;
        lse_s[q] = on ? __ldg(g.lse + (int64_t)w * g.S + q) * LOG2E : 1e30f;
        del_s[q] = on ? __ldg(g.delta + (int64_t)w * g.S + q) : 0.f;
        if (MODE == VG_ATTN_L2) qn_s[q] = on ? row_sqnorm<D>(g.q + ((int64_t)b * g.S + q) * g.ld + col0) : 0.f;
      }
      named_bar(2, NCOMP_BWD);
      for (int t = 0; t < g.n_t; ++t, ++ic) {
        const uint32_t ipar = ic & 1u;
        const int key_g = t * 128 + row;
        const bool warp_on = t * 128 + quad * 32 < g.S;
        float kk2 = 0.f, gsum = 0.f;
        if (MODE == VG_ATTN_L2 && key_g < g.S) kk2 = row_sqnorm<D>(g.k + ((int64_t)b * g.S + key_g) * g.ld + col0);
        for (int i = 0; i < g.n_b; ++i) {
          const uint32_t c = bc0 + i;
          const int buf = c & 1, ni = min(64, g.NK - 64 * i);
          uint32_t f[16];                               // P * scale (dot) or P * scale / dist (L2) of this thread's 32 queries, packed bf16
          // stage A: S^T -> P^T (smem, A operand of dV) and the factor f kept in registers
          mbar_wait(st_full(buf), (c >> 1) & 1u);
          tr(60 + i);
          tc_fence_after();
          mbar_wait(pt_empty, (c & 1u) ^ 1u);             // dV MMA of the previous block has consumed the P^T buffer
          tr(65 + i);
          const int cc = 32 * half, q0 = 64 * i + cc;
          if (warp_on && cc < ni) {
            uint32_t sv[32], pk[16];
            if (ni - cc >= 32) {
              tmem_ld32(t_lane + (uint32_t)(64 * buf + cc), sv);
              dkv_stage_a<32, MODE>(sv, q0, kk2, qn_s, lse_s, sc2, g.scale, pk, f);
              put_row<32>(pt_t, row, cc, pk);
            } else {
              tmem_ld16p(t_lane + (uint32_t)(64 * buf + cc), sv);
              tmem_ld_wait();
              dkv_stage_a<16, MODE>(sv, q0, kk2, qn_s, lse_s, sc2, g.scale, pk, f);
              put_row<16>(pt_t, row, cc, pk);
            }
          }
          tc_fence_before();
          fence_async_smem();
          mbar_arrive(pt_full);
          tr(70 + i);
          // stage B: dP^T -> dS^T = f (dP^T - delta) -> smem (A operand of dK)
          mbar_wait(dpt_full(buf), (c >> 1) & 1u);
          tr(75 + i);
          tc_fence_after();
          mbar_wait(dst_empty, (c & 1u) ^ 1u);
          tr(80 + i);
          if (warp_on && cc < ni) {
            uint32_t dv[32], pk[16];
            if (ni - cc >= 32) {
              tmem_ld32(t_lane + (uint32_t)(64 * buf + cc), dv);
              dkv_stage_b<32, MODE>(dv, q0, del_s, f, gsum, pk);
              put_row<32>(dst_t, row, cc, pk);
            } else {
              tmem_ld16p(t_lane + (uint32_t)(64 * buf + cc), dv);
              tmem_ld_wait();
              dkv_stage_b<16, MODE>(dv, q0, del_s, f, gsum, pk);
              put_row<16>(dst_t, row, cc, pk);
            }
          }
          tc_fence_before();
          mbar_arrive(st_free(buf));
          fence_async_smem();
          mbar_arrive(dst_full);
          tr(85 + i);
        }
        bc0 += g.n_b;
        // ---- drain dK, dV: TMEM -> (L2: colsum(G) k - G^T Q) -> bf16 -> staging over the K and V tiles -> TMA stores
        mbar_wait(out_full, ipar);
        tr(55);
        tc_fence_after();
        const bf16* kg = g.k + ((int64_t)b * g.S + min(key_g, g.S - 1)) * g.ld + col0;    // L2: this key's row (the smem K tile is being refilled)
        if (MODE == VG_ATTN_L2) {                        // colsum(G) of a key row = the sum of its two threads' halves
          if (half == 1) gs_s[row] = gsum;
          named_bar(2, NCOMP_BWD);
          gsum += gs_s[row];
          named_bar(2, NCOMP_BWD);
        }
        if (warp_on && half == 0) {                      // dK by the first thread of the row, dV by the second
#pragma unroll
          for (int c = 0; c < D; c += 32) {
            uint32_t v[32];
            const int n = (D - c) >= 32 ? 32 : 16;
            uint4 kv4[4];
            if (MODE == VG_ATTN_L2) {
#pragma unroll
              for (int i = 0; i < 4; ++i) if (8 * i < n) kv4[i] = __ldg(reinterpret_cast<const uint4*>(kg + c + 8 * i));
            }
            tmem_ldn(t_lane + (uint32_t)(DK_COL + c), v, n);
            float o[32];
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              o[jj] = __uint_as_float(v[jj]);
              if (MODE == VG_ATTN_L2 && jj < n) {
                const uint4 q4 = kv4[jj >> 3];
                const uint32_t u = ((jj >> 1) & 3) == 0 ? q4.x : ((jj >> 1) & 3) == 1 ? q4.y : ((jj >> 1) & 3) == 2 ? q4.z : q4.w;
                o[jj] = fmaf(gsum, (jj & 1) ? bf16_hi(u) : bf16_lo(u), -o[jj]);
              }
            }
            stg_write<D>(qr, row, c, o, n);
          }
        }
        if (warp_on && half == 1) {
#pragma unroll
          for (int c = 0; c < D; c += 32) {
            uint32_t v[32];
            const int n = (D - c) >= 32 ? 32 : 16;
            tmem_ldn(t_lane + (uint32_t)(DV_COL + c), v, n);
            float o[32];
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) o[jj] = __uint_as_float(v[jj]);
            stg_write<D>(dor, row, c, o, n);
          }
        }
        tc_fence_before();
        mbar_arrive(out_free);
        fence_async_smem();
        mbar_arrive(stg_full);
        tr(56);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_free512(tmem); }
}

// ================================================================================================ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(p);
    return (EncodeTiledFn) nullptr;
  }();
  return fn;
}
// bf16 tensor viewed as [B, S, cols] with row pitch ld elements; box {box_cols, box_rows, 1}; swizzle span = box_cols * 2 bytes
int make_map(CUtensorMap* map, const void* ptr, int B, int S, int cols, int64_t ld, int box_cols, int box_rows) {
  EncodeTiledFn enc = get_encode();
  VG_REQUIRE(enc != nullptr, VG_ERR_LAUNCH, "attention_mt: cuTensorMapEncodeTiled not available");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)S, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)S * ld * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VG_REQUIRE(r == CUDA_SUCCESS, VG_ERR_LAUNCH, "attention_mt: cuTensorMapEncodeTiled failed (%d) B=%d S=%d cols=%d ld=%lld box=%dx%d", (int)r, B, S,
             cols, (long long)ld, box_cols, box_rows);
  return VG_OK;
}
int make_out_maps(OutMaps* m, const void* ptr, int B, int S, int cols, int64_t ld, int d, int rows = 128) {
  int rc;
  memset(m, 0, sizeof(*m));
  if (d >= 64 && (rc = make_map(&m->m64, ptr, B, S, cols, ld, 64, rows))) return rc;
  if ((d % 64) >= 32 && (rc = make_map(&m->m32, ptr, B, S, cols, ld, 32, rows))) return rc;
  if ((d % 32) >= 16 && (rc = make_map(&m->m16, ptr, B, S, cols, ld, 16, rows))) return rc;
  return VG_OK;
}

unsigned long long* g_mt_trace = nullptr;

MtGeo make_geo(int B, int H, int S, int64_t ld, int64_t ldo, float scale, const void* q, const void* k, const void* o, float* lse, float* delta) {
  MtGeo g;
  g.trace = g_mt_trace;
  g.B = B; g.H = H; g.S = S; g.NK = (S + 15) / 16 * 16; g.n_t = (S + 127) / 128; g.n_b = (g.NK + 63) / 64; g.ld = ld; g.ldo = ldo;
  g.scale = scale; g.lse = lse; g.delta = delta;
  g.q = static_cast<const bf16*>(q); g.k = static_cast<const bf16*>(k); g.o = static_cast<const bf16*>(o);
  g.dq = g.dk = g.dv = nullptr; g.ldd = 0;
  static const int env_skew = [] { const char* e = getenv("VG_ATTN_SKEW"); return e ? atoi(e) : 0; }();      // experiment knobs, read once
  static const int env_nopf = [] { const char* e = getenv("VG_ATTN_NOPF"); return e ? atoi(e) : 0; }();
  g.skew_ns = env_skew; g.no_prefetch = env_nopf;
  return g;
}

template <typename K>
int set_smem_once(K kern, int bytes) {      // callers keep the result in a function-local static: one-time, thread-safe
  VG_REQUIRE(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess, VG_ERR_LAUNCH,
             "attention_mt: cudaFuncSetAttribute(%d B) failed", bytes);
  return VG_OK;
}

template <int D, int MODE>
int launch_fwd(const CUtensorMap& mq, const CUtensorMap& mk64, const CUtensorMap& mk16, const CUtensorMap& mv, const OutMaps& mo, const MtGeo& g,
               cudaStream_t st) {
  static const int attr_rc = set_smem_once(attn_fwd_mt_kernel<D, MODE>, FwdCfg<D>::SMEM);
  int rc = attr_rc;
  if (rc) return rc;
  const int grid = min(g.B * g.H, num_sms());
  launch_pdl(attn_fwd_mt_kernel<D, MODE>, dim3(grid), dim3(NTHREADS), (size_t)FwdCfg<D>::SMEM, st, mq, mk64, mk16, mv, mo, g);
  return check_launch("attention_fwd_mt");
}
template <int D, int MODE>
int launch_bwd(const CUtensorMap& mq128, const CUtensorMap& mdo128, const CUtensorMap& mk64, const CUtensorMap& mv64, const OutMaps& mdq,
               const CUtensorMap& mk128, const CUtensorMap& mv128, const CUtensorMap& mq64, const CUtensorMap& mdo64, const OutMaps& mdk,
               const OutMaps& mdv, const MtGeo& g, cudaStream_t st) {
  static const int attr_rc1 = set_smem_once(attn_bwd_dq_mt_kernel<D, MODE>, DqCfg<D>::SMEM);
  static const int attr_rc2 = set_smem_once(attn_bwd_dkv_mt_kernel<D, MODE>, DkvCfg<D>::smem(MODE));
  int rc = attr_rc1 ? attr_rc1 : attr_rc2;
  if (rc) return rc;
  const int grid = min(g.B * g.H, num_sms());
  launch_pdl(attn_bwd_dq_mt_kernel<D, MODE>, dim3(grid), dim3(NTHREADS_BWD), (size_t)DqCfg<D>::SMEM, st, mq64, mdo64, mk64, mv64, mdq, g);
  rc = check_launch("attention_bwd_dq_mt");
  if (rc) return rc;
  launch_pdl(attn_bwd_dkv_mt_kernel<D, MODE>, dim3(grid), dim3(NTHREADS_BWD), (size_t)DkvCfg<D>::smem(MODE), st, mk128, mv128, mq64, mdo64, mdk, mdv, g);
  return check_launch("attention_bwd_dkv_mt");
}

#define VG_MT_DISPATCH(d_, mode_, CALL)                                                                            \
  do {                                                                                                             \
    if (mode_ == VG_ATTN_L2) {                                                                                     \
      constexpr int MODE = VG_ATTN_L2;                                                                             \
      if (d_ == 96) { constexpr int D = 96; CALL; } else { constexpr int D = 112; CALL; }                          \
    } else {                                                                                                       \
      constexpr int MODE = VG_ATTN_DOT;                                                                            \
      if (d_ == 96) { constexpr int D = 96; CALL; } else if (d_ == 112) { constexpr int D = 112; CALL; } else { constexpr int D = 192; CALL; } \
    }                                                                                                              \
  } while (0)

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

void attention_mt_set_trace(unsigned long long* p) { g_mt_trace = p; }

// d = 192 with L2 scores is not instantiated: the dK/dV kernel's smem budget has no room for the |q|^2 table, and no
// reference configuration uses it (v1's L2 heads are 108 -> 112 wide).
bool attention_mt_supported(int dtype, int mode, int B, int H, int S, int d, const void* q, const void* k, const void* v,
                            int64_t ld_qkv, const void* o, int64_t ld_o) {
  static const int sm100 = vg_device_is_sm100();
  static const int forced_off = [] { const char* e = getenv("VG_ATTN_PATH"); return (e && !strcmp(e, "simt")) ? 1 : 0; }();
  if (!sm100 || forced_off) return false;
  if (dtype != VG_BF16) return false;
  if (!(d == 96 || d == 112 || (d == 192 && mode == VG_ATTN_DOT))) return false;
  if (S < 1 || S > MAXNK || B < 1 || H < 1) return false;
  if (ld_qkv % 8 || ld_o % 8) return false;
  return aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o);
}

int attention_fwd_mt(int mode, int B, int H, int S, int d, const void* q, const void* k, const void* v, int64_t ld, void* o, int64_t ldo,
                     float* lse, float scale, cudaStream_t st) {
  const int cols = H * d;
  CUtensorMap mq, mk64, mk16, mv;
  OutMaps mo;
  int rc;
  if ((rc = make_map(&mq, q, B, S, cols, ld, 64, 128))) return rc;
  if ((rc = make_map(&mk64, k, B, S, cols, ld, 64, 64))) return rc;
  if ((rc = make_map(&mk16, k, B, S, cols, ld, 64, 16))) return rc;
  if ((rc = make_map(&mv, v, B, S, cols, ld, 64, 64))) return rc;
  if ((rc = make_out_maps(&mo, o, B, S, cols, ldo, d))) return rc;
  const MtGeo g = make_geo(B, H, S, ld, ldo, scale, q, k, o, lse, nullptr);
  VG_MT_DISPATCH(d, mode, (rc = launch_fwd<D, MODE>(mq, mk64, mk16, mv, mo, g, st)));
  return rc;
}

int attention_bwd_mt(int mode, int B, int H, int S, int d, const void* q, const void* k, const void* v, int64_t ld, const void* o,
                     const void* d_o, int64_t ldo, const float* lse, void* dq, void* dk, void* dv, int64_t ldd, float scale, float* delta,
                     cudaStream_t st) {
  const int cols = H * d;
  CUtensorMap mq128, mdo128, mk64, mv64, mk128, mv128, mq64, mdo64;
  OutMaps mdq, mdk, mdv;
  int rc;
  if ((rc = make_map(&mq128, q, B, S, cols, ld, 64, 128))) return rc;
  if ((rc = make_map(&mdo128, d_o, B, S, cols, ldo, 64, 128))) return rc;
  if ((rc = make_map(&mk64, k, B, S, cols, ld, 64, 64))) return rc;
  if ((rc = make_map(&mv64, v, B, S, cols, ld, 64, 64))) return rc;
  if ((rc = make_map(&mk128, k, B, S, cols, ld, 64, 128))) return rc;
  if ((rc = make_map(&mv128, v, B, S, cols, ld, 64, 128))) return rc;
  if ((rc = make_map(&mq64, q, B, S, cols, ld, 64, 64))) return rc;
  if ((rc = make_map(&mdo64, d_o, B, S, cols, ldo, 64, 64))) return rc;
  if ((rc = make_out_maps(&mdq, dq, B, S, cols, ldd, d, 64))) return rc;
  if ((rc = make_out_maps(&mdk, dk, B, S, cols, ldd, d))) return rc;
  if ((rc = make_out_maps(&mdv, dv, B, S, cols, ldd, d))) return rc;
  MtGeo g = make_geo(B, H, S, ld, ldo, scale, q, k, o, const_cast<float*>(lse), delta);
  g.dq = static_cast<bf16*>(dq); g.dk = static_cast<bf16*>(dk); g.dv = static_cast<bf16*>(dv); g.ldd = ldd;
  VG_MT_DISPATCH(d, mode, (rc = launch_bwd<D, MODE>(mq128, mdo128, mk64, mv64, mdq, mk128, mv128, mq64, mdo64, mdk, mdv, g, st)));
  return rc;
}

}  // namespace vg
