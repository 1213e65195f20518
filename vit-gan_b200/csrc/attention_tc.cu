// attention_tc.cu -- dot-product multi-head self-attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA),
// forward and backward, for the single-tile regime of the ViT-GAN configs: S <= 128 tokens, head dim 32 or 64,
// bf16 operands, fp32 accumulation and softmax statistics.   (src/v2/modules.py:142-159 and its autograd backward.)
//
// Work item = (batch element b, group of 128 feature columns = 4 heads of 32 or 2 heads of 64).  All operands of
// an item are brought in by TMA as [rows x 128 B] SWIZZLE_128B tiles straight from the fused QKV projection
// output (3-D tensor maps [B, S, cols]: rows >= S are out of bounds -> zero filled on load, clipped on store), so
// there is no head split/merge copy and no S x S tensor in HBM.  One smem tile serves several operand roles:
//     forward :  S_h = Q_h K_h^T   (A = Q  K-major,  B = K  K-major)          -> TMEM, 4 heads side by side
//                softmax rows in registers (thread = query row), P_h -> bf16 -> swizzled smem
//                O_h = P_h V_h     (A = P  K-major,  B = V  MN-major)         -> TMEM -> * 1/l -> smem -> TMA store
//     backward:  S_h = Q_h K_h^T, dP_h = dO_h V_h^T (B = V K-major)           -> TMEM (double buffered per head)
//                P = exp(S*scale - lse); delta = rowsum(P*dP) (== rowsum(dO*O)); dS = P*(dP - delta)*scale
//                dV_h = P_h^T dO_h   (A = P  MN-major, B = dO MN-major)
//                dK_h = dS_h^T Q_h   (A = dS MN-major, B = Q  MN-major)
//                dQ_h = dS_h K_h     (A = dS K-major,  B = K  MN-major)        -> TMEM -> smem -> TMA store into dqkv
// Warp roles: warp 0 TMA loads, warp 1 MMA issue (+TMEM alloc), warps 2..5 softmax / drain (thread = row).
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace vg {
namespace {

constexpr int NTHREADS = 192;
constexpr int ROWS = 128;                 // query rows per tile (UMMA M)
constexpr int CHUNK_BYTES = ROWS * 128;   // one [128 rows x 64 bf16] swizzled tile
constexpr float LOG2E = 1.4426950408889634f;

struct Geo {
  int B, H, S, NK;          // NK = S rounded up to 16 (UMMA N / K granularity)
  int groups;               // column groups of 128 per batch element (= H * D / 128)
  float scale;              // softmax scale (applied to the raw dot products)
  float* lse;               // [B, H, S]
};

// ================================================================================================ forward
// smem: Q[2 chunks] K[2] V[2] (96 KB, single buffered; K/V tiles hold NK rows) + P[2 bufs][2 chunks] (64 KB; buffer 0 doubles
// as the output staging tile)
template <int D>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o, const Geo g) {
  constexpr int HPC = 128 / D;                 // heads per work item
  constexpr int S_STRIDE = 128;                // TMEM columns reserved per head for S (N <= 128)
  constexpr int O_COL = 2 * S_STRIDE;          // S regions are double buffered by head parity; O (128 cols) behind them
  constexpr int IN_BYTES = 6 * CHUNK_BYTES;    // Q,K,V x 2 chunks
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t in_base = base;                         // [IN_BYTES]
  const uint32_t p_base = base + IN_BYTES;               // [2][2 chunks]
  const uint32_t bar_base = p_base + 4 * CHUNK_BYTES;
  const uint32_t in_full = bar_base, in_empty = bar_base + 16;
  auto s_full = [&](int i) { return bar_base + 8u * (4 + i); };     // 2 (head parity)
  auto s_free = [&](int i) { return bar_base + 8u * (6 + i); };     // 2
  auto p_full = [&](int i) { return bar_base + 8u * (8 + i); };     // 2
  auto p_empty = [&](int i) { return bar_base + 8u * (10 + i); };   // 2
  const uint32_t o_full = bar_base + 8u * 12, o_free = bar_base + 8u * 13;
  const uint32_t tmem_slot = bar_base + 8u * 14;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - smem_u32(smem_dyn)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_full(i), 1); mbar_init(s_free(i), 128);
      mbar_init(p_full(i), 128); mbar_init(p_empty(i), 1);
    }
    mbar_init(in_full, 1); mbar_init(in_empty, 1);
    mbar_init(o_full, 1); mbar_init(o_free, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  pdl_trigger();
  pdl_wait();
  const int total = g.B * g.groups;
  const int nk_steps = g.NK / 16;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const int b = w / g.groups, col0 = (w % g.groups) * 128;
        mbar_wait(in_empty, ((uint32_t)it & 1u) ^ 1u);
        const uint32_t t = in_base;
        mbar_expect_tx(in_full, 2u * (ROWS * 128u) + 4u * ((uint32_t)g.NK * 128u));
        for (int c = 0; c < 2; ++c) {
          tma_load_3d(t + c * CHUNK_BYTES, &map_q, in_full, col0 + 64 * c, 0, b);
          tma_load_3d(t + (2 + c) * CHUNK_BYTES, &map_k, in_full, col0 + 64 * c, 0, b);
          tma_load_3d(t + (4 + c) * CHUNK_BYTES, &map_v, in_full, col0 + 64 * c, 0, b);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(ROWS, g.NK, 0, 0);      // S = Q K^T : both K-major
      const uint32_t idesc_o = make_idesc(ROWS, D, 0, 1);         // O = P V   : A K-major, B MN-major
      int it = 0;
      uint32_t hcount = 0;                                        // running head counter (parity bookkeeping)
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const uint32_t t = in_base;
        mbar_wait(in_full, (uint32_t)it & 1u);
        tc_fence_after();
        // software pipeline over heads: S(h+1) is issued before PV(h) so the softmax of h overlaps the next QK^T
        auto issue_s = [&](int h, uint32_t hc) {
          const int r = hc & 1;
          mbar_wait(s_free(r), ((hc >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t off = (uint32_t)((h * D) / 64) * CHUNK_BYTES + (uint32_t)((h * D) % 64) * 2u;
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            tc_mma(tmem + (uint32_t)(r * S_STRIDE), desc_k(t + off + k * 32u), desc_k(t + 2 * CHUNK_BYTES + off + k * 32u), idesc_s, k > 0);
          tc_commit(s_full(r));
        };
        issue_s(0, hcount);
        for (int h = 0; h < HPC; ++h) {
          const uint32_t hc = hcount + h;
          if (h + 1 < HPC) issue_s(h + 1, hc + 1);
          const int pb = hc & 1;
          if (h == 0) { mbar_wait(o_free, ((uint32_t)it & 1u) ^ 1u); }
          mbar_wait(p_full(pb), (hc >> 1) & 1u);
          tc_fence_after();
          const uint32_t pt = p_base + pb * 2 * CHUNK_BYTES;
          const uint32_t vt = t + 4 * CHUNK_BYTES + (uint32_t)((h * D) / 64) * CHUNK_BYTES + (uint32_t)((h * D) % 64) * 2u;
          for (int k = 0; k < nk_steps; ++k)      // K = keys: 16 per step; P chunk (k/4), 32 B per step; V rows 16k.. (2048 B per step)
            tc_mma(tmem + (uint32_t)(O_COL + h * D), desc_k(pt + (k >> 2) * CHUNK_BYTES + (k & 3) * 32u), desc_mn(vt + k * 2048u, CHUNK_BYTES), idesc_o, k > 0);
          tc_commit(p_empty(pb));
        }
        tc_commit(o_full);
        tc_commit(in_empty);
        hcount += HPC;
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax + output drain: thread = query row
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const float sc2 = g.scale * LOG2E;
    int it = 0;
    uint32_t hcount = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
      const int b = w / g.groups, grp = w % g.groups, col0 = grp * 128;
      float inv_l[HPC];
#pragma unroll
      for (int h = 0; h < HPC; ++h) {
        const uint32_t hc = hcount + h;
        const int r = hc & 1;
        mbar_wait(s_full(r), (hc >> 1) & 1u);
        tc_fence_after();
        // pass 1: row max over the valid keys
        float m = -INFINITY;
        for (int c = 0; c < g.NK; c += 16) {
          uint32_t v[16];
          tmem_ld16(tmem + lane_addr + (uint32_t)(r * S_STRIDE + c), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) if (c + j < g.S) m = fmaxf(m, __uint_as_float(v[j]));
        }
        // pass 2: p = exp2((s - m) * scale * log2e), row sum, bf16 P -> swizzled smem (keys >= S are written as 0)
        const int pb = hc & 1;
        mbar_wait(p_empty(pb), ((hc >> 1) & 1u) ^ 1u);
        const uint32_t pt = p_base + pb * 2 * CHUNK_BYTES;
        float l = 0.f;
        const float mb = m * sc2;
        for (int c = 0; c < g.NK; c += 16) {
          uint32_t v[16];
          tmem_ld16(tmem + lane_addr + (uint32_t)(r * S_STRIDE + c), v);
          tmem_ld_wait();
          float p[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            p[j] = (c + j < g.S) ? exp2f(fmaf(__uint_as_float(v[j]), sc2, -mb)) : 0.f;
            l += p[j];
          }
          const uint32_t tile = pt + (uint32_t)(c >> 6) * CHUNK_BYTES;
          const int c16 = (c & 63) >> 3;
          sts128(swz(tile, row, c16), pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
          sts128(swz(tile, row, c16 + 1), pack_bf16(p[8], p[9]), pack_bf16(p[10], p[11]), pack_bf16(p[12], p[13]), pack_bf16(p[14], p[15]));
        }
        tc_fence_before();
        mbar_arrive(s_free(r));                  // S_h fully read
        fence_async_smem();                      // P visible to the tensor-core (async) proxy
        mbar_arrive(p_full(pb));
        inv_l[h] = 1.0f / l;
        if (row < g.S) g.lse[((int64_t)b * g.H + grp * HPC + h) * g.S + row] = m * g.scale + __logf(l);
      }
      // ---- drain O (all heads of the item): TMEM -> *1/l -> bf16 -> swizzled staging (P buffer 0) -> TMA store
      mbar_wait(o_full, (uint32_t)it & 1u);
      tc_fence_after();
      if (threadIdx.x == 64) tma_wait_read();    // previous item's store has finished reading the staging tiles
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const uint32_t stg = p_base;               // P buffer 0: free, every PV MMA of this item has completed (o_full)
#pragma unroll
      for (int h = 0; h < HPC; ++h) {
#pragma unroll
        for (int c = 0; c < D; c += 16) {
          uint32_t v[16];
          tmem_ld16(tmem + lane_addr + (uint32_t)(O_COL + h * D + c), v);
          tmem_ld_wait();
          float o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]) * inv_l[h];
          const int col = h * D + c;
          const uint32_t tile = stg + (uint32_t)(col >> 6) * CHUNK_BYTES;
          const int c16 = (col & 63) >> 3;
          sts128(swz(tile, row, c16), pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
          sts128(swz(tile, row, c16 + 1), pack_bf16(o[8], o[9]), pack_bf16(o[10], o[11]), pack_bf16(o[12], o[13]), pack_bf16(o[14], o[15]));
        }
      }
      tc_fence_before();
      mbar_arrive(o_free);
      fence_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        tma_store_3d(&map_o, stg, col0, 0, b);
        tma_store_3d(&map_o, stg + CHUNK_BYTES, col0 + 64, 0, b);
        tma_commit();
      }
      hcount += HPC;
      // NOTE: P buffer 0 is reused by the next item's first head only after p_empty/its MMA, and its softmax write is
      // ordered behind this store's smem read by the tma_wait_read() + bar.sync at the top of the next drain ... which is
      // too late for head 0 of the next item -> wait here instead (cheap: the store reads 32 KB of smem).
      if (threadIdx.x == 64) tma_wait_read();
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    if (threadIdx.x == 64) tma_wait_read();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ================================================================================================ backward
// smem: Q[2] K[2] V[2] dO[2] chunks (single buffered, 128 KB) + P[2 chunks] + dS[2 chunks] (64 KB).  The dQ/dK/dV staging
// tiles alias P/dS: they are written after the head's last MMA has consumed P/dS and before the next head's P is produced.
template <int D>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_do,
                   const __grid_constant__ CUtensorMap map_dq, const __grid_constant__ CUtensorMap map_dk,
                   const __grid_constant__ CUtensorMap map_dv, const Geo g) {
  constexpr int HPC = 128 / D;
  // TMEM columns: [S | dP] double buffered by head parity (2 x 256), no room left for outputs at NK = 128 ->
  // S and dP use 96-column slots when NK <= 96 ... keep it simple: slots of 128 for S and dP of ONE head (256),
  // outputs dV,dK,dQ of one head behind them (3*D <= 192): 448 <= 512.  Heads are processed one at a time.
  constexpr int DP_COL = 128, DV_COL = 256, DK_COL = 256 + D, DQ_COL = 256 + 2 * D;
  constexpr int STG_BYTES = ROWS * D * 2;      // one [128 x D] bf16 staging tile, plain row-major
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t q_t = base, k_t = base + 2 * CHUNK_BYTES, v_t = base + 4 * CHUNK_BYTES, do_t = base + 6 * CHUNK_BYTES;
  const uint32_t p_t = base + 8 * CHUNK_BYTES, ds_t = base + 10 * CHUNK_BYTES;
  const uint32_t stg = p_t;                                // 3 tiles of STG_BYTES (<= 48 KB) over P | dS
  const uint32_t bar_base = base + 12 * CHUNK_BYTES;
  const uint32_t in_full = bar_base, in_empty = bar_base + 8, sdp_full = bar_base + 16, sdp_free = bar_base + 24,
                 pds_full = bar_base + 32, pds_empty = bar_base + 40, out_full = bar_base + 48, out_free = bar_base + 56;
  const uint32_t tmem_slot = bar_base + 64;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - smem_u32(smem_dyn)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    mbar_init(in_full, 1); mbar_init(in_empty, 1); mbar_init(sdp_full, 1); mbar_init(sdp_free, 128);
    mbar_init(pds_full, 128); mbar_init(pds_empty, 1); mbar_init(out_full, 1); mbar_init(out_free, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  pdl_trigger();
  pdl_wait();
  const int total = g.B * g.groups;
  const int nk_steps = g.NK / 16;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const int b = w / g.groups, col0 = (w % g.groups) * 128;
        mbar_wait(in_empty, ((uint32_t)it & 1u) ^ 1u);
        mbar_expect_tx(in_full, 4u * (ROWS * 128u) + 4u * ((uint32_t)g.NK * 128u));
        for (int c = 0; c < 2; ++c) {
          tma_load_3d(q_t + c * CHUNK_BYTES, &map_q, in_full, col0 + 64 * c, 0, b);
          tma_load_3d(do_t + c * CHUNK_BYTES, &map_do, in_full, col0 + 64 * c, 0, b);
          tma_load_3d(k_t + c * CHUNK_BYTES, &map_k, in_full, col0 + 64 * c, 0, b);
          tma_load_3d(v_t + c * CHUNK_BYTES, &map_v, in_full, col0 + 64 * c, 0, b);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(ROWS, g.NK, 0, 0);     // S, dP: [q x keys], A,B K-major
      const uint32_t idesc_kv = make_idesc(ROWS, D, 1, 1);       // dV, dK: [keys(128) x D], A MN-major (P/dS), B MN-major (dO/Q)
      const uint32_t idesc_q = make_idesc(ROWS, D, 0, 1);        // dQ: [q x D], A K-major (dS), B MN-major (K)
      int it = 0;
      uint32_t hc = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        mbar_wait(in_full, (uint32_t)it & 1u);
        tc_fence_after();
        for (int h = 0; h < HPC; ++h, ++hc) {
          const uint32_t off = (uint32_t)((h * D) / 64) * CHUNK_BYTES + (uint32_t)((h * D) % 64) * 2u;
          mbar_wait(sdp_free, (hc & 1u) ^ 1u);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < D / 16; ++k) tc_mma(tmem, desc_k(q_t + off + k * 32u), desc_k(k_t + off + k * 32u), idesc_s, k > 0);
#pragma unroll
          for (int k = 0; k < D / 16; ++k) tc_mma(tmem + DP_COL, desc_k(do_t + off + k * 32u), desc_k(v_t + off + k * 32u), idesc_s, k > 0);
          tc_commit(sdp_full);
          mbar_wait(pds_full, hc & 1u);
          mbar_wait(out_free, (hc & 1u) ^ 1u);
          tc_fence_after();
          // dV = P^T dO, dK = dS^T Q : K dimension = query rows (8 steps of 16 rows = 2048 B in both operands)
#pragma unroll
          for (int k = 0; k < ROWS / 16; ++k) tc_mma(tmem + DV_COL, desc_mn(p_t + k * 2048u, CHUNK_BYTES), desc_mn(do_t + off + k * 2048u, CHUNK_BYTES), idesc_kv, k > 0);
#pragma unroll
          for (int k = 0; k < ROWS / 16; ++k) tc_mma(tmem + DK_COL, desc_mn(ds_t + k * 2048u, CHUNK_BYTES), desc_mn(q_t + off + k * 2048u, CHUNK_BYTES), idesc_kv, k > 0);
          // dQ = dS K : K dimension = keys
          for (int k = 0; k < nk_steps; ++k)
            tc_mma(tmem + DQ_COL, desc_k(ds_t + (k >> 2) * CHUNK_BYTES + (k & 3) * 32u), desc_mn(k_t + off + k * 2048u, CHUNK_BYTES), idesc_q, k > 0);
          tc_commit(out_full);
          tc_commit(pds_empty);
        }
        tc_commit(in_empty);
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const float sc2 = g.scale * LOG2E;
    uint32_t hc = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int b = w / g.groups, grp = w % g.groups, col0 = grp * 128;
      for (int h = 0; h < HPC; ++h, ++hc) {
        const bool valid_row = row < g.S;
        const float lse2 = valid_row ? g.lse[((int64_t)b * g.H + grp * HPC + h) * g.S + row] * LOG2E : 0.f;
        mbar_wait(sdp_full, hc & 1u);
        tc_fence_after();
        mbar_wait(pds_empty, (hc & 1u) ^ 1u);        // previous head's dV/dK/dQ MMAs have finished reading P / dS
        if (threadIdx.x == 64) tma_wait_read();      // ... and the previous head's output stores have read the aliased staging
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // pass 1: P = exp2(s*scale*log2e - lse*log2e) -> smem (bf16); delta = sum_j P_ij dP_ij
        float delta = 0.f;
        for (int c = 0; c < g.NK; c += 16) {
          uint32_t sv[16], dv[16];
          tmem_ld16(tmem + lane_addr + (uint32_t)c, sv);
          tmem_ld16(tmem + lane_addr + (uint32_t)(DP_COL + c), dv);
          tmem_ld_wait();
          float p[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            p[j] = (valid_row && c + j < g.S) ? exp2f(fmaf(__uint_as_float(sv[j]), sc2, -lse2)) : 0.f;
            delta = fmaf(p[j], __uint_as_float(dv[j]), delta);
          }
          const uint32_t tile = p_t + (uint32_t)(c >> 6) * CHUNK_BYTES;
          const int c16 = (c & 63) >> 3;
          sts128(swz(tile, row, c16), pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
          sts128(swz(tile, row, c16 + 1), pack_bf16(p[8], p[9]), pack_bf16(p[10], p[11]), pack_bf16(p[12], p[13]), pack_bf16(p[14], p[15]));
        }
        // pass 2: dS = P * (dP - delta) * scale   (P re-read from this thread's own smem row, dP from TMEM)
        for (int c = 0; c < g.NK; c += 16) {
          uint32_t dv[16];
          tmem_ld16(tmem + lane_addr + (uint32_t)(DP_COL + c), dv);
          tmem_ld_wait();
          const uint32_t tile_p = p_t + (uint32_t)(c >> 6) * CHUNK_BYTES, tile_d = ds_t + (uint32_t)(c >> 6) * CHUNK_BYTES;
          const int c16 = (c & 63) >> 3;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t a0, a1, a2, a3;
            lds128(swz(tile_p, row, c16 + hh), a0, a1, a2, a3);
            const float p[8] = {bf16_lo(a0), bf16_hi(a0), bf16_lo(a1), bf16_hi(a1), bf16_lo(a2), bf16_hi(a2), bf16_lo(a3), bf16_hi(a3)};
            float s[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] = p[j] * (__uint_as_float(dv[hh * 8 + j]) - delta) * g.scale;
            sts128(swz(tile_d, row, c16 + hh), pack_bf16(s[0], s[1]), pack_bf16(s[2], s[3]), pack_bf16(s[4], s[5]), pack_bf16(s[6], s[7]));
          }
        }
        // key columns NK..127 of P / dS feed the M dimension of dV / dK (rows that are never stored) but must be finite
        for (int c = g.NK; c < 128; c += 8) {
          sts128(swz(p_t + (uint32_t)(c >> 6) * CHUNK_BYTES, row, (c & 63) >> 3), 0u, 0u, 0u, 0u);
          sts128(swz(ds_t + (uint32_t)(c >> 6) * CHUNK_BYTES, row, (c & 63) >> 3), 0u, 0u, 0u, 0u);
        }
        tc_fence_before();
        mbar_arrive(sdp_free);
        fence_async_smem();
        mbar_arrive(pds_full);
        // ---- drain dV, dK, dQ of this head: TMEM -> bf16 -> plain row-major staging -> TMA stores (rows >= S clipped)
        mbar_wait(out_full, hc & 1u);              // dV/dK/dQ complete => P/dS (== staging) no longer read by the tensor core
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const uint32_t tcol = (t == 0) ? DQ_COL : (t == 1 ? DK_COL : DV_COL);
          const uint32_t dst = stg + t * STG_BYTES + (uint32_t)row * (D * 2);
#pragma unroll
          for (int c = 0; c < D; c += 16) {
            uint32_t v[16];
            tmem_ld16(tmem + lane_addr + tcol + (uint32_t)c, v);
            tmem_ld_wait();
            sts128(dst + c * 2, pack_bf16(__uint_as_float(v[0]), __uint_as_float(v[1])), pack_bf16(__uint_as_float(v[2]), __uint_as_float(v[3])),
                   pack_bf16(__uint_as_float(v[4]), __uint_as_float(v[5])), pack_bf16(__uint_as_float(v[6]), __uint_as_float(v[7])));
            sts128(dst + c * 2 + 16, pack_bf16(__uint_as_float(v[8]), __uint_as_float(v[9])), pack_bf16(__uint_as_float(v[10]), __uint_as_float(v[11])),
                   pack_bf16(__uint_as_float(v[12]), __uint_as_float(v[13])), pack_bf16(__uint_as_float(v[14]), __uint_as_float(v[15])));
          }
        }
        tc_fence_before();
        mbar_arrive(out_free);
        fence_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) {
          const int c = col0 + h * D;
          tma_store_3d(&map_dq, stg, c, 0, b);
          tma_store_3d(&map_dk, stg + STG_BYTES, c, 0, b);
          tma_store_3d(&map_dv, stg + 2 * STG_BYTES, c, 0, b);
          tma_commit();
        }
      }
    }
    if (threadIdx.x == 64) tma_wait_read();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ================================================================================================ head-parallel kernels (S <= 96)
// Second generation of the single-tile kernels, built for the latency-bound E=128 regime (S=65, d=32: 2048 tiny problems).
// Work item = (batch element b, ONE 64-column SW128 chunk = 64/D heads).  Every head of the item has its own warpgroup
// (thread = query row), its own TMEM region and its own barriers, so heads progress independently; the single MMA thread is
// an event loop that issues whichever head's next MMA has its operands ready.  Tiles are packed to NK = ceil16(S) rows
// (10 KB instead of 16 KB at S=65): the rows NK..127 an M=128 MMA reads beyond a tile are whatever follows in smem -- they
// only produce accumulator rows that are never stored.  ~110 KB smem and 256 TMEM columns per CTA -> TWO CTAs per SM, which
// overlap each other's TMA latency.  O (forward) and dK/dQ (backward) accumulate over the consumed S / dP columns.
// Warp roles: warp 4h+q (q < 3) = rows 32q.. of head h; S <= 96 leaves the fourth warp of every warpgroup without rows, so
// warp 3 (lane 0) is the control thread (TMA producer + MMA issuer in one event loop) -- it spins on a scheduler (SMSP 3)
// that no softmax warp uses.
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void tmem_ld16p(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
// 16-byte chunk `c16` of row `row` in a [rows x D] bf16 staging tile whose TMA map swizzles with the span of one row
// (D = 64: SWIZZLE_128B, D = 32: SWIZZLE_64B -- byte-address bits [4, 4+B) ^= bits [7, 7+B)); bank-conflict-free row writes
template <int D>
__device__ __forceinline__ uint32_t stg_addr(uint32_t tile, int row, int c16) {
  if (D == 64) return tile + (uint32_t)row * 128u + (((uint32_t)c16 ^ ((uint32_t)row & 7u)) << 4);
  return tile + (uint32_t)row * 64u + (((uint32_t)c16 ^ (((uint32_t)row >> 1) & 3u)) << 4);
}
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int D, int NKG>
__global__ void __launch_bounds__(128 * (64 / D), 2)
attn_fwd_hp_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o, const Geo g) {
  constexpr int HPC = 64 / D, NK = NKG * 16, TS = NK * 128, RS = 256 / HPC, STAGES = 2;
  constexpr int STG_BYTES = NK * D * 2;                              // one head's [NK x D] bf16 output tile, plain row-major
  constexpr int CTRL_WARP = 3, NSM = 96;                             // control warp; softmax threads per head
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t in_base = base;                                     // [STAGES][Q, K, V][TS]
  const uint32_t p_base = in_base + STAGES * 3 * TS;                 // [HPC][2 key chunks][TS]
  const uint32_t stg_base = p_base + HPC * 2 * TS;                   // [HPC][STG_BYTES]
  const uint32_t bar_base = stg_base + HPC * STG_BYTES;
  auto in_full = [&](int i) { return bar_base + 8u * i; };
  auto in_empty = [&](int i) { return bar_base + 8u * (2 + i); };
  auto s_full = [&](int h) { return bar_base + 8u * (4 + h); };
  auto p_full = [&](int h) { return bar_base + 8u * (6 + h); };
  auto o_full = [&](int h) { return bar_base + 8u * (8 + h); };
  auto o_free = [&](int h) { return bar_base + 8u * (10 + h); };
  const uint32_t tmem_slot = bar_base + 8u * 12;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - smem_u32(smem_dyn)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {      // tensor-map fetches are ~0.9 us each when cold and serialise behind the first TMA that needs them
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_o) : "memory");
  }
  if (warp == CTRL_WARP && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(in_full(i), 1); mbar_init(in_empty(i), 1); }
    for (int h = 0; h < HPC; ++h) { mbar_init(s_full(h), 1); mbar_init(p_full(h), NSM); mbar_init(o_full(h), 1); mbar_init(o_free(h), NSM); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == CTRL_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  pdl_trigger();
  pdl_wait();
  const int total = g.B * g.groups;
  const int n_items = (int)blockIdx.x < total ? (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == CTRL_WARP) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(ROWS, NK, 0, 0);        // S = Q K^T : both K-major
      const uint32_t idesc_o = make_idesc(ROWS, D, 0, 1);         // O = P V   : A K-major, B MN-major
      int s_it[HPC], pv_it[HPC];
#pragma unroll
      for (int h = 0; h < HPC; ++h) { s_it[h] = 0; pv_it[h] = 0; }
      int loaded = 0, released = 0, tma_it = 0;
      while (released < n_items) {
        if (tma_it < n_items && mbar_try_wait(in_empty(tma_it % STAGES), (((uint32_t)(tma_it / STAGES)) & 1u) ^ 1u)) {
          const int w = blockIdx.x + tma_it * gridDim.x, b = w / g.groups, col0 = (w % g.groups) * 64, st = tma_it % STAGES;
          const uint32_t t = in_base + st * 3 * TS;
          mbar_expect_tx(in_full(st), 3u * TS);
          tma_load_3d(t, &map_q, in_full(st), col0, 0, b);
          tma_load_3d(t + TS, &map_k, in_full(st), col0, 0, b);
          tma_load_3d(t + 2 * TS, &map_v, in_full(st), col0, 0, b);
          ++tma_it;
        }
#pragma unroll
        for (int h = 0; h < HPC; ++h) {
          if (pv_it[h] < s_it[h]) {                               // S issued, softmax pending -> O_h = P_h V_h
            const int it = pv_it[h];
            if (mbar_try_wait(p_full(h), (uint32_t)it & 1u)) {
              tc_fence_after();
              const uint32_t vt = in_base + (it % STAGES) * 3 * TS + 2 * TS + (uint32_t)(h * D) * 2u;
              const uint32_t pt = p_base + h * 2 * TS;
#pragma unroll
              for (int k = 0; k < NKG; ++k)      // K = keys, 16 per step: P chunk (k/4) + 32 B per step; V rows 16k.. (2048 B per step)
                tc_mma(tmem + (uint32_t)(RS * h), desc_k(pt + (k >> 2) * TS + (k & 3) * 32u), desc_mn(vt + k * 2048u, TS), idesc_o, k > 0);
              tc_commit(o_full(h));
              pv_it[h] = it + 1;
            }
          } else if (s_it[h] < n_items) {                         // next item of this head: S_h = Q_h K_h^T
            const int it = s_it[h];
            if (it == loaded && mbar_try_wait(in_full(it % STAGES), ((uint32_t)(it / STAGES)) & 1u)) ++loaded;
            if (it < loaded && mbar_try_wait(o_free(h), ((uint32_t)it & 1u) ^ 1u)) {
              tc_fence_after();
              const uint32_t t = in_base + (it % STAGES) * 3 * TS + (uint32_t)(h * D) * 2u;
#pragma unroll
              for (int k = 0; k < D / 16; ++k)
                tc_mma(tmem + (uint32_t)(RS * h), desc_k(t + k * 32u), desc_k(t + TS + k * 32u), idesc_s, k > 0);
              tc_commit(s_full(h));
              s_it[h] = it + 1;
            }
          }
        }
        int mn = pv_it[0];
#pragma unroll
        for (int h = 1; h < HPC; ++h) mn = min(mn, pv_it[h]);
        if (mn > released) { tc_commit(in_empty(released % STAGES)); ++released; }
      }
    }
  } else if ((warp & 3) != 3) {
    // ---------------------------------------------------------------- softmax + output drain: warpgroup = head, thread = query row
    const int h = warp >> 2, quad = warp & 3, row = quad * 32 + lane;
    const uint32_t t_s = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(RS * h);
    const bool quad_on = quad * 32 < NK;                         // warp-uniform: this warp owns rows of the tile
    const uint32_t pt = p_base + h * 2 * TS, stg = stg_base + h * STG_BYTES;
    const bool leader = quad == 0 && lane == 0;
    const float sc2 = g.scale * LOG2E;
    for (int it = 0; it < n_items; ++it) {
      const int w = blockIdx.x + it * gridDim.x, b = w / g.groups, grp = w % g.groups;
      const uint32_t par = (uint32_t)it & 1u;
      mbar_wait(s_full(h), par);
      tc_fence_after();
      float inv_l = 0.f;
      if (quad_on) {
        uint32_t v[NK];
#pragma unroll
        for (int c = 0; c < NKG; ++c) tmem_ld16p(t_s + 16 * c, &v[16 * c]);
        tmem_ld_wait();
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < NKG; ++c) {
          if (16 * (c + 1) <= g.S) {
#pragma unroll
            for (int j = 0; j < 16; ++j) m = fmaxf(m, __uint_as_float(v[16 * c + j]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (16 * c + j < g.S) m = fmaxf(m, __uint_as_float(v[16 * c + j]));
          }
        }
        const float mb = m * sc2;
        float l = 0.f;
#pragma unroll
        for (int c = 0; c < NKG; ++c) {
          uint32_t pk[8];
          if (16 * (c + 1) <= g.S) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float p0 = ex2_approx(fmaf(__uint_as_float(v[16 * c + 2 * j]), sc2, -mb));
              const float p1 = ex2_approx(fmaf(__uint_as_float(v[16 * c + 2 * j + 1]), sc2, -mb));
              l += p0 + p1;
              pk[j] = pack_bf16(p0, p1);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float p0 = (16 * c + 2 * j < g.S) ? ex2_approx(fmaf(__uint_as_float(v[16 * c + 2 * j]), sc2, -mb)) : 0.f;
              const float p1 = (16 * c + 2 * j + 1 < g.S) ? ex2_approx(fmaf(__uint_as_float(v[16 * c + 2 * j + 1]), sc2, -mb)) : 0.f;
              l += p0 + p1;
              pk[j] = pack_bf16(p0, p1);
            }
          }
          if (row < NK) {
            const uint32_t tile = pt + (uint32_t)(c >> 2) * TS;
            const int c16 = (c & 3) * 2;
            sts128(swz(tile, row, c16), pk[0], pk[1], pk[2], pk[3]);
            sts128(swz(tile, row, c16 + 1), pk[4], pk[5], pk[6], pk[7]);
          }
        }
        inv_l = 1.0f / l;
        if (row < g.S) g.lse[((int64_t)b * g.H + grp * HPC + h) * g.S + row] = m * g.scale + __logf(l);
      }
      tc_fence_before();
      fence_async_smem();                      // P visible to the tensor-core (async) proxy
      mbar_arrive(p_full(h));
      // ---- drain O_h: TMEM -> * 1/l -> bf16 -> staging -> TMA store (rows >= S clipped by the tensor map)
      mbar_wait(o_full(h), par);
      tc_fence_after();
      if (leader) tma_wait_read();             // the previous item's store has finished reading the staging tile
      named_bar(1 + h, NSM);
      if (quad_on) {
        uint32_t o[D];
#pragma unroll
        for (int c = 0; c < D; c += 16) tmem_ld16p(t_s + c, &o[c]);
        tmem_ld_wait();
        if (row < NK) {
#pragma unroll
          for (int c = 0; c < D; c += 8)
            sts128(stg_addr<D>(stg, row, c >> 3),
                   pack_bf16(__uint_as_float(o[c]) * inv_l, __uint_as_float(o[c + 1]) * inv_l),
                   pack_bf16(__uint_as_float(o[c + 2]) * inv_l, __uint_as_float(o[c + 3]) * inv_l),
                   pack_bf16(__uint_as_float(o[c + 4]) * inv_l, __uint_as_float(o[c + 5]) * inv_l),
                   pack_bf16(__uint_as_float(o[c + 6]) * inv_l, __uint_as_float(o[c + 7]) * inv_l));
        }
      }
      tc_fence_before();
      mbar_arrive(o_free(h));                  // O_h (== the S_h columns) may be overwritten by the next item's S_h
      fence_async_smem();
      named_bar(1 + h, NSM);
      if (leader) { tma_store_3d(&map_o, stg, grp * 64 + h * D, 0, b); tma_commit(); }
    }
    if (leader) tma_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == CTRL_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

// Backward, same organisation.  Per head: S -> P (registers + smem) ; dP = dO V^T over the consumed S columns, dV = P^T dO
// beside them ; dS overwrites P in smem (P stays in registers as packed bf16) ; dK = dS^T Q and dQ = dS K over the consumed
// dP columns ; drain dQ | dK | dV -> staging -> three TMA stores.  TMEM region of a head: [0, NK) S/dP then dK [0, D) and
// dQ [D, 2D); dV at [max(NK, 2D), +D).
template <int D, int NKG>
__global__ void __launch_bounds__(128 * (64 / D), 2)
attn_bwd_hp_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_do,
                   const __grid_constant__ CUtensorMap map_dq, const __grid_constant__ CUtensorMap map_dk,
                   const __grid_constant__ CUtensorMap map_dv, const Geo g) {
  constexpr int HPC = 64 / D, NK = NKG * 16, TS = NK * 128, RS = 256 / HPC;
  constexpr int STG_BYTES = NK * D * 2;
  constexpr int DK_COL = 0, DQ_COL = D, DV_COL = NK > 2 * D ? NK : 2 * D;
  constexpr int CTRL_WARP = 3, NSM = 96;                             // control warp; softmax threads per head
  static_assert(DV_COL + D <= RS, "TMEM region layout");
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t q_t = base, k_t = base + TS, v_t = base + 2 * TS, do_t = base + 3 * TS;   // single stage (the other CTA of the SM overlaps)
  const uint32_t p_base = base + 4 * TS;                               // [HPC][2 key chunks][TS]  P, then dS
  const uint32_t stg_base = p_base + HPC * 2 * TS;                     // [HPC][3][STG_BYTES]
  const uint32_t bar_base = stg_base + HPC * 3 * STG_BYTES;
  const uint32_t in_full = bar_base, in_empty = bar_base + 8;
  auto s_full = [&](int h) { return bar_base + 8u * (2 + h); };
  auto p_ready = [&](int h) { return bar_base + 8u * (4 + h); };
  auto dp_full = [&](int h) { return bar_base + 8u * (6 + h); };
  auto ds_ready = [&](int h) { return bar_base + 8u * (8 + h); };
  auto out_full = [&](int h) { return bar_base + 8u * (10 + h); };
  auto out_free = [&](int h) { return bar_base + 8u * (12 + h); };
  const uint32_t tmem_slot = bar_base + 8u * 14;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - smem_u32(smem_dyn)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_do) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dq) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dk) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dv) : "memory");
  }
  if (warp == CTRL_WARP && lane == 0) {
    mbar_init(in_full, 1); mbar_init(in_empty, 1);
    for (int h = 0; h < HPC; ++h) {
      mbar_init(s_full(h), 1); mbar_init(p_ready(h), NSM); mbar_init(dp_full(h), 1); mbar_init(ds_ready(h), NSM);
      mbar_init(out_full(h), 1); mbar_init(out_free(h), NSM);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == CTRL_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  pdl_trigger();
  pdl_wait();
  const int total = g.B * g.groups;
  const int n_items = (int)blockIdx.x < total ? (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == CTRL_WARP) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(ROWS, NK, 0, 0);       // S, dP: [q x keys], A,B K-major
      const uint32_t idesc_kv = make_idesc(ROWS, D, 1, 1);       // dV, dK: [keys(128) x D], A MN-major (P/dS), B MN-major (dO/Q)
      const uint32_t idesc_q = make_idesc(ROWS, D, 0, 1);        // dQ: [q x D], A K-major (dS), B MN-major (K)
      int item[HPC], phase[HPC];                                  // per head: current item, next MMA group (0: S, 1: dP+dV, 2: dK+dQ)
#pragma unroll
      for (int h = 0; h < HPC; ++h) { item[h] = 0; phase[h] = 0; }
      int loaded = 0, released = 0, tma_it = 0;
      while (released < n_items) {
        if (tma_it < n_items && mbar_try_wait(in_empty, ((uint32_t)tma_it & 1u) ^ 1u)) {
          const int w = blockIdx.x + tma_it * gridDim.x, b = w / g.groups, col0 = (w % g.groups) * 64;
          mbar_expect_tx(in_full, 4u * TS);
          tma_load_3d(q_t, &map_q, in_full, col0, 0, b);
          tma_load_3d(k_t, &map_k, in_full, col0, 0, b);
          tma_load_3d(do_t, &map_do, in_full, col0, 0, b);
          tma_load_3d(v_t, &map_v, in_full, col0, 0, b);
          ++tma_it;
        }
#pragma unroll
        for (int h = 0; h < HPC; ++h) {
          const int it = item[h];
          if (it >= n_items) continue;
          const uint32_t par = (uint32_t)it & 1u;
          const uint32_t off = (uint32_t)(h * D) * 2u;
          const uint32_t rt = tmem + (uint32_t)(RS * h);
          const uint32_t pt = p_base + h * 2 * TS;
          if (phase[h] == 0) {
            if (it == loaded && mbar_try_wait(in_full, par)) ++loaded;
            if (it < loaded && mbar_try_wait(out_free(h), par ^ 1u)) {
              tc_fence_after();
#pragma unroll
              for (int k = 0; k < D / 16; ++k) tc_mma(rt, desc_k(q_t + off + k * 32u), desc_k(k_t + off + k * 32u), idesc_s, k > 0);
              tc_commit(s_full(h));
              phase[h] = 1;
            }
          } else if (phase[h] == 1) {
            if (mbar_try_wait(p_ready(h), par)) {
              tc_fence_after();
#pragma unroll
              for (int k = 0; k < D / 16; ++k) tc_mma(rt, desc_k(do_t + off + k * 32u), desc_k(v_t + off + k * 32u), idesc_s, k > 0);
#pragma unroll
              for (int k = 0; k < NKG; ++k)      // dV = P^T dO : K dimension = query rows, 16 rows = 2048 B per step in both operands
                tc_mma(rt + DV_COL, desc_mn(pt + k * 2048u, TS), desc_mn(do_t + off + k * 2048u, TS), idesc_kv, k > 0);
              tc_commit(dp_full(h));
              phase[h] = 2;
            }
          } else {
            if (mbar_try_wait(ds_ready(h), par)) {
              tc_fence_after();
#pragma unroll
              for (int k = 0; k < NKG; ++k)      // dK = dS^T Q
                tc_mma(rt + DK_COL, desc_mn(pt + k * 2048u, TS), desc_mn(q_t + off + k * 2048u, TS), idesc_kv, k > 0);
#pragma unroll
              for (int k = 0; k < NKG; ++k)      // dQ = dS K : K dimension = keys
                tc_mma(rt + DQ_COL, desc_k(pt + (k >> 2) * TS + (k & 3) * 32u), desc_mn(k_t + off + k * 2048u, TS), idesc_q, k > 0);
              tc_commit(out_full(h));
              phase[h] = 0;
              item[h] = it + 1;
            }
          }
        }
        int mn = item[0];
#pragma unroll
        for (int h = 1; h < HPC; ++h) mn = min(mn, item[h]);
        if (mn > released) { tc_commit(in_empty); ++released; }
      }
    }
  } else if ((warp & 3) != 3) {
    const int h = warp >> 2, quad = warp & 3, row = quad * 32 + lane;
    const uint32_t t_r = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(RS * h);
    const bool quad_on = quad * 32 < NK;
    const uint32_t pt = p_base + h * 2 * TS, stg = stg_base + h * 3 * STG_BYTES;
    const bool leader = quad == 0 && lane == 0;
    const float sc2 = g.scale * LOG2E;
    const bool valid_row = row < g.S;
    for (int it = 0; it < n_items; ++it) {
      const int w = blockIdx.x + it * gridDim.x, b = w / g.groups, grp = w % g.groups;
      const uint32_t par = (uint32_t)it & 1u;
      const float lse2 = valid_row ? __ldg(g.lse + ((int64_t)b * g.H + grp * HPC + h) * g.S + row) * LOG2E : 0.f;
      uint32_t pk[NK / 2];                                         // this row of P as packed bf16
      mbar_wait(s_full(h), par);
      tc_fence_after();
      if (quad_on) {
        // pass 1: P = exp2(s*scale*log2e - lse*log2e) -> registers + smem (rows >= S and keys >= S are exact zeros)
#pragma unroll
        for (int c = 0; c < NKG; ++c) {
          uint32_t sv[16];
          tmem_ld16p(t_r + 16 * c, sv);
          tmem_ld_wait();
          if (16 * (c + 1) <= g.S) {               // warp-uniform: every key of the group is real
#pragma unroll
            for (int j = 0; j < 8; ++j)
              pk[8 * c + j] = pack_bf16(ex2_approx(fmaf(__uint_as_float(sv[2 * j]), sc2, -lse2)), ex2_approx(fmaf(__uint_as_float(sv[2 * j + 1]), sc2, -lse2)));
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float p0 = (16 * c + 2 * j < g.S) ? ex2_approx(fmaf(__uint_as_float(sv[2 * j]), sc2, -lse2)) : 0.f;
              const float p1 = (16 * c + 2 * j + 1 < g.S) ? ex2_approx(fmaf(__uint_as_float(sv[2 * j + 1]), sc2, -lse2)) : 0.f;
              pk[8 * c + j] = pack_bf16(p0, p1);
            }
          }
          if (!valid_row) {                        // rows S..NK-1 enter dV / dK as contraction rows: exact zeros
#pragma unroll
            for (int j = 0; j < 8; ++j) pk[8 * c + j] = 0u;
          }
          if (row < NK) {
            const uint32_t tile = pt + (uint32_t)(c >> 2) * TS;
            const int c16 = (c & 3) * 2;
            sts128(swz(tile, row, c16), pk[8 * c], pk[8 * c + 1], pk[8 * c + 2], pk[8 * c + 3]);
            sts128(swz(tile, row, c16 + 1), pk[8 * c + 4], pk[8 * c + 5], pk[8 * c + 6], pk[8 * c + 7]);
          }
        }
      }
      tc_fence_before();
      fence_async_smem();
      mbar_arrive(p_ready(h));
      mbar_wait(dp_full(h), par);                  // dP complete; dV complete => P in smem no longer read
      tc_fence_after();
      if (quad_on) {
        // pass 2: delta = sum_j P_ij dP_ij
        float delta = 0.f;
#pragma unroll
        for (int c = 0; c < NKG; ++c) {
          uint32_t dv[16];
          tmem_ld16p(t_r + 16 * c, dv);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            delta = fmaf(bf16_lo(pk[8 * c + j]), __uint_as_float(dv[2 * j]), delta);
            delta = fmaf(bf16_hi(pk[8 * c + j]), __uint_as_float(dv[2 * j + 1]), delta);
          }
        }
        // pass 3: dS = P * (dP - delta) * scale -> smem, over P
        const float nds = -delta * g.scale;
#pragma unroll
        for (int c = 0; c < NKG; ++c) {
          uint32_t dv[16];
          tmem_ld16p(t_r + 16 * c, dv);
          tmem_ld_wait();
          uint32_t ds[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {      // p * (dp - delta) * scale  as  p * fma(dp, scale, -delta * scale): 2 instead of 3 per element
            const float s0 = bf16_lo(pk[8 * c + j]) * fmaf(__uint_as_float(dv[2 * j]), g.scale, nds);
            const float s1 = bf16_hi(pk[8 * c + j]) * fmaf(__uint_as_float(dv[2 * j + 1]), g.scale, nds);
            ds[j] = pack_bf16(s0, s1);
          }
          if (row < NK) {
            const uint32_t tile = pt + (uint32_t)(c >> 2) * TS;
            const int c16 = (c & 3) * 2;
            sts128(swz(tile, row, c16), ds[0], ds[1], ds[2], ds[3]);
            sts128(swz(tile, row, c16 + 1), ds[4], ds[5], ds[6], ds[7]);
          }
        }
      }
      tc_fence_before();
      fence_async_smem();
      mbar_arrive(ds_ready(h));
      // ---- drain dQ, dK, dV of this head: TMEM -> bf16 -> plain row-major staging -> TMA stores (rows >= S clipped)
      mbar_wait(out_full(h), par);
      tc_fence_after();
      if (leader) tma_wait_read();
      named_bar(1 + h, NSM);
      if (quad_on) {
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const uint32_t tcol = (t == 0) ? DQ_COL : (t == 1 ? DK_COL : DV_COL);
          uint32_t o[D];
#pragma unroll
          for (int c = 0; c < D; c += 16) tmem_ld16p(t_r + tcol + c, &o[c]);
          tmem_ld_wait();
          if (row < NK) {
#pragma unroll
            for (int c = 0; c < D; c += 8)
              sts128(stg_addr<D>(stg + t * STG_BYTES, row, c >> 3), pack_bf16(__uint_as_float(o[c]), __uint_as_float(o[c + 1])), pack_bf16(__uint_as_float(o[c + 2]), __uint_as_float(o[c + 3])),
                     pack_bf16(__uint_as_float(o[c + 4]), __uint_as_float(o[c + 5])), pack_bf16(__uint_as_float(o[c + 6]), __uint_as_float(o[c + 7])));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(out_free(h));
      fence_async_smem();
      named_bar(1 + h, NSM);
      if (leader) {
        const int c = grp * 64 + h * D;
        tma_store_3d(&map_dq, stg, c, 0, b);
        tma_store_3d(&map_dk, stg + STG_BYTES, c, 0, b);
        tma_store_3d(&map_dv, stg + 2 * STG_BYTES, c, 0, b);
        tma_commit();
      }
    }
    if (leader) tma_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == CTRL_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

// ================================================================================================ next: S = 257, d = 192 (C4) -- design notes
// Not built yet; these shapes (and v1's d = 96 / 108 and the L2-distance scores) still run on the CUDA-core kernels of
// attention.cu.  What the budget analysis of this round says the tcgen05 version has to look like:
//   forward, work item = (b, h, 128-row q tile), one CTA per SM:
//     TMEM  S [128 x NK<=272] fp32 (272 columns: one MMA of N = 256 plus one of N = 16) | O [128 x 192] (192 columns) = 464 <= 512;
//     smem  Q [128 x 192] = 3 SW128 chunks (48 KB) | P [128 x 272] bf16 K-major = 5 chunks (80 KB) | ring of 64-key K / V blocks,
//           3 chunks x 8 KB = 24 KB each, 3 stages (72 KB) = 200 KB;
//     flow  stream K blocks -> S (12 k-steps per block) ; one softmax over the whole key range (no online rescaling: 257 keys fit
//           in one S tile), thread = row, two TMEM passes ; stream V blocks -> O += P_j V_j (B MN-major, N = 192 through LBO) ;
//           drain through the P region (free once the last PV has committed) -> TMA store.  ~4.5 us per item, 3072 items per
//           GPU at C4 -> ~95 us per layer against ~600 us of GEMMs.  The third q tile holds one row (S = 256 + CLS): skip the
//           softmax work of empty TMEM lane quadrants as the head-parallel kernels do.
//   backward: d = 192 does not leave room for dK / dV accumulators of all 272 keys (3 M tiles x 192 columns x 2 = 1152 TMEM
//     columns), so either (a) item = (b, h, q tile): S and dP share [0, 272), one [128 x 192] output tile at a time behind them,
//     dQ stored, partial dK / dV of the three q tiles combined with bf16 TMA reduce-add into zero-initialised buffers; smem is
//     the constraint (Q 48 + dO 48 + P/dS 80 + ring 48 = 224 KB, staging must alias the ring); or (b) the H = 12 / d = 64
//     variant of the same model (identical FLOPs), where K / V of a head (34 KB each) stay resident and only
//     S / dP need streaming.  The L2-distance variant adds the row norms |q|^2, |k|^2 (one extra N = 16 ones-MMA each, as in
//     gemm_tc's a_rowsum) and a sqrt in the score epilogue; d = 108 needs the head padded to 112 columns at projection time.

// ================================================================================================ host
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
// bf16 tensor viewed as [B, S, cols] with row pitch ld elements; box {box_cols, box_rows, 1}
int make_map3(CUtensorMap* map, const void* ptr, int B, int S, int cols, int64_t ld, int box_cols, int box_rows, bool swizzle, bool swz_row = false) {
  EncodeTiledFn enc = get_encode();
  VG_REQUIRE(enc != nullptr, VG_ERR_LAUNCH, "attention_tc: cuTensorMapEncodeTiled not available");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)S, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)S * ld * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz_row ? (box_cols * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)
                           : (swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VG_REQUIRE(r == CUDA_SUCCESS, VG_ERR_LAUNCH, "attention_tc: cuTensorMapEncodeTiled failed (%d) B=%d S=%d cols=%d ld=%lld box=%dx%d", (int)r, B, S,
             cols, (long long)ld, box_cols, box_rows);
  return VG_OK;
}

constexpr int FWD_SMEM = 6 * CHUNK_BYTES + 4 * CHUNK_BYTES + 1024 + 256;
template <int D> constexpr int bwd_smem() { return 12 * CHUNK_BYTES + 1024 + 256; }

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

constexpr int HP_MAX_S = 96;
// dynamic smem of the head-parallel kernels: tiles + barriers, and the M=128 over-read of the last K-major A tile
int hp_fwd_smem(int D, int NK) {
  const int HPC = 64 / D, TS = NK * 128;
  const int p_end = 2 * 3 * TS + HPC * 2 * TS, bar_end = p_end + HPC * NK * D * 2 + 128;
  return 1024 + max(bar_end, p_end + (128 - NK) * 128);
}
int hp_bwd_smem(int D, int NK) {
  const int HPC = 64 / D, TS = NK * 128;
  const int p_end = 4 * TS + HPC * 2 * TS, bar_end = p_end + HPC * 3 * NK * D * 2 + 128;
  return 1024 + max(bar_end, p_end + (128 - NK) * 128);
}
bool hp_disabled() {
  static const bool off = [] { const char* e = getenv("VG_ATTN_HP"); return e && !strcmp(e, "0"); }();   // read once, thread-safe
  return off;
}
int hp_ctas_per_sm() {
  static const int n = [] { const char* e = getenv("VG_ATTN_HP_CTAS"); const int v = e ? atoi(e) : 2; return (v < 1 || v > 4) ? 2 : v; }();
  return n;
}

template <int D, int NKG>
int launch_fwd_hp(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mo, const Geo& g, cudaStream_t st) {
  const int smem = hp_fwd_smem(D, NKG * 16);
  static const cudaError_t attr_e = cudaFuncSetAttribute(attn_fwd_hp_kernel<D, NKG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   // one-time, thread-safe
    VG_REQUIRE(attr_e == cudaSuccess, VG_ERR_LAUNCH, "attention_tc: smem attr");
  const int total = g.B * g.groups, grid = min(total, hp_ctas_per_sm() * num_sms());
  launch_pdl(attn_fwd_hp_kernel<D, NKG>, dim3(grid), dim3(128 * (64 / D)), (size_t)smem, st, mq, mk, mv, mo, g);
  return check_launch("attention_fwd_hp");
}
template <int D, int NKG>
int launch_bwd_hp(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mdo, const CUtensorMap& mdq,
                  const CUtensorMap& mdk, const CUtensorMap& mdv, const Geo& g, cudaStream_t st) {
  const int smem = hp_bwd_smem(D, NKG * 16);
  static const cudaError_t attr_e = cudaFuncSetAttribute(attn_bwd_hp_kernel<D, NKG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   // one-time, thread-safe
    VG_REQUIRE(attr_e == cudaSuccess, VG_ERR_LAUNCH, "attention_tc: smem attr");
  const int total = g.B * g.groups, grid = min(total, hp_ctas_per_sm() * num_sms());
  launch_pdl(attn_bwd_hp_kernel<D, NKG>, dim3(grid), dim3(128 * (64 / D)), (size_t)smem, st, mq, mk, mv, mdo, mdq, mdk, mdv, g);
  return check_launch("attention_bwd_hp");
}
#define VG_HP_DISPATCH(D_, NKG_, CALL)                                   \
  switch (NKG_) {                                                        \
    case 1: { constexpr int NKG = 1; if (D_ == 32) { constexpr int D = 32; CALL; } else { constexpr int D = 64; CALL; } } break; \
    case 2: { constexpr int NKG = 2; if (D_ == 32) { constexpr int D = 32; CALL; } else { constexpr int D = 64; CALL; } } break; \
    case 3: { constexpr int NKG = 3; if (D_ == 32) { constexpr int D = 32; CALL; } else { constexpr int D = 64; CALL; } } break; \
    case 4: { constexpr int NKG = 4; if (D_ == 32) { constexpr int D = 32; CALL; } else { constexpr int D = 64; CALL; } } break; \
    case 5: { constexpr int NKG = 5; if (D_ == 32) { constexpr int D = 32; CALL; } else { constexpr int D = 64; CALL; } } break; \
    default: { constexpr int NKG = 6; if (D_ == 32) { constexpr int D = 32; CALL; } else { constexpr int D = 64; CALL; } } break; \
  }

}  // namespace

bool attention_tc_supported(int dtype, int mode, int B, int H, int S, int d, const void* q, const void* k, const void* v,
                            int64_t ld_qkv, const void* o, int64_t ld_o) {
  static const int sm100 = vg_device_is_sm100();                                                               // one-time, thread-safe
  static const bool forced_off = [] { const char* e = getenv("VG_ATTN_PATH"); return e && !strcmp(e, "simt"); }();
  if (!sm100 || forced_off) return false;
  if (dtype != VG_BF16 || mode != VG_ATTN_DOT) return false;
  if (!(d == 32 || d == 64) || S < 1 || S > 128) return false;
  if ((H * d) % 128 != 0 && !((H * d) % 64 == 0 && S <= HP_MAX_S && !hp_disabled())) return false;   // head-parallel kernels work on 64-column items; the single-tile ones on 128
  if (ld_qkv % 8 || ld_o % 8) return false;
  return aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o);
}

int attention_fwd_tc(int B, int H, int S, int d, const void* q, const void* k, const void* v, int64_t ld, void* o, int64_t ldo,
                     float* lse, float scale, cudaStream_t st) {
  const int cols = H * d, NK = (S + 15) / 16 * 16;
  CUtensorMap mq, mk, mv, mo;
  int rc;
  if (S <= HP_MAX_S && !hp_disabled()) {
    if ((rc = make_map3(&mq, q, B, S, cols, ld, 64, NK, true))) return rc;
    if ((rc = make_map3(&mk, k, B, S, cols, ld, 64, NK, true))) return rc;
    if ((rc = make_map3(&mv, v, B, S, cols, ld, 64, NK, true))) return rc;
    if ((rc = make_map3(&mo, o, B, S, cols, ldo, d, NK, false, true))) return rc;
    Geo g; g.B = B; g.H = H; g.S = S; g.NK = NK; g.groups = cols / 64; g.scale = scale; g.lse = lse;
    VG_HP_DISPATCH(d, NK / 16, (rc = launch_fwd_hp<D, NKG>(mq, mk, mv, mo, g, st)));
    return rc;
  }
  if ((rc = make_map3(&mq, q, B, S, cols, ld, 64, ROWS, true))) return rc;
  if ((rc = make_map3(&mk, k, B, S, cols, ld, 64, NK, true))) return rc;
  if ((rc = make_map3(&mv, v, B, S, cols, ld, 64, NK, true))) return rc;
  if ((rc = make_map3(&mo, o, B, S, cols, ldo, 64, ROWS, true))) return rc;
  Geo g; g.B = B; g.H = H; g.S = S; g.NK = NK; g.groups = cols / 128; g.scale = scale; g.lse = lse;
  const int total = B * g.groups, grid = min(total, num_sms());
  if (d == 32) {
    static const cudaError_t attr_e = cudaFuncSetAttribute(attn_fwd_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);   // one-time, thread-safe
    VG_REQUIRE(attr_e == cudaSuccess, VG_ERR_LAUNCH, "attention_tc: smem attr");
    launch_pdl(attn_fwd_tc_kernel<32>, dim3(grid), dim3(NTHREADS), FWD_SMEM, st, mq, mk, mv, mo, g);
  } else {
    static const cudaError_t attr_e = cudaFuncSetAttribute(attn_fwd_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);   // one-time, thread-safe
    VG_REQUIRE(attr_e == cudaSuccess, VG_ERR_LAUNCH, "attention_tc: smem attr");
    launch_pdl(attn_fwd_tc_kernel<64>, dim3(grid), dim3(NTHREADS), FWD_SMEM, st, mq, mk, mv, mo, g);
  }
  return check_launch("attention_fwd_tc");
}

int attention_bwd_tc(int B, int H, int S, int d, const void* q, const void* k, const void* v, int64_t ld, const void* d_o,
                     int64_t ldo, const float* lse, void* dq, void* dk, void* dv, int64_t ldd, float scale, cudaStream_t st) {
  const int cols = H * d, NK = (S + 15) / 16 * 16;
  CUtensorMap mq, mk, mv, mdo, mdq, mdk, mdv;
  int rc;
  if (S <= HP_MAX_S && !hp_disabled()) {
    if ((rc = make_map3(&mq, q, B, S, cols, ld, 64, NK, true))) return rc;
    if ((rc = make_map3(&mk, k, B, S, cols, ld, 64, NK, true))) return rc;
    if ((rc = make_map3(&mv, v, B, S, cols, ld, 64, NK, true))) return rc;
    if ((rc = make_map3(&mdo, d_o, B, S, cols, ldo, 64, NK, true))) return rc;
    if ((rc = make_map3(&mdq, dq, B, S, cols, ldd, d, NK, false, true))) return rc;
    if ((rc = make_map3(&mdk, dk, B, S, cols, ldd, d, NK, false, true))) return rc;
    if ((rc = make_map3(&mdv, dv, B, S, cols, ldd, d, NK, false, true))) return rc;
    Geo g; g.B = B; g.H = H; g.S = S; g.NK = NK; g.groups = cols / 64; g.scale = scale; g.lse = const_cast<float*>(lse);
    VG_HP_DISPATCH(d, NK / 16, (rc = launch_bwd_hp<D, NKG>(mq, mk, mv, mdo, mdq, mdk, mdv, g, st)));
    return rc;
  }
  if ((rc = make_map3(&mq, q, B, S, cols, ld, 64, ROWS, true))) return rc;
  if ((rc = make_map3(&mk, k, B, S, cols, ld, 64, NK, true))) return rc;
  if ((rc = make_map3(&mv, v, B, S, cols, ld, 64, NK, true))) return rc;
  if ((rc = make_map3(&mdo, d_o, B, S, cols, ldo, 64, ROWS, true))) return rc;
  if ((rc = make_map3(&mdq, dq, B, S, cols, ldd, d, ROWS, false))) return rc;
  if ((rc = make_map3(&mdk, dk, B, S, cols, ldd, d, ROWS, false))) return rc;
  if ((rc = make_map3(&mdv, dv, B, S, cols, ldd, d, ROWS, false))) return rc;
  Geo g; g.B = B; g.H = H; g.S = S; g.NK = NK; g.groups = cols / 128; g.scale = scale; g.lse = const_cast<float*>(lse);
  const int total = B * g.groups, grid = min(total, num_sms());
  if (d == 32) {
    static const cudaError_t attr_e = cudaFuncSetAttribute(attn_bwd_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem<32>());   // one-time, thread-safe
    VG_REQUIRE(attr_e == cudaSuccess, VG_ERR_LAUNCH, "attention_tc: smem attr");
    launch_pdl(attn_bwd_tc_kernel<32>, dim3(grid), dim3(NTHREADS), bwd_smem<32>(), st, mq, mk, mv, mdo, mdq, mdk, mdv, g);
  } else {
    static const cudaError_t attr_e = cudaFuncSetAttribute(attn_bwd_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem<64>());   // one-time, thread-safe
    VG_REQUIRE(attr_e == cudaSuccess, VG_ERR_LAUNCH, "attention_tc: smem attr");
    launch_pdl(attn_bwd_tc_kernel<64>, dim3(grid), dim3(NTHREADS), bwd_smem<64>(), st, mq, mk, mv, mdo, mdq, mdk, mdv, g);
  }
  return check_launch("attention_bwd_tc");
}

}  // namespace vg
