/* vitgan_b200.h -- C ABI of libvitgan_b200.so: sm_100a kernels for the ViT-GAN train-step hot path.
 *
 * The reference (krzkro4122/vit-gan) is pure PyTorch and has no FFI layer; its "operator interface"
 * for this path is torch's own ATen calls made from nn.Module.forward.  Each entry point below
 * names the ATen call sequence / reference lines it replaces (paths relative to /root/reference).
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller;
 *    the library never allocates, frees or retains caller memory;
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), performs no host
 *    synchronisation and is CUDA-graph capturable;
 *  - returns 0 on success, a negative vg_status otherwise; vg_last_error() (thread-local) has text;
 *  - activations are row-major `dtype` (VG_F32 parity path / VG_BF16 fast path); parameters and
 *    parameter gradients are always fp32 unless stated; statistics (mean, rstd, lse) are fp32;
 *  - there is no CPU fallback: unsupported shapes are an error.
 */
#ifndef VITGAN_B200_H
#define VITGAN_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VG_ABI_VERSION 5

typedef enum { VG_F32 = 0, VG_BF16 = 1 } vg_dtype;
typedef enum {
  VG_OK = 0, VG_ERR_SHAPE = -1, VG_ERR_ALIGN = -2, VG_ERR_UNSUPPORTED = -3, VG_ERR_LAUNCH = -4, VG_ERR_ARG = -5
} vg_status;

/* GEMM epilogue activation (applied after bias, before residual) */
typedef enum {
  VG_ACT_NONE = 0,
  VG_ACT_GELU = 1,        /* exact erf GELU, nn.GELU() default (src/v2/modules.py:174)            */
  VG_ACT_TANH = 2,        /* Classifier (src/v2/modules.py:195-198)                               */
  VG_ACT_SIN = 3,         /* sin(act_param * x), SIREN (src/v1/siren.py:45)                       */
  VG_ACT_SIGMOID = 4,     /* D head (src/v1/discriminatorViT.py:51)                               */
  VG_ACT_MUL_DGELU = 5,   /* x * gelu'(aux)            (backward of GELU, aux = pre-activation)   */
  VG_ACT_MUL_DTANH = 6,   /* x * (1 - aux^2)           (aux = tanh output)                        */
  VG_ACT_MUL_DSIN = 7,    /* x * act_param*cos(act_param*aux)  (aux = pre-activation)             */
  VG_ACT_MUL_DSIGMOID = 8 /* x * aux*(1-aux)           (aux = sigmoid output)                     */
} vg_act;

typedef enum { VG_GEMM_AUTO = -1, VG_GEMM_SIMT = 0, VG_GEMM_TCGEN05 = 1 } vg_gemm_path;
typedef enum { VG_ATTN_DOT = 0, VG_ATTN_L2 = 1 } vg_attn_mode;

/* Parameter block of vg_gemm.  C[M,N] = opA(A)[M,K] * opB(B)[K,N], row-major storage:
 *   trans_a == 0: A stored [M,K] (lda >= K)      trans_a == 1: A stored [K,M] (lda >= M)
 *   trans_b == 0: B stored [K,N] (ldb >= N)      trans_b == 1: B stored [N,K] (ldb >= K)   (= nn.Linear weight)
 * epilogue, per element (m,n):  v = acc (+ bias[n]);  if (c_pre) c_pre[m,n] = v;  v = act(v | aux[m,n]);
 *                               v += residual[res_row(m), n];  C[out_row(m), n] (+)= v
 *   out_row(m) = m + (m / c_row_group) + 1            if c_row_group > 0 (leaves one gap row per group: CLS slot)
 *   res_row(m) = (m % res_row_mod) + res_row_off      if res_row_mod > 0 (broadcast positional embedding)
 *   accumulate != 0: C must be fp32; partial sums are atomically ADDED into C (split-K weight gradients).
 * Replaces: F.linear / addmm / mm / bmm calls of src/v2/modules.py:128-139,161,173-175,195-198,370 and
 *           src/v1/attention.py:46-48,102, muilti_layer_perceptron.py:39, siren.py:45, patch_encoder.py:44,
 *           and their autograd backward (mm for dgrad / wgrad).
 */
typedef struct {
  int path;                 /* vg_gemm_path */
  int ab_dtype;             /* dtype of A and B */
  int c_dtype;              /* dtype of C, c_pre, residual, aux */
  int trans_a, trans_b;
  int M, N, K;
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  void* C; int64_t ldc;
  const float* bias;        /* [N] fp32 or NULL */
  int act; float act_param;
  const void* aux; int64_t ldaux;
  const void* residual; int64_t ldres;
  void* c_pre; int64_t ldpre;
  int c_row_group;
  int res_row_mod, res_row_off;
  int accumulate;
  float* a_rowsum;          /* optional [M] fp32, accumulated: a_rowsum[m] += sum_k opA(A)[m,k].  With the wgrad form
                             * (trans_a=1, A = dY) this is the bias gradient colsum(dY), obtained on the tensor cores from an
                             * extra N=16 MMA against a tile of ones.  tcgen05 path in accumulate mode only
                             * (VG_ERR_UNSUPPORTED otherwise, nothing launched). */
  /* optional fused LayerNorm of the OUTPUT rows (src/v2/modules.py:168,172: the norm that follows out-proj / fc2 + skip):
   * ln_out = LN(C_row) * ln_gamma + ln_beta (same dtype as C), ln_mean / ln_rstd [M] fp32 saved for the backward.  The
   * statistics are taken over the bf16-rounded C values, i.e. exactly what a separate vg_layernorm_fwd would read.
   * tcgen05 path, N == 128 (one tile = one complete row), bf16 C, no aux / c_pre (VG_ERR_UNSUPPORTED otherwise). */
  const float* ln_gamma; const float* ln_beta;
  void* ln_out; int64_t ld_ln;
  float* ln_mean; float* ln_rstd;
  float ln_eps;
} vg_gemm_args;

int vg_version(void);
const char* vg_last_error(void);
/* 1 if the running device is compute capability 10.x (tcgen05 available) */
int vg_device_is_sm100(void);

int vg_gemm(const vg_gemm_args* args, void* stream);
/* Development aid: while `buffer` (device memory, >= 16 * gridDim u64, caller-zeroed) is set, every tcgen05 GEMM CTA records
 * %globaltimer at its pipeline milestones (entry, prologue done, dependency wait done, first tile's TMA issued / landed / MMA
 * committed / stored, stores drained, exit) into buffer[cta * 16 + k].  NULL switches it off.  Process-global, not thread-safe. */
int vg_gemm_set_trace(void* buffer);

/* dst[i] = (dst_dtype) (scale ? *scale : 1) * src[i]; used to make bf16 copies of fp32 parameters and to
 * apply the spectral rescale W <- sigma0/sigma * W (src/v1/attention.py:60-64) while packing. */
int vg_cast_scale(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n,
                  const float* num, const float* den, void* stream);

/* n fp32 [rows, cols] tensors, given as a DEVICE array of n pointers, -> dst [n * rows_pad, cols_pad] in dst_dtype: block t is
 * srcs[t] * num[t] / den[t] (either may be NULL) zero-padded to rows_pad x cols_pad.  One launch packs the 3H per-head q/k/v
 * weights of a v1 MultiHeadSelfAttention (src/v1/attention.py:46-48, spectral rescale :60-64) into the grouped projection
 * operand with head widths padded to the tensor-core granularity (108 -> 112); with n = 1 it pads columns (out-proj weight). */
int vg_pack_pad(const void* const* srcs, int n, int rows, int cols, int rows_pad, int cols_pad, const float* num,
                const float* den, void* dst, int dst_dtype, void* stream);

/* out[n] += sum_m x[m,n]   (bias gradients; fp32 out, accumulated: zero it first).
 * workspace (optional): persistent ZERO-INITIALISED [ws_rows = R, N] fp32 replicated accumulators + a zero-initialised
 * ticket `counter`: CTA partial sums go to row blockIdx %% R (same-address atomic contention / R), the last CTA folds the R
 * rows into `out` and re-zeroes them and the counter.  Without it the kernel falls back to direct global atomics. */
int vg_colsum(const void* x, int dtype, int64_t M, int N, int64_t ldx, float* out, float* workspace, int ws_rows,
              unsigned* counter, void* stream);

/* LayerNorm over the last dim (eps 1e-5, biased variance, affine): src/v2/modules.py:168,172,225;
 * src/v1/transformer.py:18-19.  Saves mean / rstd per row for the backward. */
int vg_layernorm_fwd(int dtype, int64_t rows, int E, const void* x, const float* gamma, const float* beta,
                     void* y, float* mean, float* rstd, float eps, void* stream);
/* dx = (dres ? dres : 0) + LN'(dy);  dgamma/dbeta are atomically accumulated (zero them first).
 * dres_colsum / dx_colsum (optional, fp32 [E], accumulated): column sums of the skip-path gradient `dres` and of the
 * produced `dx` -- the bias gradients of the Linear layers on either side of the norm, computed here for free.
 * dgamma == dbeta == NULL: dx only (the dgrad-only discriminator pass of the generator update), no reductions at all. */
int vg_layernorm_bwd(int dtype, int64_t rows, int E, const void* dy, const void* x, const float* mean,
                     const float* rstd, const float* gamma, const void* dres, void* dx,
                     float* dgamma, float* dbeta, float* dres_colsum, float* dx_colsum,
                     float* workspace /* persistent zeroed [ws_rows = R, 4*E] fp32 or NULL */, int ws_rows, unsigned* counter, void* stream);

/* vg_layernorm_bwd with DEFERRED column reductions (E <= 128): dx as above; every CTA stores its partial
 * [dgamma | dbeta | colsum(dres) | colsum(dx)] (4*E floats) to partials[cta] with plain stores -- no atomics and no
 * last-CTA fold on the critical path.  Returns the number of partial rows written (1..max_parts) or a negative status.
 * Fold them later (typically on another stream) with vg_fold_partials. */
int vg_layernorm_bwd_partials(int dtype, int64_t rows, int E, const void* dy, const void* x, const float* mean,
                              const float* rstd, const float* gamma, const void* dres, void* dx, float* partials,
                              int max_parts, void* stream);
/* outK[i] += sum_p partials[p][K*E + i], K = 0..3 (NULL outputs are skipped; fp32, accumulated atomically) */
int vg_fold_partials(const float* partials, int n_parts, int E, float* out0, float* out1, float* out2, float* out3,
                     void* stream);

/* Self-modulated LayerNorm  y = w * (gamma_s * (LN(h)*g + b) + beta_s)  (src/v1/spectral_layer_norm.py:19-20).
 * h has h_rows rows (h_rows == rows, or rows % h_rows == 0 for the first G layer where h is (S,F) and
 * broadcasts over the batch, src/v1/transformer.py:86).  gamma_s / beta_s are device scalars. */
int vg_sln_fwd(int dtype, int64_t rows, int64_t h_rows, int F, const void* h, const void* w,
               const float* ln_g, const float* ln_b, const float* gamma_s, const float* beta_s,
               void* y, float* mean, float* rstd, float eps, void* stream);
/* dh (+= dh_res if given), dw (+= dw_res if given); dgamma_s, dbeta_s, dln_g, dln_b atomically accumulated.
 * When h_rows < rows, dh is fp32 [h_rows,F] and atomically accumulated (zero it first). */
int vg_sln_bwd(int dtype, int64_t rows, int64_t h_rows, int F, const void* dy, const void* h, const void* w,
               const float* mean, const float* rstd, const float* ln_g, const float* ln_b,
               const float* gamma_s, const float* beta_s, const void* dh_res, const void* dw_res,
               void* dh, void* dw, float* dgamma_s, float* dbeta_s, float* dln_g, float* dln_b, void* stream);

/* Multi-head self-attention core, flash style (no SxS tensor in HBM).
 *   mode DOT: scores = q.k        (src/v2/modules.py:142-155; src/v1/attention.py:69-70)
 *   mode L2 : scores = +||q-k||_2 (src/v1/attention.py:66-67, torch.cdist matmul path semantics)
 *   out = softmax(scores * scale) @ v;  scale = 1/sqrt(d) (v2) or 1/sqrt(H*d) (v1).
 * q/k/v/o element (b,s,h,j) lives at base + (b*S+s)*ld + h*d + j  -> heads are read in place from the
 * fused projection output, and merged heads are written in place (no permute/cat copies).
 * lse: [B,H,S] fp32 log-sum-exp of the scaled scores, saved for the backward. */
int vg_attention_fwd(int dtype, int mode, int B, int H, int S, int d, const void* q, const void* k,
                     const void* v, int64_t ld_qkv, void* o, int64_t ld_o, float* lse, float scale, void* stream);
int vg_attention_bwd(int dtype, int mode, int B, int H, int S, int d, const void* q, const void* k,
                     const void* v, int64_t ld_qkv, const void* o, const void* d_o, int64_t ld_o,
                     const float* lse, void* dq, void* dk, void* dv, int64_t ld_dqkv, float scale,
                     float* delta_ws /* [B,H,S] fp32 scratch */, void* stream);
/* Which kernel family the two calls above take for a bf16/fp32 problem of this shape with 16-byte aligned operands:
 * 0 = CUDA-core flash kernel (fp32 parity path, odd head sizes), 1 = single-tile tcgen05 (S <= 128, d in {32, 64}, dot),
 * 2 = multi-tile tcgen05 (d in {96, 112, 192}, S <= 272; L2-distance scores for d = 96 / 112).  Test / bench introspection. */
int vg_attention_path(int dtype, int mode, int B, int H, int S, int d);
/* Development aid (like vg_gemm_set_trace): while `buffer` (device memory, 9 * 1024 u64, caller-zeroed) is set, CTA 0 of the
 * multi-tile attention kernels (roles 0-2 forward, 3-5 backward dQ, 6-8 backward dK/dV) logs (event code, %globaltimer) pairs per role -- producer, MMA issuer, compute leader -- at
 * buffer[role * 1024 + 2 i].  NULL switches it off.  Process-global, not thread-safe (profiles/trace_attn.py). */
int vg_attention_set_trace(void* buffer);

/* v2 patchify: img fp32 (B,C,I,I) -> patches[B*N, C*P*P] in `dtype`, k = c*P*P + i*P + j (conv weight order).
 * Together with vg_gemm (bias, c_row_group = N, residual = pos_embedding with res_row_mod = N) and
 * vg_fill_rows it replaces conv2d + reshape/permute + pos add + cat(cls) of src/v2/modules.py:82-100. */
int vg_im2col_patches(int dtype, int B, int C, int I, int P, const float* img, void* patches, void* stream);
/* inverse scatter (no overlap): dpatches -> dimg fp32 (B,C,I,I) */
int vg_col2im_patches(int dtype, int B, int C, int I, int P, const void* dpatches, float* dimg, void* stream);

/* v1 token gather in the reference's scrambled layout (src/v1/patch_encoder.py:54-73):
 * tokens[b, t, f] = flat element 432*t+f of the (C, n_h, n_w, win, win) unfold tensor of image b. */
int vg_v1_tokens_fwd(int dtype, int B, int C, int I, int win, int stride, int n_side, const float* img,
                     void* tokens, void* stream);
/* adjoint: dimg (fp32, zeroed by the caller) += scatter(dtokens) with atomics (windows overlap) */
int vg_v1_tokens_bwd(int dtype, int B, int C, int I, int win, int stride, int n_side, const void* dtokens,
                     float* dimg, void* stream);

/* x[b, row, :] = v[:] (+ v2[:]) for every b: writes the CLS row of a (B,S,E) token tensor
 * (src/v2/modules.py:96-98; src/v1/patch_encoder.py:45-50 where CLS also gets pos[0]). */
int vg_fill_rows(int dtype, int B, int S, int E, int row, const float* v, const float* v2, void* x, void* stream);
/* Embedding backward split: dx (B,S,E) -> dtok[B*(S-1), E] contiguous (dtype), dcls[E] += sum_b dx[b,0,:],
 * dpos[(S-1 or S), E] += sum_b dx[b, s, :]  (pos_has_cls: v1 positional embedding covers the CLS row). */
int vg_embed_bwd_split(int dtype, int B, int S, int E, const void* dx, void* dtok, float* dcls, float* dpos,
                       int pos_has_cls, void* stream);

/* sigma_max of n_mats row-major fp32 matrices [rows, cols] by power iteration with a persistent left
 * vector u (state [n_mats, rows], updated in place; all-zero state = cold start from a fixed vector).
 * Replaces the 3 full torch.svd calls per head per forward of src/v1/attention.py:54-58. */
int vg_sigma_max(const float* const* mats, int n_mats, int rows, int cols, float* u_state, int n_iters,
                 float* sigma_out, void* stream);

/* generic elementwise helpers used by heads and by the data-parallel runner */
int vg_add_inplace(int dtype, void* x, const void* y, int64_t n, void* stream);           /* x += y */
int vg_broadcast_rows(int dtype, const void* src, int64_t src_rows, int64_t cols, void* dst, int64_t reps, void* stream);
/* out[i] = dy[i] * act'(aux[i]) for act in {GELU, TANH, SIN, SIGMOID} (aux as in VG_ACT_MUL_D*): backward of the
 * activation of standalone Linear layers (heads, SIREN) whose dgrad GEMM belongs to the previous layer. */
int vg_act_backward(int dtype, int64_t n, const void* dy, const void* aux, int act, float act_param, void* out, void* stream);

/* Fused nn.CrossEntropyLoss (mean reduction, class-index int64 targets; src/v2/training.py:159) over rows/rows_per_group
 * groups of consecutive rows: losses[g] = mean_{r in group g} CE(logits[r], targets[r]);  dlogits = d(sum_g losses[g])/dlogits.
 * One launch instead of log_softmax + nll_loss forward/backward + reductions (loss head inside the
 * captured step).  fp32 logits [rows, C]; rows %% rows_per_group == 0; at most 64 groups. */
int vg_softmax_ce(const float* logits, const int64_t* targets, int rows, int C, int rows_per_group, float* losses,
                  float* dlogits, void* stream);

/* Fused nn.BCELoss (mean reduction, float targets in [0,1]; src/v1/gan.py:16-20 pick_criterion, used at gan.py:222-252) over
 * rows/rows_per_group groups of consecutive probabilities (D's sigmoid output [rows, 1]): losses[g] = mean BCE of group g with
 * torch's log clamp at -100; dprob = d(sum_g losses[g])/dprob with torch's 1e-12 denominator clamp.  At most 64 groups. */
int vg_bce(const float* prob, const float* target, int rows, int rows_per_group, float* losses, float* dprob, void* stream);

/* utils.convert_to_uint8 (src/v2/utils.py:194-196): out = uint8(clamp(x * 127.5 + 127.5, 0, 255)), the de-normalisation at
 * the end of the sampling path (src/v2/generation.py:47-56, utils.py:165-166); byte-exact vs torch on identical fp32 input. */
int vg_denorm_u8(int dtype, const void* x, int64_t n, uint8_t* out, void* stream);
/* dst[r*ld_dst + c] = src[r*ld_src + c]: strided row gather/scatter (CLS-token rows: x[:,0,:], src/v2/modules.py:195) */
int vg_copy_rows(int dtype, int64_t rows, int cols, const void* src, int64_t ld_src, void* dst, int64_t ld_dst, void* stream);

/* Fused multi-tensor Adam/AdamW over flat fp32 buffers (torch.optim.AdamW of
 * src/v2/training.py:150-157 and Adam of src/v1/gan.py:316-328).  step_count is a device int incremented
 * by the kernel so the update is graph-capturable.  decoupled != 0 -> AdamW weight decay. */
int vg_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                 float beta1, float beta2, float eps, float weight_decay, int decoupled, float grad_scale,
                 int* step_count, void* bf16_shadow /* optional [n] bf16 copy of the updated parameters */, void* stream);

/* self test of the tcgen05 GEMM against an in-kernel SIMT reference; returns 0 if max rel err < tol */
int vg_selftest_tcgen05(int M, int N, int K, int trans_a, int trans_b, float tol, float* max_err_out);

#ifdef __cplusplus
}
#endif
#endif /* VITGAN_B200_H */
