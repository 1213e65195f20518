"""Tensor-level wrappers over the C ABI: allocate outputs with torch, pass raw pointers + the current stream.

PyTorch is plumbing here (device memory, streams); every arithmetic kernel launched below is one of this
repository's own sm_100a kernels.  All functions require CUDA tensors and raise on anything else.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import lib as L
from .lib import lib, check

_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16}
_TORCH = {L.F32: torch.float32, L.BF16: torch.bfloat16}

# launches of this library's kernels since the last reset (bench.py reports it as gpu_launches)
launch_count = 0


def _count(n=1):
    global launch_count
    launch_count += n


def stream():
    return torch.cuda.current_stream().cuda_stream


# scratch for the low-contention column reductions (see cta_replica_reduce in csrc/common.cuh): per device one
# persistent ZEROED buffer of R replicated accumulator rows and a zeroed ticket counter; kernels leave both zeroed.
_REPLICAS = 16
_MAX_COLS = 4 * 1024
_scratch: dict = {}


def _reduce_scratch(device, ncols):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    sc = _scratch.get(idx)
    if sc is None:
        sc = (torch.zeros(_REPLICAS * _MAX_COLS, dtype=torch.float32, device=device), torch.zeros(16, dtype=torch.int32, device=device))
        _scratch[idx] = sc
    if ncols > _MAX_COLS:
        return None, 0, sc[1]
    return sc[0], _REPLICAS, sc[1]


def dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"vitgan_b200: unsupported dtype {t.dtype}") from None


def _req(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"vitgan_b200: `{name}` must be a CUDA tensor (there is no CPU path)")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _rowmajor2d(t: torch.Tensor, name: str):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"vitgan_b200: `{name}` must be 2-D with unit inner stride, got {tuple(t.shape)}/{t.stride()}")
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def gemm(a, b, *, trans_a=False, trans_b=True, bias=None, act=L.ACT_NONE, act_param=0.0, aux=None, residual=None,
         want_pre=False, out=None, out_dtype=None, accumulate=False, c_row_group=0, res_row_mod=0, res_row_off=0,
         path=L.GEMM_AUTO, rowsum_out=None, ln=None):
    """C[M,N] = opA(a) @ opB(b) with the fused epilogue of vg_gemm (see include/vitgan_b200.h).

    a: [M,K] (or [K,M] if trans_a); b: [N,K] if trans_b (an nn.Linear weight) else [K,N].
    Returns C, or (C, pre_activation) when want_pre.
    rowsum_out: optional fp32 [M] vector, accumulated with the row sums of opA(a) (the bias gradient when this is a
    weight-gradient GEMM); tcgen05 accumulate mode only -- raises VitganError (nothing launched) when unsupported.
    ln: optional (gamma, beta, eps): also return LayerNorm(C) and its statistics -> (C, ln_out, mean, rstd); tcgen05 path with
    N == 128 only -- raises VitganError (nothing launched) when unsupported.
    """
    _req(a, "a"); _req(b, "b")
    lda, ldb = _rowmajor2d(a, "a"), _rowmajor2d(b, "b")
    M, K = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[0], b.shape[1]) if trans_b else (b.shape[1], b.shape[0])
    if K != Kb:
        raise ValueError(f"vitgan_b200.gemm: inner dims differ ({K} vs {Kb})")
    if a.dtype != b.dtype:
        raise TypeError(f"vitgan_b200.gemm: A/B dtypes differ ({a.dtype} vs {b.dtype})")
    if out is None:
        rows = M + (M // c_row_group if c_row_group > 0 else 0)
        odt = out_dtype or (torch.float32 if accumulate else a.dtype)
        out = (torch.zeros if accumulate else torch.empty)((rows, N), dtype=odt, device=a.device)
    ldc = _rowmajor2d(out, "out")
    pre = torch.empty((M, N), dtype=out.dtype, device=a.device) if want_pre else None
    g = L.GemmArgs()
    g.path, g.ab_dtype, g.c_dtype = path, dt(a), dt(out)
    g.trans_a, g.trans_b, g.M, g.N, g.K = int(trans_a), int(trans_b), M, N, K
    g.A, g.lda, g.B, g.ldb, g.C, g.ldc = a.data_ptr(), lda, b.data_ptr(), ldb, out.data_ptr(), ldc
    if bias is not None:
        if bias.dtype != torch.float32 or bias.numel() != N:
            raise ValueError("vitgan_b200.gemm: bias must be fp32 of length N")
        g.bias = bias.data_ptr()
    g.act, g.act_param = act, float(act_param)
    for name, t in (("aux", aux), ("residual", residual)):
        if t is not None and t.dtype != out.dtype:
            raise TypeError(f"vitgan_b200.gemm: `{name}` dtype {t.dtype} must equal the output dtype {out.dtype}")
    if aux is not None:
        g.aux, g.ldaux = aux.data_ptr(), _rowmajor2d(aux, "aux")
    if residual is not None:
        g.residual, g.ldres = residual.data_ptr(), _rowmajor2d(residual, "residual")
    if pre is not None:
        g.c_pre, g.ldpre = pre.data_ptr(), N
    g.c_row_group, g.res_row_mod, g.res_row_off, g.accumulate = c_row_group, res_row_mod, res_row_off, int(accumulate)
    if rowsum_out is not None:
        if rowsum_out.dtype != torch.float32 or rowsum_out.numel() != M or not rowsum_out.is_contiguous():
            raise ValueError("vitgan_b200.gemm: rowsum_out must be a contiguous fp32 vector of length M")
        g.a_rowsum = rowsum_out.data_ptr()
    ln_res = None
    if ln is not None:
        gamma, beta, eps = ln
        ln_out = torch.empty((M, N), dtype=out.dtype, device=a.device)
        mean = torch.empty(M, dtype=torch.float32, device=a.device)
        rstd = torch.empty_like(mean)
        g.ln_gamma, g.ln_beta, g.ln_out, g.ld_ln = gamma.data_ptr(), beta.data_ptr(), ln_out.data_ptr(), N
        g.ln_mean, g.ln_rstd, g.ln_eps = mean.data_ptr(), rstd.data_ptr(), float(eps)
        ln_res = (ln_out, mean, rstd)
    check(lib.vg_gemm(C.byref(g), stream()), "vg_gemm")
    _count()
    if ln_res is not None:
        return (out,) + ln_res
    return (out, pre) if want_pre else out


def cast(src: torch.Tensor, dtype: torch.dtype, out=None, num=None, den=None):
    """out = dtype(src * num/den)  (num/den: optional 1-element fp32 device tensors)."""
    _req(src, "src")
    src = src.contiguous()
    if out is None:
        out = torch.empty(src.shape, dtype=dtype, device=src.device)
    check(lib.vg_cast_scale(src.data_ptr(), dt(src), out.data_ptr(), dt(out), src.numel(), _ptr(num), _ptr(den), stream()),
          "vg_cast_scale")
    _count()
    return out


def pack_pad(ptrs: torch.Tensor, n: int, rows: int, cols: int, rows_pad: int, cols_pad: int, dtype: torch.dtype, num=None, den=None):
    """ptrs: int64 device tensor of n fp32 [rows, cols] tensor addresses -> [n * rows_pad, cols_pad] of `dtype`, scaled by
    num[t] / den[t] and zero padded (one launch)."""
    _req(ptrs, "ptrs")
    out = torch.empty(n * rows_pad, cols_pad, dtype=dtype, device=ptrs.device)
    check(lib.vg_pack_pad(ptrs.data_ptr(), n, rows, cols, rows_pad, cols_pad, _ptr(num), _ptr(den), out.data_ptr(), dt(out), stream()),
          "vg_pack_pad")
    _count()
    return out


def colsum(x2d: torch.Tensor, out=None):
    _req(x2d, "x")
    ld = _rowmajor2d(x2d, "x")
    if out is None:
        out = torch.zeros(x2d.shape[1], dtype=torch.float32, device=x2d.device)
    ws, ws_rows, counter = _reduce_scratch(x2d.device, x2d.shape[1])
    check(lib.vg_colsum(x2d.data_ptr(), dt(x2d), x2d.shape[0], x2d.shape[1], ld, out.data_ptr(), _ptr(ws), ws_rows,
                        counter.data_ptr(), stream()), "vg_colsum")
    _count()
    return out


def layernorm_fwd(x2d, gamma, beta, eps=1e-5):
    _req(x2d, "x")
    rows, e = x2d.shape
    y = torch.empty_like(x2d)
    mean = torch.empty(rows, dtype=torch.float32, device=x2d.device)
    rstd = torch.empty_like(mean)
    check(lib.vg_layernorm_fwd(dt(x2d), rows, e, x2d.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(),
                               mean.data_ptr(), rstd.data_ptr(), eps, stream()), "vg_layernorm_fwd")
    _count()
    return y, mean, rstd


def layernorm_bwd(dy2d, x2d, mean, rstd, gamma, dres=None, dgamma_acc=None, dbeta_acc=None, dres_colsum=None, dx_colsum=None,
                  dx_only=False):
    """dgamma_acc / dbeta_acc: optional fp32 buffers the kernel atomically ADDS into (e.g. the parameters' .grad).
    dres_colsum / dx_colsum: optional fp32 [E] buffers accumulating the column sums of `dres` / of the returned dx.
    dx_only: no parameter gradients at all (returns (dx, None, None))."""
    _req(x2d, "x2d")
    rows, e = x2d.shape
    dx = torch.empty_like(x2d)
    if dx_only:
        check(lib.vg_layernorm_bwd(dt(x2d), rows, e, dy2d.data_ptr(), x2d.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                   gamma.data_ptr(), _ptr(dres), dx.data_ptr(), None, None, None, None, None, 0, None, stream()),
              "vg_layernorm_bwd")
        _count()
        return dx, None, None
    if dgamma_acc is None:
        dgb = torch.zeros(2, e, dtype=torch.float32, device=x2d.device)
        dgamma_acc, dbeta_acc = dgb[0], dgb[1]
    ws, ws_rows, counter = _reduce_scratch(x2d.device, 4 * e)
    check(lib.vg_layernorm_bwd(dt(x2d), rows, e, dy2d.data_ptr(), x2d.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                               gamma.data_ptr(), _ptr(dres), dx.data_ptr(), dgamma_acc.data_ptr(), dbeta_acc.data_ptr(), _ptr(dres_colsum),
                               _ptr(dx_colsum), _ptr(ws), ws_rows, counter[4:].data_ptr(), stream()),
          "vg_layernorm_bwd")
    _count()
    return dx, dgamma_acc, dbeta_acc


_MAX_PARTS = {}


def layernorm_bwd_partials(dy2d, x2d, mean, rstd, gamma, dres=None):
    """LayerNorm backward whose column reductions are deferred: returns (dx, partials[n, 4E]); see vg_fold_partials."""
    _req(x2d, "x2d")
    rows, e = x2d.shape
    idx = x2d.device.index if x2d.device.index is not None else torch.cuda.current_device()
    mp = _MAX_PARTS.get(idx)
    if mp is None:
        mp = _MAX_PARTS[idx] = 2 * torch.cuda.get_device_properties(idx).multi_processor_count
    dx = torch.empty_like(x2d)
    part = torch.empty(mp, 4 * e, dtype=torch.float32, device=x2d.device)
    n = lib.vg_layernorm_bwd_partials(dt(x2d), rows, e, dy2d.data_ptr(), x2d.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                      gamma.data_ptr(), _ptr(dres), dx.data_ptr(), part.data_ptr(), mp, stream())
    if n <= 0:
        check(n if n < 0 else -1, "vg_layernorm_bwd_partials")
    _count()
    return dx, part[:n]


def fold_partials(part, e, out0, out1, out2=None, out3=None):
    check(lib.vg_fold_partials(part.data_ptr(), part.shape[0], e, _ptr(out0), _ptr(out1), _ptr(out2), _ptr(out3), stream()), "vg_fold_partials")
    _count()


def sln_fwd(h2d, w2d, ln_g, ln_b, gamma_s, beta_s, eps=1e-5):
    _req(w2d, "w2d")
    rows, f = w2d.shape
    y = torch.empty_like(w2d)
    mean = torch.empty(rows, dtype=torch.float32, device=w2d.device)
    rstd = torch.empty_like(mean)
    check(lib.vg_sln_fwd(dt(w2d), rows, h2d.shape[0], f, h2d.data_ptr(), w2d.data_ptr(), ln_g.data_ptr(), ln_b.data_ptr(),
                         gamma_s.data_ptr(), beta_s.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), eps, stream()),
          "vg_sln_fwd")
    _count()
    return y, mean, rstd


def sln_bwd(dy2d, h2d, w2d, mean, rstd, ln_g, ln_b, gamma_s, beta_s, dh_res=None, dw_res=None):
    _req(w2d, "w2d")
    rows, f = w2d.shape
    h_rows = h2d.shape[0]
    bcast = h_rows < rows
    dh = torch.zeros(h_rows, f, dtype=torch.float32, device=w2d.device) if bcast else torch.empty_like(h2d)
    dw = torch.empty_like(w2d)
    small = torch.zeros(2 + 2 * f, dtype=torch.float32, device=w2d.device)   # dgamma_s, dbeta_s, dln_g, dln_b
    check(lib.vg_sln_bwd(dt(w2d), rows, h_rows, f, dy2d.data_ptr(), h2d.data_ptr(), w2d.data_ptr(), mean.data_ptr(),
                         rstd.data_ptr(), ln_g.data_ptr(), ln_b.data_ptr(), gamma_s.data_ptr(), beta_s.data_ptr(),
                         _ptr(dh_res), _ptr(dw_res), dh.data_ptr(), dw.data_ptr(), small[0:1].data_ptr(),
                         small[1:2].data_ptr(), small[2:2 + f].data_ptr(), small[2 + f:].data_ptr(), stream()), "vg_sln_bwd")
    _count()
    return dh, dw, small[0:1], small[1:2], small[2:2 + f], small[2 + f:]


def attention_fwd(q, k, v, B, H, S, d, scale, mode=L.ATTN_DOT, ld_qkv=None, out=None):
    """q,k,v: 2-D row-major views [B*S, >=H*d] sharing the leading dim (slices of a fused projection output)."""
    _req(q, "q")
    ld = ld_qkv or q.stride(0)
    if out is None:
        out = torch.empty(B * S, H * d, dtype=q.dtype, device=q.device)
    lse = torch.empty(B * H * S, dtype=torch.float32, device=q.device)
    check(lib.vg_attention_fwd(dt(q), mode, B, H, S, d, q.data_ptr(), k.data_ptr(), v.data_ptr(), ld, out.data_ptr(),
                               out.stride(0), lse.data_ptr(), scale, stream()), "vg_attention_fwd")
    _count()
    return out, lse


def attention_bwd(q, k, v, o, d_o, lse, B, H, S, d, scale, mode=L.ATTN_DOT, dqkv=None):
    """Returns dqkv [B*S, 3*H*d] laid out like a fused projection output (dq | dk | dv)."""
    _req(q, "q")
    hd = H * d
    if dqkv is None:
        dqkv = torch.empty(B * S, 3 * hd, dtype=q.dtype, device=q.device)
    delta = torch.empty(B * H * S, dtype=torch.float32, device=q.device)
    check(lib.vg_attention_bwd(dt(q), mode, B, H, S, d, q.data_ptr(), k.data_ptr(), v.data_ptr(), q.stride(0), o.data_ptr(),
                               d_o.data_ptr(), o.stride(0), lse.data_ptr(), dqkv.data_ptr(), dqkv[:, hd:].data_ptr(),
                               dqkv[:, 2 * hd:].data_ptr(), dqkv.stride(0), scale, delta.data_ptr(), stream()),
          "vg_attention_bwd")
    _count(3)
    return dqkv


def im2col(img: torch.Tensor, P: int, dtype: torch.dtype):
    _req(img, "img")
    B, Cc, I, _ = img.shape
    img = img.contiguous().float()
    out = torch.empty(B * (I // P) ** 2, Cc * P * P, dtype=dtype, device=img.device)
    check(lib.vg_im2col_patches(_DT[dtype], B, Cc, I, P, img.data_ptr(), out.data_ptr(), stream()), "vg_im2col_patches")
    _count()
    return out


def col2im(dpatches: torch.Tensor, B, Cc, I, P):
    _req(dpatches, "dpatches")
    dimg = torch.empty(B, Cc, I, I, dtype=torch.float32, device=dpatches.device)
    check(lib.vg_col2im_patches(dt(dpatches), B, Cc, I, P, dpatches.data_ptr(), dimg.data_ptr(), stream()), "vg_col2im_patches")
    _count()
    return dimg


def v1_tokens_fwd(img, win, stride, n_side, dtype):
    _req(img, "img")
    B, Cc, I, _ = img.shape
    img = img.contiguous().float()
    out = torch.empty(B * n_side * n_side, Cc * win * win, dtype=dtype, device=img.device)
    check(lib.vg_v1_tokens_fwd(_DT[dtype], B, Cc, I, win, stride, n_side, img.data_ptr(), out.data_ptr(), stream()), "vg_v1_tokens_fwd")
    _count()
    return out


def v1_tokens_bwd(dtokens, B, Cc, I, win, stride, n_side):
    _req(dtokens, "dtokens")
    dimg = torch.zeros(B, Cc, I, I, dtype=torch.float32, device=dtokens.device)
    check(lib.vg_v1_tokens_bwd(dt(dtokens), B, Cc, I, win, stride, n_side, dtokens.data_ptr(), dimg.data_ptr(), stream()), "vg_v1_tokens_bwd")
    _count()
    return dimg


def fill_rows(x3d, row, v, v2=None):
    _req(x3d, "x3d")
    B, S, E = x3d.shape
    check(lib.vg_fill_rows(dt(x3d), B, S, E, row, v.data_ptr(), _ptr(v2), x3d.data_ptr(), stream()), "vg_fill_rows")
    _count()


def embed_bwd_split(dx3d, pos_has_cls: bool):
    _req(dx3d, "dx3d")
    B, S, E = dx3d.shape
    dtok = torch.empty(B * (S - 1), E, dtype=dx3d.dtype, device=dx3d.device)
    dcls = torch.zeros(E, dtype=torch.float32, device=dx3d.device)
    dpos = torch.zeros(S if pos_has_cls else S - 1, E, dtype=torch.float32, device=dx3d.device)
    check(lib.vg_embed_bwd_split(dt(dx3d), B, S, E, dx3d.data_ptr(), dtok.data_ptr(), dcls.data_ptr(), dpos.data_ptr(),
                                 int(pos_has_cls), stream()), "vg_embed_bwd_split")
    _count()
    return dtok, dcls, dpos


def copy_rows(src, ld_src, rows, cols, dst, ld_dst):
    _req(src, "src")
    check(lib.vg_copy_rows(dt(src), rows, cols, src.data_ptr(), ld_src, dst.data_ptr(), ld_dst, stream()), "vg_copy_rows")
    _count()


def act_backward(dy, aux, act, act_param=0.0):
    _req(dy, "dy")
    if dy.dtype != aux.dtype:
        raise TypeError("vitgan_b200.act_backward: dy/aux dtypes differ")
    out = torch.empty_like(dy)
    check(lib.vg_act_backward(dt(dy), dy.numel(), dy.data_ptr(), aux.data_ptr(), act, float(act_param), out.data_ptr(), stream()),
          "vg_act_backward")
    _count()
    return out


def add_(x, y):
    _req(x, "x")
    check(lib.vg_add_inplace(dt(x), x.data_ptr(), y.data_ptr(), x.numel(), stream()), "vg_add_inplace")
    _count()
    return x


def broadcast_rows(src2d, reps):
    _req(src2d, "src2d")
    out = torch.empty(reps * src2d.shape[0], src2d.shape[1], dtype=src2d.dtype, device=src2d.device)
    check(lib.vg_broadcast_rows(dt(src2d), src2d.data_ptr(), src2d.shape[0], src2d.shape[1], out.data_ptr(), reps, stream()),
          "vg_broadcast_rows")
    _count()
    return out


def sigma_max(mat_ptrs: torch.Tensor, n_mats, rows, cols, u_state, n_iters, out=None):
    """mat_ptrs: int64 device tensor of fp32 matrix addresses."""
    if out is None:
        out = torch.empty(n_mats, dtype=torch.float32, device=u_state.device)
    check(lib.vg_sigma_max(mat_ptrs.data_ptr(), n_mats, rows, cols, u_state.data_ptr(), n_iters, out.data_ptr(), stream()),
          "vg_sigma_max")
    _count()
    return out


def adam_step(p, g, m, v, step_count, lr, b1, b2, eps, wd, decoupled, grad_scale=1.0, shadow=None):
    check(lib.vg_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, b1, b2, eps, wd,
                           int(decoupled), grad_scale, step_count.data_ptr(), _ptr(shadow), stream()), "vg_adam_step")
    _count(2)


def softmax_ce(logits, targets, rows_per_group):
    """-> (losses[G], dlogits[rows, C]) of the fused cross-entropy head (see vg_softmax_ce)."""
    _req(logits, "logits")
    if logits.dtype != torch.float32 or targets.dtype != torch.int64 or not logits.is_contiguous():
        raise TypeError("vitgan_b200.softmax_ce: contiguous fp32 logits and int64 targets required")
    rows, c = logits.shape
    losses = torch.empty(rows // rows_per_group, dtype=torch.float32, device=logits.device)
    dlog = torch.empty_like(logits)
    check(lib.vg_softmax_ce(logits.data_ptr(), targets.data_ptr(), rows, c, rows_per_group, losses.data_ptr(), dlog.data_ptr(), stream()),
          "vg_softmax_ce")
    _count()
    return losses, dlog


def bce(prob, target, rows_per_group):
    """-> (losses[G], dprob[rows]) of the fused binary-cross-entropy head (see vg_bce)."""
    _req(prob, "prob")
    if prob.dtype != torch.float32 or target.dtype != torch.float32 or not prob.is_contiguous() or not target.is_contiguous():
        raise TypeError("vitgan_b200.bce: contiguous fp32 probabilities and targets required")
    if prob.numel() != target.numel():
        raise ValueError("vitgan_b200.bce: probabilities and targets differ in size")
    rows = prob.numel()
    losses = torch.empty(rows // rows_per_group, dtype=torch.float32, device=prob.device)
    dprob = torch.empty_like(prob)
    check(lib.vg_bce(prob.data_ptr(), target.data_ptr(), rows, rows_per_group, losses.data_ptr(), dprob.data_ptr(), stream()), "vg_bce")
    _count()
    return losses, dprob


def denorm_u8(x):
    """uint8(clamp(x * 127.5 + 127.5, 0, 255)) -- utils.convert_to_uint8 of the reference, same shape as x."""
    _req(x, "x")
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    check(lib.vg_denorm_u8(dt(x), x.data_ptr(), x.numel(), out.data_ptr(), stream()), "vg_denorm_u8")
    _count()
    return out
