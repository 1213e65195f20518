"""torch.autograd.Function wrappers: one per reference block, forward and backward both made of this
library's sm_100a kernels (ops.py -> C ABI).  Parameters stay the reference's fp32 nn.Parameters with the
reference's names/shapes; packed / bf16 operand copies are built at call time and cached per parameter
version (state_dict must not change).

Precision: ``set_precision("bf16")`` (fast path: bf16 activations and GEMM operands on tcgen05, fp32
accumulate/statistics, 2e-2 tolerance) or ``"fp32"`` (parity path: fp32 everywhere on CUDA cores, 1e-4).
Double backward is not supported (the reference's active path never needs it).
"""
from __future__ import annotations

import math
import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import lib as L
from . import ops

_PRECISION = "bf16"
_SKIP_PARAM_GRADS = False


def set_precision(p: str):
    global _PRECISION
    if p not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    _PRECISION = p


def get_precision() -> str:
    return _PRECISION


def act_dtype() -> torch.dtype:
    return torch.bfloat16 if _PRECISION == "bf16" else torch.float32


class skip_param_grads:
    """Context: blocks return None for parameter gradients (used for the discriminator during the generator
    update, where the reference computes D weight gradients only to zero them at the next iteration:
    src/v2/training.py:177,204-210; src/v1/gan.py:222,247-251)."""

    def __init__(self, enabled=True):
        self.enabled = enabled

    def __enter__(self):
        global _SKIP_PARAM_GRADS
        self.prev, _SKIP_PARAM_GRADS = _SKIP_PARAM_GRADS, self.enabled

    def __exit__(self, *a):
        global _SKIP_PARAM_GRADS
        _SKIP_PARAM_GRADS = self.prev


# --------------------------------------------------------------------------------------------------
# operand cache: packed (and, on the fast path, bf16) copies of parameters, rebuilt when a parameter's
# version counter changes (optimizer.step()).  Disable while capturing CUDA graphs so the casts are captured.
# --------------------------------------------------------------------------------------------------
_cache: dict = {}
_cache_enabled = True


# --------------------------------------------------------------------------------------------------
# dropout policy.  The reference applies nn.Dropout in training mode (src/v2/modules.py:99,179-180, p = 0.1;
# src/v1/transformer.py:42,86 and muilti_layer_perceptron.py:27, p = 0.2).  The fused blocks do not implement it, so a
# module in training mode whose dropout probability is > 0 is an ERROR by default ("unsupported => error") --
# never a silently different model.  "off" is the parity protocol of SURVEY Q11 (p = 0 on both sides) made explicit.
# --------------------------------------------------------------------------------------------------
_DROPOUT_POLICY = "error"


def set_dropout_policy(policy: str):
    """'error' (default): raise when a patched block in training mode owns an nn.Dropout with p > 0;
    'off': treat every dropout as p = 0 (what zeroing nn.Dropout.p on both sides does in the parity runs)."""
    global _DROPOUT_POLICY
    if policy not in ("error", "off"):
        raise ValueError("dropout policy must be 'error' or 'off'")
    _DROPOUT_POLICY = policy


def check_dropout(module, *extra):
    """Raise if `module` (training mode) owns an active nn.Dropout, directly or through `extra` sub-modules."""
    if _DROPOUT_POLICY == "off" or not module.training:
        return
    for holder in (module,) + extra:
        for name, child in holder._modules.items():
            if isinstance(child, torch.nn.Dropout) and child.p > 0:
                raise NotImplementedError(
                    f"vitgan_b200: {type(module).__name__}.{name} is nn.Dropout(p={child.p}) in training mode, which the fused CUDA "
                    "blocks do not implement.  Set p = 0 (as the parity protocol does), call .eval(), or "
                    "vitgan_b200.set_dropout_policy('off') to run the blocks without dropout explicitly.")


def operand_cache_enabled() -> bool:
    return _cache_enabled


def set_operand_cache(enabled: bool):
    global _cache_enabled
    _cache_enabled = enabled
    _cache.clear()


def invalidate_operands(param_ids):
    """Drop cached operand copies that involve any of the given parameter ids (after an out-of-band update)."""
    for key in [k for k in _cache if any(i in param_ids for i in k[0])]:
        del _cache[key]


# parameters re-homed by train.FlatNet: id(param) -> (flatnet, element offset).  When the parameters asked for sit back to
# back in the flat buffer, the packed operand is a VIEW of the flat fp32 buffer or of its bf16 shadow (kept in sync by the
# fused Adam kernel): no cast, no copy, no launch -- also under CUDA-graph capture, where the version-keyed cache is off.
# The registry holds WEAK references and every lookup re-checks that the parameter's storage really is the slot it was
# registered with: ids of freed Parameters get recycled, and a stale entry must never hand out another network's buffer.
flat_registry: dict = {}       # id(param) -> (weakref to the FlatNet, element offset)


def register_flat(net, params, offsets):
    import weakref
    ref = weakref.ref(net)
    ids = [id(p) for p in params]
    for i, o in zip(ids, offsets):
        flat_registry[i] = (ref, o)

    def _drop(ids=ids, ref=ref):
        for i in ids:
            if flat_registry.get(i, (None,))[0] is ref:
                del flat_registry[i]
    weakref.finalize(net, _drop)


def _flat_entry(p):
    ent = flat_registry.get(id(p))
    if ent is None:
        return None
    net = ent[0]()
    if net is None or p.data_ptr() != net.flat_param.data_ptr() + 4 * ent[1]:      # dead network, or a recycled id
        return None
    return net, ent[1]


def _flat_view(params, dtype):
    ent = _flat_entry(params[0])
    if ent is None:
        return None
    net, off0 = ent
    buf = net.flat_param if dtype == torch.float32 else (net.flat_shadow if dtype == torch.bfloat16 else None)
    if buf is None or not net.shadow_valid(params):
        return None
    off = off0
    for p in params:
        e = _flat_entry(p)
        if e is None or e[0] is not net or e[1] != off:
            return None
        off += p.numel()
    rows = sum(p.shape[0] if p.dim() > 1 else p.numel() for p in params)
    return buf[off0:off].view(rows, -1)


def packed(params, dtype, cols=None, scales=None):
    """Stack 2-D (or flattenable) fp32 parameters along dim 0 into one [sum(rows), cols] tensor of `dtype`.
    scales: optional list of (num, den) 1-element device tensors (spectral rescale) per parameter."""
    if len(params) == 1 and params[0].dtype == dtype and scales is None:
        p = params[0].detach()
        return p if p.dim() == 2 else p.reshape(p.shape[0], -1)
    if scales is None:
        v = _flat_view(params, dtype)
        if v is not None:
            return v if params[0].dim() != 1 else v
    key = (tuple(id(p) for p in params), dtype)
    ver = tuple((p._version, p.data_ptr()) for p in params)
    if _cache_enabled and scales is None:
        hit = _cache.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
    with torch.no_grad():
        flat = [p.detach().reshape(p.shape[0], -1) if p.dim() != 1 else p.detach().reshape(-1, 1) for p in params]
        cols = flat[0].shape[1]
        out = torch.empty(sum(f.shape[0] for f in flat), cols, dtype=dtype, device=flat[0].device)
        r = 0
        for i, f in enumerate(flat):
            num, den = scales[i] if scales is not None else (None, None)
            ops.cast(f, dtype, out=out[r:r + f.shape[0]], num=num, den=den)
            r += f.shape[0]
    if _cache_enabled and scales is None:
        _cache[key] = (ver, out)
    return out


def packed_vec(params):
    """Concatenate fp32 bias vectors (stay fp32)."""
    if len(params) == 1:
        return params[0].detach()
    v = _flat_view(params, torch.float32)
    if v is not None:
        return v.reshape(-1)
    return packed(params, torch.float32).reshape(-1)


def _want(ctx, idx):
    return ctx.needs_input_grad[idx] and not ctx.skip_pg


# --------------------------------------------------------------------------------------------------
# fused gradient accumulation: when a parameter already owns a contiguous fp32 .grad (train.FlatNet, or any
# optimizer used with zero_grad(set_to_none=False)), the split-K weight-gradient GEMM / bias column-sum /
# LayerNorm reductions ADD straight into it (they are atomic-accumulating kernels anyway) and the Function
# returns None for that input -> no temporary, no zero-fill, no autograd add kernel per parameter.
# Data-parallel bucketing is told through `grad_ready_hooks` since autograd hooks do not fire for None grads.
# --------------------------------------------------------------------------------------------------
_FUSE_GRAD_ACC = True
grad_ready_hooks: list = []


def set_fused_grad_accumulation(enabled: bool):
    global _FUSE_GRAD_ACC
    _FUSE_GRAD_ACC = enabled


def _grad_view(params, rows, cols):
    """A [rows, cols] fp32 view over the .grad of `params` if they exist and are laid out back to back, else None."""
    if not _FUSE_GRAD_ACC:
        return None
    g0 = params[0].grad
    if g0 is None or g0.dtype != torch.float32 or not g0.is_contiguous():
        return None
    ptr = g0.data_ptr()
    for p in params:
        g = p.grad
        if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.data_ptr() != ptr:
            return None
        ptr += 4 * g.numel()
    if len(params) == 1:
        return g0.view(rows, cols)
    return g0.as_strided((rows, cols), (cols, 1))


def _notify(params):
    for hook in grad_ready_hooks:
        for p in params:
            hook(p)


# --------------------------------------------------------------------------------------------------
# parameter-gradient side stream: the weight-gradient GEMMs feed nothing but .grad, so they need not sit on the
# dgrad critical path.  When enabled they are launched on a second stream that forks from the current one (the
# GEMM's inputs are ready there) and are joined back by join_param_grad_stream() -- which the step calls before the
# optimizer / gradient all-reduce.  The kernels of the two streams fill each other's launch / tail gaps (the E=128
# kernels leave SMs idle ~35 % of their duration).  Inputs are kept alive until the join instead of record_stream()
# so that the caching allocator cannot recycle them early -- also correct under whole-step graph capture.
# Off by default: code that calls backward() and then a torch optimizer directly has no join point.
# --------------------------------------------------------------------------------------------------
_PG_STREAM_ON = False
_pg_streams: dict = {}
_pg_keepalive: list = []
_pg_forked: set = set()


def set_param_grad_stream(enabled: bool):
    global _PG_STREAM_ON
    join_param_grad_stream()
    _PG_STREAM_ON = enabled


def param_grad_stream_enabled() -> bool:
    return _PG_STREAM_ON


def _pg_side(dev):
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _pg_streams.get(idx)
    if st is None:
        st = _pg_streams[idx] = torch.cuda.Stream(device=idx)
    return idx, st


def join_param_grad_stream():
    """The current stream waits for every weight-gradient GEMM launched on the side stream since the last join."""
    for idx in list(_pg_forked):
        torch.cuda.current_stream(idx).wait_stream(_pg_streams[idx])
    _pg_forked.clear()
    _pg_keepalive.clear()


def _wgrad_gemm(dy2, x2, view, rowsum=None):
    if not _PG_STREAM_ON:
        return ops.gemm(dy2, x2, trans_a=True, trans_b=False, accumulate=True, out=view, rowsum_out=rowsum)
    idx, side = _pg_side(dy2.device)
    side.wait_stream(torch.cuda.current_stream(idx))
    with torch.cuda.stream(side):
        ops.gemm(dy2, x2, trans_a=True, trans_b=False, accumulate=True, out=view, rowsum_out=rowsum)
    _pg_forked.add(idx)
    _pg_keepalive.append((dy2, x2))
    return view


_FUSE_BIAS_IN_WGRAD = True     # bias gradient as an extra N=16 MMA against ones inside the wgrad GEMM (vg_gemm a_rowsum)
_wgrad_bias_unsupported: set = set()


def acc_wgrad(dy2, x2, params, bias_params=None):
    """dW = dy2^T x2 for one parameter or a row-stack of parameters.  Returns the gradient tensor ([sum N, K]),
    or None if it was accumulated in place into the parameters' .grad.
    bias_params: the matching bias parameter(s); their gradient colsum(dy2) is then produced by the same GEMM when both
    gradients accumulate in place on the tcgen05 path -- returns (None, None) in that case, else (acc_wgrad(...), acc_colsum(...))."""
    n = sum(p.shape[0] for p in params)
    view = _grad_view(params, n, x2.shape[1])
    if bias_params is None:
        if view is not None:
            _wgrad_gemm(dy2, x2, view)
            _notify(params)
            return None
        return ops.gemm(dy2, x2, trans_a=True, trans_b=False, accumulate=True)
    bview = _grad_view(bias_params, 1, dy2.shape[1]) if view is not None else None
    key = (dy2.dtype, dy2.shape[1], x2.shape[1])
    # The fused bias gradient (a_rowsum: 16 more TMEM columns) only exists in the 128-wide GEMM.  Compute-bound weight gradients
    # (the scaled config: K = 65 792 tokens, N = 768 a multiple of 256) run ~1.6x faster on the 256-wide tile, whose two accumulator
    # stages fill the TMEM -- there the bias gradient is the memory-bound column-sum kernel, issued like the GEMM on the
    # parameter-gradient stream where it overlaps compute-bound work.  Mirrors the tile heuristic of gemm_tc_launch.
    wide = (x2.shape[1] % 256 == 0 and dy2.shape[0] >= 4096 and ((dy2.shape[1] + 127) // 128) * (x2.shape[1] // 256) >= 8)
    if _FUSE_BIAS_IN_WGRAD and not wide and bview is not None and dy2.dtype == torch.bfloat16 and key not in _wgrad_bias_unsupported:
        try:
            _wgrad_gemm(dy2, x2, view, bview.view(-1))
            _notify(params)
            _notify(bias_params)
            return None, None
        except L.VitganError as ex:    # rejected before any launch (shape / alignment): separate column-sum kernel
            if not ex.rejected:
                raise                  # a launch / CUDA error is not a reason to try another path
            _wgrad_bias_unsupported.add(key)
    return acc_wgrad(dy2, x2, params), acc_colsum(dy2, bias_params)


def acc_colsum(dy2, params):
    view = _grad_view(params, 1, dy2.shape[1])
    if view is not None:
        if _PG_STREAM_ON and dy2.shape[1] > 256:    # feeds nothing but .grad: off the critical path, like the weight-gradient GEMMs
            # (wider than 256 columns the kernel takes no shared replica workspace, so it may overlap a main-stream column sum)
            idx, side = _pg_side(dy2.device)
            side.wait_stream(torch.cuda.current_stream(idx))
            with torch.cuda.stream(side):
                ops.colsum(dy2, out=view.view(-1))
            _pg_forked.add(idx)
            _pg_keepalive.append((dy2,))
        else:
            ops.colsum(dy2, out=view.view(-1))
        _notify(params)
        return None
    return ops.colsum(dy2)


def _split_rows(t, sizes):
    if t is None:
        return (None,) * len(sizes)
    out, r = [], 0
    for n in sizes:
        out.append(t[r:r + n])
        r += n
    return tuple(out)


def _acc_target(params):
    """-> (buffer to accumulate a bias gradient into, in_place?)"""
    n = sum(p.numel() for p in params)
    view = _grad_view(params, 1, n)
    if view is not None:
        return view.view(-1), True
    return torch.zeros(n, dtype=torch.float32, device=params[0].device), False


def ln_bwd(dy2, x2, mean, rstd, weight, bias, dres, want_pg, bias_of_dres=None, bias_of_dx=None):
    """LayerNorm backward; returns (dx, dweight|None, dbias|None, d_bias_of_dres|None, d_bias_of_dx|None).
    bias_of_dres / bias_of_dx: bias parameters whose gradient is colsum(dres) / colsum(dx) (fused into this kernel).
    Gradients are accumulated in place into .grad when possible (then None is returned for them)."""
    if not want_pg:
        dx, _, _ = ops.layernorm_bwd(dy2, x2, mean, rstd, weight.detach(), dres=dres, dx_only=True)
        return dx, None, None, None, None
    gw, w_in = _acc_target([weight])
    gb, b_in = _acc_target([bias])
    cr, cr_in = _acc_target([bias_of_dres]) if bias_of_dres is not None else (None, True)
    cx, cx_in = _acc_target([bias_of_dx]) if bias_of_dx is not None else (None, True)
    if _PG_STREAM_ON and x2.shape[1] <= 128 and w_in and b_in and cr_in and cx_in:
        # deferred reductions: per-CTA partials now, folded into .grad on the parameter-gradient stream
        dx, part = ops.layernorm_bwd_partials(dy2, x2, mean, rstd, weight.detach(), dres=dres)
        idx, side = _pg_side(x2.device)
        side.wait_stream(torch.cuda.current_stream(idx))
        with torch.cuda.stream(side):
            ops.fold_partials(part, x2.shape[1], gw, gb, cr, cx)
        _pg_forked.add(idx)
        _pg_keepalive.append(part)
    else:
        dx, _, _ = ops.layernorm_bwd(dy2, x2, mean, rstd, weight.detach(), dres=dres, dgamma_acc=gw, dbeta_acc=gb,
                                     dres_colsum=cr, dx_colsum=cx)
    _notify([p for p, flag in ((weight, w_in), (bias, b_in), (bias_of_dres, cr_in), (bias_of_dx, cx_in)) if p is not None and flag])
    return (dx, None if w_in else gw, None if b_in else gb, None if cr_in else cr, None if cx_in else cx)


# --------------------------------------------------------------------------------------------------
# Linear (+ activation): y = act(x W^T + b)
# --------------------------------------------------------------------------------------------------
_ACT_BWD = {L.ACT_GELU: L.ACT_MUL_DGELU, L.ACT_TANH: L.ACT_MUL_DTANH, L.ACT_SIN: L.ACT_MUL_DSIN,
            L.ACT_SIGMOID: L.ACT_MUL_DSIGMOID}


class LinearFn(Function):
    """nn.Linear (+ fused activation).  x: (..., K) activation dtype or fp32; returns (..., N) in `out_dtype`.
    compute='act' -> operands in the activation dtype; compute='fp32' -> fp32 CUDA-core GEMM (tiny heads)."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, act_param, compute, out_dtype):
        cdt = act_dtype() if compute == "act" else torch.float32
        x2 = x.reshape(-1, x.shape[-1])
        if x2.dtype != cdt:
            x2 = ops.cast(x2, cdt)
        w = packed([weight], cdt)
        odt = out_dtype or cdt
        need_pre = act in (L.ACT_GELU, L.ACT_SIN)
        res = ops.gemm(x2, w, bias=None if bias is None else bias.detach(), act=act, act_param=act_param,
                       want_pre=need_pre, out_dtype=odt)
        y, pre = res if need_pre else (res, None)
        ctx.save_for_backward(x2, weight, pre if need_pre else (y if act != L.ACT_NONE else None))
        ctx.meta = (act, act_param, cdt, x.shape, x.dtype, bias is not None)
        ctx.bias_ref = bias
        ctx.skip_pg = _SKIP_PARAM_GRADS
        return y.reshape(*x.shape[:-1], weight.shape[0])

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2, weight, aux = ctx.saved_tensors
        act, act_param, cdt, xshape, xdtype, has_bias = ctx.meta
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        if act != L.ACT_NONE:   # dpre = dy * act'(.)
            dy2 = _act_backward(dy2, aux, act, act_param)
        if dy2.dtype != cdt:
            dy2 = ops.cast(dy2, cdt)
        w = packed([weight], cdt)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            if cdt == torch.float32 and w.shape[1] * 16 <= w.shape[0]:
                # tiny fan-in (G head 10 -> 3072): dX[M,10] = dY[M,3072] W has 8 output tiles and a long K: split-K
                dx = ops.gemm(dy2, w, trans_b=False, accumulate=True)
            else:
                dx = ops.gemm(dy2, w, trans_b=False, out_dtype=cdt)
            if dx.dtype != xdtype:
                dx = ops.cast(dx, xdtype)
            dx = dx.reshape(xshape)
        if _want(ctx, 1) and has_bias and _want(ctx, 2):
            dw, db = acc_wgrad(dy2, x2, [weight], [ctx.bias_ref])        # bias gradient from the same GEMM where supported
            dw = None if dw is None else dw.reshape(weight.shape)
        else:
            if _want(ctx, 1):
                dw = acc_wgrad(dy2, x2, [weight])
                dw = None if dw is None else dw.reshape(weight.shape)
            if has_bias and _want(ctx, 2):
                db = acc_colsum(dy2, [ctx.bias_ref])
        return dx, dw, db, None, None, None, None


def _act_backward(dy2, aux, act, act_param):
    """dpre = dy * act'(aux) for a standalone Linear(+activation) layer (vg_act_backward kernel)."""
    aux2 = aux.reshape(dy2.shape)
    if dy2.dtype != aux2.dtype:
        dy2 = ops.cast(dy2, aux2.dtype)
    return ops.act_backward(dy2, aux2.contiguous(), act, act_param)


# --------------------------------------------------------------------------------------------------
# LayerNorm
# --------------------------------------------------------------------------------------------------
class LayerNormFn(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        y, mean, rstd = ops.layernorm_fwd(x2, weight.detach(), bias.detach(), eps)
        ctx.save_for_backward(x2, mean, rstd, weight)
        ctx.bias_ref = bias
        ctx.skip_pg = _SKIP_PARAM_GRADS
        return y.reshape(x.shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2, mean, rstd, weight = ctx.saved_tensors
        dy2 = dy.reshape(x2.shape).contiguous()
        dx, dg, db, _, _ = ln_bwd(dy2, x2, mean, rstd, weight, ctx.bias_ref, None, not ctx.skip_pg)
        return dx.reshape(dy.shape), dg, db, None


# --------------------------------------------------------------------------------------------------
# v2 EmbedLayer: conv(k=s=P)+bias, +pos (patch rows), CLS row       src/v2/modules.py:82-100
# --------------------------------------------------------------------------------------------------
class EmbedV2Fn(Function):
    @staticmethod
    def forward(ctx, img, conv_w, conv_b, pos, cls, patch):
        adt = act_dtype()
        B, Cc, I, _ = img.shape
        E = conv_w.shape[0]
        N = (I // patch) ** 2
        patches = ops.im2col(img, patch, adt)                        # [B*N, C*P*P]
        w = packed([conv_w], adt)                                     # [E, C*P*P]
        posd = packed([pos], adt).reshape(N, E)
        x = torch.empty(B, N + 1, E, dtype=adt, device=img.device)
        ops.gemm(patches, w, bias=conv_b.detach(), residual=posd, res_row_mod=N, c_row_group=N, out=x.view(B * (N + 1), E))
        ops.fill_rows(x, 0, cls.detach().reshape(E))
        ctx.save_for_backward(patches, conv_w)
        ctx.bias_ref = conv_b
        ctx.meta = (B, Cc, I, patch, N, E, pos.shape, cls.shape)
        ctx.skip_pg = _SKIP_PARAM_GRADS
        return x

    @staticmethod
    @once_differentiable
    def backward(ctx, dx):
        patches, conv_w = ctx.saved_tensors
        B, Cc, I, patch, N, E, pos_shape, cls_shape = ctx.meta
        dtok, dcls, dpos = ops.embed_bwd_split(dx.contiguous(), pos_has_cls=False)
        dimg = dw = db = None
        if ctx.needs_input_grad[0]:
            w = packed([conv_w], dtok.dtype)
            dpatches = ops.gemm(dtok, w, trans_b=False)
            dimg = ops.col2im(dpatches, B, Cc, I, patch)
        if ctx.skip_pg:
            return dimg, None, None, None, None, None
        dw, db = acc_wgrad(dtok, patches, [conv_w], [ctx.bias_ref])
        dw = None if dw is None else dw.reshape(conv_w.shape)
        return dimg, dw, db, dpos.reshape(pos_shape), dcls.reshape(cls_shape), None


# --------------------------------------------------------------------------------------------------
# attention sub-layer helpers (shared by SelfAttentionFn and EncoderFn)
# --------------------------------------------------------------------------------------------------
# LayerNorm of a GEMM's output rows inside its epilogue (vg_gemm ln_*; N == 128 on the tcgen05 path): the out-proj GEMM also emits
# LayerNorm-2 of its block and the fc2 GEMM LayerNorm-1 of the NEXT block (v2.vit_forward chains the blocks).  Measured at C2 on
# B200: 4.57 -> 4.51 ms per step (33 launches fewer).  A small win only: with 1.76 tiles per CTA the epilogue is on the critical
# path of every tile and the two-pass statistics need two 64-thread barriers between the warps sharing a row.  VG_FUSE_LN=0 or
# set_fused_layernorm_epilogue(False) falls back to GEMM + LayerNorm kernel (parity-tested both ways).
_FUSE_LN_IN_GEMM = os.environ.get("VG_FUSE_LN", "1") != "0"
_gemm_ln_unsupported: set = set()


def fused_layernorm_epilogue() -> bool:
    return _FUSE_LN_IN_GEMM


def set_fused_layernorm_epilogue(enabled: bool):
    global _FUSE_LN_IN_GEMM
    _FUSE_LN_IN_GEMM = enabled


def gemm_ln(a, w, bias, residual, ln):
    """C = a w^T + bias + residual followed by LayerNorm(C): -> (C, (xn, mean, rstd)).  One launch when the GEMM epilogue can
    normalise the row (bf16, N == 128, tcgen05 path); otherwise the GEMM and the LayerNorm kernel."""
    gamma, beta, eps = ln
    key = (a.dtype, w.shape[0], w.shape[1])
    if _FUSE_LN_IN_GEMM and a.dtype == torch.bfloat16 and w.shape[0] == 128 and key not in _gemm_ln_unsupported:
        try:
            y, xn, mean, rstd = ops.gemm(a, w, bias=bias, residual=residual, ln=(gamma, beta, eps))
            return y, (xn, mean, rstd)
        except L.VitganError as ex:    # rejected before any launch
            if not ex.rejected:
                raise
            _gemm_ln_unsupported.add(key)
    y = ops.gemm(a, w, bias=bias, residual=residual)
    return y, ops.layernorm_fwd(y, gamma, beta, eps)


def _attn_fwd(xn, B, S, H, wqkv, bqkv, wo, bo, residual, scale, mode=L.ATTN_DOT, ln=None):
    """ln = (gamma, beta, eps) of the LayerNorm that follows the sub-layer: then a fifth result (xn, mean, rstd) is returned."""
    E3 = wqkv.shape[0]
    hd = E3 // 3
    d = hd // H
    qkv = ops.gemm(xn, wqkv, bias=bqkv)                                            # fused Q|K|V projection
    o, lse = ops.attention_fwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], B, H, S, d, scale, mode)
    if ln is not None:
        y, stats = gemm_ln(o, wo, bo, residual, ln)                                # out-proj (+bias, +skip) + next LayerNorm
        return y, qkv, o, lse, stats
    y = ops.gemm(o, wo, bias=bo, residual=residual)                                # out-proj (+bias, +skip)
    return y, qkv, o, lse


def _attn_bwd(dy, xn, qkv, o, lse, B, S, H, wqkv, wo, scale, want_pg, prm, mode=L.ATTN_DOT, bo_done=False):
    """prm = dict(wq, wk, wv, wo, bq, bk, bv, bo) of the nn.Parameters (targets of the in-place accumulation).
    bo_done: the out-proj bias gradient colsum(dy) was already produced by the LayerNorm backward that made dy."""
    hd = qkv.shape[1] // 3
    d = hd // H
    g = {"wo": None, "bo": None, "wqkv": None, "bqkv": None}
    if want_pg:
        g["wo"] = acc_wgrad(dy, o, [prm["wo"]])
        if not bo_done:
            g["bo"] = acc_colsum(dy, [prm["bo"]])
    d_o = ops.gemm(dy, wo, trans_b=False)
    dqkv = ops.attention_bwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], o, d_o, lse, B, H, S, d, scale, mode)
    if want_pg:
        g["wqkv"], g["bqkv"] = acc_wgrad(dqkv, xn, [prm["wq"], prm["wk"], prm["wv"]], [prm["bq"], prm["bk"], prm["bv"]])
    dxn = ops.gemm(dqkv, wqkv, trans_b=False)
    return dxn, g


class SelfAttentionFn(Function):
    """v2 SelfAttention.forward (src/v2/modules.py:123-162): fused QKV GEMM -> flash attention -> out-proj."""

    @staticmethod
    def forward(ctx, x, n_heads, wq, bq, wk, bk, wv, bv, wo, bo):
        adt = act_dtype()
        B, S, E = x.shape
        x2 = x.reshape(B * S, E)
        if x2.dtype != adt:
            x2 = ops.cast(x2, adt)
        wqkv, bqkv = packed([wq, wk, wv], adt), packed_vec([bq, bk, bv])
        wo_ = packed([wo], adt)
        scale = 1.0 / math.sqrt(E // n_heads)
        y, qkv, o, lse = _attn_fwd(x2.contiguous(), B, S, n_heads, wqkv, bqkv, wo_, bo.detach(), None, scale)
        ctx.save_for_backward(x2, qkv, o, lse, wq, wk, wv, wo)
        ctx.prm = dict(wq=wq, wk=wk, wv=wv, wo=wo, bq=bq, bk=bk, bv=bv, bo=bo)
        ctx.meta = (B, S, E, n_heads, scale, x.dtype)
        ctx.skip_pg = _SKIP_PARAM_GRADS
        return y.reshape(B, S, E)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2, qkv, o, lse, wq, wk, wv, wo = ctx.saved_tensors
        B, S, E, H, scale, xdtype = ctx.meta
        adt = x2.dtype
        dy2 = dy.reshape(B * S, E).contiguous()
        if dy2.dtype != adt:
            dy2 = ops.cast(dy2, adt)
        dxn, g = _attn_bwd(dy2, x2, qkv, o, lse, B, S, H, packed([wq, wk, wv], adt), packed([wo], adt), scale,
                           not ctx.skip_pg, ctx.prm)
        dx = dxn if dxn.dtype == xdtype else ops.cast(dxn, xdtype)
        dwq, dwk, dwv = _split_rows(g["wqkv"], (E, E, E))
        dbq, dbk, dbv = _split_rows(g["bqkv"], (E, E, E))
        return (dx.reshape(B, S, E), None, dwq, dbq, dwk, dbk, dwv, dbv, g["wo"], g["bo"])


# --------------------------------------------------------------------------------------------------
# v2 Encoder block (src/v2/modules.py:165-183): pre-LN attention + GELU MLP, both with skip connections
# --------------------------------------------------------------------------------------------------
class EncoderFn(Function):
    @staticmethod
    def forward(ctx, x, ln1_xn, ln1_mean, ln1_rstd, next_g, next_b, n_heads, n1w, n1b, wq, bq, wk, bk, wv, bv, wo, bo, n2w, n2b, w1, b1, w2, b2):
        """ln1_*: LayerNorm-1 of x already produced by the previous block's fc2 epilogue (or None).  next_g / next_b: affine
        parameters of the NEXT block's LayerNorm-1: then its (xn, mean, rstd) are returned as three extra, non-differentiable
        outputs (they are functions of y; the gradient reaches y through the next block's own backward)."""
        adt = act_dtype()
        B, S, E = x.shape
        x2 = x.reshape(B * S, E)
        if x2.dtype != adt:
            x2 = ops.cast(x2, adt)
        x2 = x2.contiguous()
        scale = 1.0 / math.sqrt(E // n_heads)
        eps = 1e-5
        if ln1_xn is not None:
            xn1, mean1, rstd1 = ln1_xn, ln1_mean, ln1_rstd
        else:
            xn1, mean1, rstd1 = ops.layernorm_fwd(x2, n1w.detach(), n1b.detach())
        x1, qkv, o, lse, (xn2, mean2, rstd2) = _attn_fwd(xn1, B, S, n_heads, packed([wq, wk, wv], adt), packed_vec([bq, bk, bv]),
                                                         packed([wo], adt), bo.detach(), x2, scale, ln=(n2w.detach(), n2b.detach(), eps))
        g, u = ops.gemm(xn2, packed([w1], adt), bias=b1.detach(), act=L.ACT_GELU, want_pre=True)   # fc1 + GELU
        nxt = (None, None, None)
        if next_g is not None:
            y, nxt = gemm_ln(g, packed([w2], adt), b2.detach(), x1, (next_g.detach(), next_b.detach(), eps))   # fc2 + skip + next LN1
        else:
            y = ops.gemm(g, packed([w2], adt), bias=b2.detach(), residual=x1)                       # fc2 + skip
        ctx.save_for_backward(x2, mean1, rstd1, xn1, qkv, o, lse, x1, mean2, rstd2, xn2, u, g,
                              n1w, wq, wk, wv, wo, n2w, w1, w2)
        ctx.prm = dict(wq=wq, wk=wk, wv=wv, wo=wo, bq=bq, bk=bk, bv=bv, bo=bo, n1b=n1b, n2b=n2b, b1=b1, b2=b2)
        ctx.meta = (B, S, E, n_heads, scale, x.dtype)
        ctx.skip_pg = _SKIP_PARAM_GRADS
        if nxt[0] is not None:
            ctx.mark_non_differentiable(*nxt)
            ctx.set_materialize_grads(False)      # no zero-filled "gradients" for the three statistics outputs
        return (y.reshape(B, S, E),) + tuple(nxt)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, *_unused):
        if dy is None:
            return (None,) * 23
        (x2, mean1, rstd1, xn1, qkv, o, lse, x1, mean2, rstd2, xn2, u, g, n1w, wq, wk, wv, wo, n2w, w1, w2) = ctx.saved_tensors
        B, S, E, H, scale, xdtype = ctx.meta
        adt = x2.dtype
        pg = not ctx.skip_pg
        dy2 = dy.reshape(B * S, E).contiguous()
        if dy2.dtype != adt:
            dy2 = ops.cast(dy2, adt)
        # ---- MLP half
        P = ctx.prm
        dw2 = db2 = dw1 = db1 = None
        if pg:
            dw2 = acc_wgrad(dy2, g, [w2])
        du = ops.gemm(dy2, packed([w2], adt), trans_b=False, act=L.ACT_MUL_DGELU, aux=u)       # dgrad fc2 x gelu'(u)
        if pg:
            dw1, db1 = acc_wgrad(du, xn2, [w1], [P["b1"]])
        dxn2 = ops.gemm(du, packed([w1], adt), trans_b=False)
        # LN2 backward also yields db2 = colsum(dy2) (its skip input) and dbo = colsum(dx1) (its output)
        dx1, dn2w, dn2b, db2, dbo = ln_bwd(dxn2, x1, mean2, rstd2, n2w, P["n2b"], dy2, pg, bias_of_dres=P["b2"], bias_of_dx=P["bo"])
        # ---- attention half
        dxn1, ga = _attn_bwd(dx1, xn1, qkv, o, lse, B, S, H, packed([wq, wk, wv], adt), packed([wo], adt), scale, pg, P, bo_done=True)
        ga["bo"] = dbo
        dx, dn1w, dn1b, _, _ = ln_bwd(dxn1, x2, mean1, rstd1, n1w, P["n1b"], dx1, pg)
        if dx.dtype != xdtype:
            dx = ops.cast(dx, xdtype)
        dx = dx.reshape(B, S, E)
        dwq, dwk, dwv = _split_rows(ga["wqkv"], (E, E, E))
        dbq, dbk, dbv = _split_rows(ga["bqkv"], (E, E, E))
        return (dx, None, None, None, None, None, None, dn1w, dn1b, dwq, dbq, dwk, dbk, dwv, dbv, ga["wo"], ga["bo"], dn2w, dn2b, dw1, db1, dw2, db2)


# --------------------------------------------------------------------------------------------------
# CLS-row gather (x[:, 0, :]) with scatter backward
# --------------------------------------------------------------------------------------------------
class ClsRowFn(Function):
    @staticmethod
    def forward(ctx, x):
        B, S, E = x.shape
        xc = x.contiguous()
        out = torch.empty(B, E, dtype=x.dtype, device=x.device)
        ops.copy_rows(xc, S * E, B, E, out, E)
        ctx.shape = (B, S, E)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        B, S, E = ctx.shape
        dx = torch.zeros(B, S, E, dtype=dy.dtype, device=dy.device)
        ops.copy_rows(dy.contiguous(), E, B, E, dx, S * E)
        return dx


# --------------------------------------------------------------------------------------------------
# loss head: nn.CrossEntropyLoss (mean) per group of rows, one launch forward, one multiply backward
# --------------------------------------------------------------------------------------------------
class SoftmaxCEFn(Function):
    """losses[g] = CrossEntropyLoss(logits[g*n:(g+1)*n], targets[...]) (src/v2/training.py:159,182-210)."""

    @staticmethod
    def forward(ctx, logits, targets, rows_per_group):
        losses, dlog = ops.softmax_ce(logits.contiguous(), targets, rows_per_group)
        ctx.save_for_backward(dlog)
        return losses

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (dlog,) = ctx.saved_tensors
        G = g.shape[0]
        return (dlog.view(G, -1) * g.reshape(G, 1)).view_as(dlog), None, None


def softmax_ce(logits, targets, rows_per_group=None):
    return SoftmaxCEFn.apply(logits, targets, rows_per_group or logits.shape[0])


class BCEFn(Function):
    """losses[g] = nn.BCELoss()(prob[g*n:(g+1)*n], target[...]) (src/v1/gan.py:16-20, 222-252), one launch forward,
    one multiply backward."""

    @staticmethod
    def forward(ctx, prob, target, rows_per_group):
        losses, dprob = ops.bce(prob.contiguous(), target.contiguous(), rows_per_group)
        ctx.save_for_backward(dprob)
        return losses

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (dprob,) = ctx.saved_tensors
        G = g.shape[0]
        return (dprob.view(G, -1) * g.reshape(G, 1)).view_as(dprob), None, None


def bce(prob, target, rows_per_group=None):
    return BCEFn.apply(prob, target, rows_per_group or prob.numel())
