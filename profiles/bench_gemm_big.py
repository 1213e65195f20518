"""tcgen05 GEMM at the compute-bound C4 shapes (E=768, M = 256 images x 257 tokens) and 8192^3; VG_TC_BN=128|256 forces the tile width."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitgan_b200 as vb
from bench import time_graph
bf, L = torch.bfloat16, vb.lib
mk = lambda *s: torch.randn(*s, device="cuda").to(bf)
M, E = 256 * 257, 768
peak = 1621.0
def run(name, M_, N_, K_, make, lib=None):
    t = time_graph(make, 2, iters=8) * 1e-3
    tf = 2.0 * M_ * N_ * K_ / t / 1e12
    if lib is not None:      # same-box comparison point (vendor library through torch), not part of the product
        tl = time_graph(lib, 2, iters=8) * 1e-3
        name = f"{name}  [torch/cuBLAS {tl * 1e6:.1f} us = {2.0 * M_ * N_ * K_ / tl / 1e12:.0f} TF]"
    print(f"bn={os.environ.get('VG_TC_BN', 'auto'):>4s} {name:34s} {t * 1e6:8.1f} us  {tf:7.1f} TFLOP/s  {100 * tf / peak:5.1f}% of burst bf16", flush=True)
wq, bq = mk(3 * E, E), torch.randn(3 * E, device="cuda")
w1, b1 = mk(2 * E, E), torch.randn(2 * E, device="cuda")
def f_qkv(i):
    x, o = mk(M, E), torch.empty(M, 3 * E, device="cuda", dtype=bf)
    return lambda: vb.ops.gemm(x, wq, bias=bq, out=o, path=L.GEMM_TCGEN05)
def f_fc1(i):
    x = mk(M, E)
    return lambda: vb.ops.gemm(x, w1, bias=b1, act=L.ACT_GELU, want_pre=True, path=L.GEMM_TCGEN05)
def f_dgrad(i):
    dy, o = mk(M, 3 * E), torch.empty(M, E, device="cuda", dtype=bf)
    return lambda: vb.ops.gemm(dy, wq, trans_b=False, out=o, path=L.GEMM_TCGEN05)
def f_wgrad(i):
    dy, x, o = mk(M, 3 * E), mk(M, E), torch.zeros(3 * E, E, device="cuda")
    return lambda: vb.ops.gemm(dy, x, trans_a=True, trans_b=False, accumulate=True, out=o, path=L.GEMM_TCGEN05)
w2, b2 = mk(E, 2 * E), torch.randn(E, device="cuda")
def f_fc2_res(i):      # fc2 forward + bias + residual (side tile prefetched one group ahead)
    g, x, o = mk(M, 2 * E), mk(M, E), torch.empty(M, E, device="cuda", dtype=bf)
    return lambda: vb.ops.gemm(g, w2, bias=b2, residual=x, out=o, path=L.GEMM_TCGEN05)
def f_fc2_dgrad(i):    # dgrad of fc2 fused with GELU'(u): dU = (dY W2) * gelu'(u), aux = pre-activation u
    dy, u, o = mk(M, E), mk(M, 2 * E), torch.empty(M, 2 * E, device="cuda", dtype=bf)
    return lambda: vb.ops.gemm(dy, w2, trans_b=False, act=L.ACT_MUL_DGELU, aux=u, out=o, path=L.GEMM_TCGEN05)
def f_sq(i):
    a, b, o = mk(8192, 8192), mk(8192, 8192), torch.empty(8192, 8192, device="cuda", dtype=bf)
    return lambda: vb.ops.gemm(a, b, out=o, path=L.GEMM_TCGEN05)
def l_qkv(i):
    x, o = mk(M, E), torch.empty(M, 3 * E, device="cuda", dtype=bf)
    bb = bq.to(bf)
    return lambda: torch.addmm(bb, x, wq.t(), out=o)
def l_dgrad(i):
    dy, o = mk(M, 3 * E), torch.empty(M, E, device="cuda", dtype=bf)
    return lambda: torch.mm(dy, wq, out=o)
def l_wgrad(i):
    dy, x, o = mk(M, 3 * E), mk(M, E), torch.empty(3 * E, E, device="cuda", dtype=bf)
    return lambda: torch.mm(dy.t(), x, out=o)
def l_sq(i):
    a, b, o = mk(8192, 8192), mk(8192, 8192), torch.empty(8192, 8192, device="cuda", dtype=bf)
    return lambda: torch.mm(a, b.t(), out=o)
run("fwd qkv [65792x768]x[768x2304]+b", M, 3 * E, E, f_qkv, l_qkv)
run("fwd fc1+gelu(+pre) [..x768]x[768x1536]", M, 2 * E, E, f_fc1)
run("dgrad qkv [65792x2304]x[2304x768]", M, E, 3 * E, f_dgrad, l_dgrad)
run("wgrad qkv [2304x65792]x[65792x768]", 3 * E, E, M, f_wgrad, l_wgrad)
run("fwd fc2+b+residual [..x1536]x[1536x768]", M, E, 2 * E, f_fc2_res)
run("dgrad fc2 * gelu'(u) [..x768]x[768x1536]", M, 2 * E, E, f_fc2_dgrad)
run("8192^3", 8192, 8192, 8192, f_sq, l_sq)
