"""Fixed cost per dependent kernel launch inside a CUDA graph (tiny LN launches), and LN / colsum streaming cost vs rows."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitgan_b200 as vb
from bench import time_graph
bf = torch.bfloat16
E = 128
gam, bet = torch.ones(E, device="cuda"), torch.zeros(E, device="cuda")

def ln_f(M):
    def mk(i):
        x = torch.randn(M, E, device="cuda").to(bf)
        return lambda: vb.ops.layernorm_fwd(x, gam, bet)
    return mk

def ln_b(M):
    def mk(i):
        x, dy, dr = [torch.randn(M, E, device="cuda").to(bf) for _ in range(3)]
        _, mean, rstd = vb.ops.layernorm_fwd(x, gam, bet)
        cr, cx = torch.zeros(E, device="cuda"), torch.zeros(E, device="cuda")
        return lambda: vb.ops.layernorm_bwd(dy, x, mean, rstd, gam, dres=dr, dres_colsum=cr, dx_colsum=cx)
    return mk

def gemm(M, N):
    w, b = torch.randn(N, E, device="cuda").to(bf), torch.randn(N, device="cuda")
    def mk(i):
        x, o = torch.randn(M, E, device="cuda").to(bf), torch.empty(M, N, device="cuda", dtype=bf)
        return lambda: vb.ops.gemm(x, w, bias=b, out=o, path=vb.lib.GEMM_TCGEN05)
    return mk

print("pdl", os.environ.get("VG_PDL", "1"))
for M in (8, 33280, 66560, 133120):
    n = max(2, int(300e6 // (M * E * 2 * 4)) + 1) if M > 8 else 2
    print(f"M={M:7d}  ln_fwd {time_graph(ln_f(M), n) * 1e3:6.2f} us   ln_bwd {time_graph(ln_b(M), n) * 1e3:6.2f} us   "
          f"gemm N=128 {time_graph(gemm(M, 128), n) * 1e3:6.2f} us   gemm N=384 {time_graph(gemm(M, 384), n) * 1e3:6.2f} us", flush=True)
