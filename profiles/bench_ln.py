"""Micro-benchmark of layernorm_bwd at the C2 shape (M=33280, E=128, bf16): graph-batched, rotating >L2 buffers."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitgan_b200 as vb
from bench import time_graph
M, E = 33280, 128
bf = torch.bfloat16
gam, bet = torch.ones(E, device="cuda"), torch.zeros(E, device="cuda")
def mk(i, ws):
    x, dy, dr = [torch.randn(M, E, device="cuda").to(bf) for _ in range(3)]
    _, mean, rstd = vb.ops.layernorm_fwd(x, gam, bet)
    cr, cx = torch.zeros(E, device="cuda"), torch.zeros(E, device="cuda")
    if ws:
        return lambda: vb.ops.layernorm_bwd(dy, x, mean, rstd, gam, dres=dr, dres_colsum=cr, dx_colsum=cx)
    return lambda: vb.ops.layernorm_bwd(dy, x, mean, rstd, gam, dres=dr)
print("ctas/sm", os.environ.get("VG_LN_BWD_CTAS_PER_SM", "2"), "ln_bwd plain %.1f us   with fused colsums %.1f us   ln_fwd %.1f us" % (
    time_graph(lambda i: mk(i, False), 10) * 1e3, time_graph(lambda i: mk(i, True), 10) * 1e3,
    time_graph(lambda i: (lambda x=torch.randn(M, E, device="cuda").to(bf): vb.ops.layernorm_fwd(x, gam, bet)), 20) * 1e3))
