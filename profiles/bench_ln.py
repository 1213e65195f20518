"""Micro-benchmark of layernorm_bwd (default: the C2 shape M=33280, E=128, bf16; `python profiles/bench_ln.py 65792 768` for the C4
micro-batch): graph-batched, rotating >L2 buffers; prints us per launch and the fraction of the measured HBM copy bandwidth."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitgan_b200 as vb
from bench import time_graph
M, E = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (33280, 128)      # e.g. 65792 768 for the C4 micro-batch
bf = torch.bfloat16
gam, bet = torch.ones(E, device="cuda"), torch.zeros(E, device="cuda")
def mk(i, ws):
    x, dy, dr = [torch.randn(M, E, device="cuda").to(bf) for _ in range(3)]
    _, mean, rstd = vb.ops.layernorm_fwd(x, gam, bet)
    cr, cx = torch.zeros(E, device="cuda"), torch.zeros(E, device="cuda")
    if ws:
        return lambda: vb.ops.layernorm_bwd(dy, x, mean, rstd, gam, dres=dr, dres_colsum=cr, dx_colsum=cx)
    return lambda: vb.ops.layernorm_bwd(dy, x, mean, rstd, gam, dres=dr)
nset = max(2, int(400e6 // (M * E * 2 * 4)) + 1)
it = 48 if M * E < 2e7 else 8
t_plain, t_cs = time_graph(lambda i: mk(i, False), nset, iters=it) * 1e3, time_graph(lambda i: mk(i, True), nset, iters=it) * 1e3
t_fwd = time_graph(lambda i: (lambda x=torch.randn(M, E, device="cuda").to(bf): vb.ops.layernorm_fwd(x, gam, bet)), nset, iters=it) * 1e3
bw = lambda us, n: n * M * E * 2 / us / 1e3     # GB/s for n tensor passes
print("M %d E %d  ln_bwd plain %.1f us (%.0f GB/s)   with fused colsums %.1f us (%.0f GB/s)   ln_fwd %.1f us (%.0f GB/s)" % (
    M, E, t_plain, bw(t_plain, 4), t_cs, bw(t_cs, 4), t_fwd, bw(t_fwd, 2)))
