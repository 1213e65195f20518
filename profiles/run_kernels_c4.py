"""Launch every hot kernel of the scaled config (BASELINE.json configs[3]: E = 768, H = 4, d = 192, 256 images x 257 tokens per
micro-batch) ONCE, for one `ncu --set full` pass:
   ncu --set full --clock-control none --import-source on -k regex:'gemm_tc|attn_|ln_|colsum|adam' -o gpurun_out/prof_c4 \
       python profiles/run_kernels_c4.py
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitgan_b200 as vb  # noqa: E402

L, bf, dev = vb.lib, torch.bfloat16, "cuda"
mk = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(bf)
B, H, S, d, E, m = 256, 4, 257, 192, 768, 2
M = B * S
x, x2, x3 = mk(M, E), mk(M, E), mk(M, E)
wqkv, w1, w2, wo = mk(3 * E, E), mk(m * E, E), mk(E, m * E), mk(E, E)
bq, b1, b2 = torch.randn(3 * E, device=dev), torch.randn(m * E, device=dev), torch.randn(E, device=dev)
gam, bet = torch.ones(E, device=dev), torch.zeros(E, device=dev)
# ---- forward of one encoder block
xn, mean, rstd = vb.ops.layernorm_fwd(x, gam, bet)                                               # ln_fwd
qkv = vb.ops.gemm(xn, wqkv, bias=bq, path=L.GEMM_TCGEN05)                                        # gemm_tc2<1,0>  QKV
o, lse = vb.ops.attention_fwd(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], B, H, S, d, d ** -0.5)  # attn_fwd_mt
x1 = vb.ops.gemm(o, wo, bias=b2, residual=x, path=L.GEMM_TCGEN05)                                # gemm_tc2<1,0>  out-proj + residual
g, u = vb.ops.gemm(xn, w1, bias=b1, act=L.ACT_GELU, want_pre=True, path=L.GEMM_TCGEN05)          # gemm_tc2<1,1>  fc1 + GELU (+pre)
y = vb.ops.gemm(g, w2, bias=b2, residual=x1, path=L.GEMM_TCGEN05)                                # gemm_tc2<1,0>  fc2 + residual
# ---- backward
dy = mk(M, E)
dg = vb.ops.gemm(dy, w2, trans_b=False, act=L.ACT_MUL_DGELU, aux=u, path=L.GEMM_TCGEN05)         # gemm_tc2<1,5>  fc2 dgrad x GELU'
dw2 = torch.zeros(E, m * E, device=dev)
vb.ops.gemm(dy, g, trans_a=True, trans_b=False, accumulate=True, out=dw2, path=L.GEMM_TCGEN05)   # gemm_tc2<2,0>  wgrad split-K
db1 = vb.ops.colsum(dg)                                                                          # colsum_vec (bias gradient of fc1)
dxn = vb.ops.gemm(dg, w1, trans_b=False, path=L.GEMM_TCGEN05)                                    # gemm_tc2<1,0>  fc1 dgrad
cr, cx = torch.zeros(E, device=dev), torch.zeros(E, device=dev)
vb.ops.layernorm_bwd(dxn, x1, mean, rstd, gam, dres=dy, dres_colsum=cr, dx_colsum=cx)            # ln_bwd_wide<6,true>
vb.ops.layernorm_bwd(dxn, x, mean, rstd, gam, dres=dy)                                           # ln_bwd_wide<6,false>
vb.ops.attention_bwd(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], o, dy, lse, B, H, S, d, d ** -0.5)   # attn_bwd_dq_mt, attn_bwd_dkv_mt
torch.cuda.synchronize()
print("ok")
