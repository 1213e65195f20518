"""Launch the multi-tile tcgen05 attention kernels and the C4-shape side kernels ONCE so that `ncu --set full` can capture them:
   ncu --set full --clock-control none --import-source on -k regex:'attn_.*_mt|ln_bwd_wide|gemm_tc' -o gpurun_out/prof_mt \
       python profiles/run_kernels_mt.py --once
Shapes: C4 micro-batch (256 images, H=4, S=257, d=192, dot scores); C3 discriminator (128 images, H=4, S=65, d=108 padded to 112,
L2-distance scores) and generator (128 images, H=4, S=64, d=96, dot scores); fc1+GELU(+pre) and LayerNorm backward at E=768.
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitgan_b200 as vb  # noqa: E402

L, bf, dev = vb.lib, torch.bfloat16, "cuda"
mk = lambda *s: (torch.randn(*s, device=dev) * 0.7).to(bf)
REPS = 1 if "--once" in sys.argv else 2


def attn(B, H, S, d, mode):
    hd = H * d
    qkv, d_o = mk(B * S, 3 * hd), mk(B * S, hd)
    scale = 1.0 / math.sqrt(d if mode == 0 else hd)
    for _ in range(REPS):
        o, lse = vb.ops.attention_fwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], B, H, S, d, scale, mode)
        vb.ops.attention_bwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], o, d_o, lse, B, H, S, d, scale, mode)


attn(256, 4, 257, 192, 0)
attn(128, 4, 65, 112, 1)
attn(128, 4, 64, 96, 0)
M4, E4, m = 256 * 257, 768, 2
x4, w1 = mk(M4, E4), mk(m * E4, E4)
b1 = torch.randn(m * E4, device=dev)
gam = torch.ones(E4, device=dev)
for _ in range(REPS):
    vb.ops.gemm(x4, w1, bias=b1, act=L.ACT_GELU, want_pre=True, path=L.GEMM_TCGEN05)        # fc1 + GELU (+ pre-activation)
    y, mean, rstd = vb.ops.layernorm_fwd(x4, gam, torch.zeros_like(gam))
    vb.ops.layernorm_bwd(x4, x4, mean, rstd, gam, dres=x4)
torch.cuda.synchronize()
print("ok")
