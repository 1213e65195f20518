"""Timeline of CTA 0 of the multi-tile attention forward kernel (vg_attention_set_trace): where a tile's time goes.

    python profiles/trace_attn.py B,H,S,d,mode [n_events]
Event codes -- producer: 1 Q load issued, 100+kb K block, 110+kb V block issued; MMA: 10 q_full, 11 s_free, 20+kb K block landed
(S MMAs issued right after), 12 S committed, 13 o_free, 30+kb V block landed, 40+kb p_full (PV MMAs issued), 14 O committed;
compute leader: 50 item start, 51 norms done, 52 s_full, 53 pass 1 done, 54 staging free, 60+kb P chunk written, 55 o_full,
56 drain done (store issued).
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitgan_b200 as vb  # noqa: E402

B, H, S, d, mode = [int(t) for t in sys.argv[1].split(",")]
nev = int(sys.argv[2]) if len(sys.argv) > 2 else 120
what = sys.argv[3] if len(sys.argv) > 3 else "fwd"      # fwd | bwd (roles 3-5: dQ kernel, 6-8: dK/dV kernel)
hd = H * d
qkv = (torch.randn(B * S, 3 * hd) * 0.7).bfloat16().cuda()
scale = 1.0 / math.sqrt(d if mode == 0 else hd)
fwd = lambda: vb.ops.attention_fwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], B, H, S, d, scale, mode)
o, lse = fwd()
d_o = torch.randn(B * S, hd).bfloat16().cuda()
bwd = lambda: vb.ops.attention_bwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], o, d_o, lse, B, H, S, d, scale, mode)
run = fwd if what == "fwd" else bwd
for _ in range(3):
    run()
torch.cuda.synchronize()
tr = torch.zeros(9 * 1024, dtype=torch.int64, device="cuda")
vb.lib.lib.vg_attention_set_trace(tr.data_ptr())
run()
torch.cuda.synchronize()
vb.lib.lib.vg_attention_set_trace(None)
t = tr.cpu().view(9, 512, 2)
ev = []
for r in range(9):
    for i in range(512):
        if t[r, i, 1] == 0:
            break
        ev.append((int(t[r, i, 1]), r, int(t[r, i, 0])))
ev.sort()
t0 = ev[0][0]
names = ["prod", "mma ", "comp"] * 3
for ts, r, code in ev[:nev]:
    print(f"{(ts - t0) / 1000:9.3f} us  k{r // 3} {names[r]}  {'    ' * (r % 3)}{code}")
print("total events", len(ev), "span us", (ev[-1][0] - t0) / 1000)
