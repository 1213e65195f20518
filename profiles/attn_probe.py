"""Bring-up / timing probe for the attention kernels (profiles/ tooling, not a test): per shape, the relative error of the
forward (o, lse) and backward (dq, dk, dv) against the fp32 formula evaluated by torch on the same device, the kernel family
that ran (vg_attention_path) and, with --time, CUDA-event timings and algorithmic TFLOP/s (4 S^2 d fwd, 8 S^2 d bwd per head).

    python profiles/attn_probe.py [--what fwd|bwd|both] [--time] [--shapes B,H,S,d,mode ...]
"""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitgan_b200 as vb  # noqa: E402

DEFAULT = ["1,1,64,192,0", "1,1,128,192,0", "2,2,257,192,0", "3,4,64,96,0", "2,4,65,112,0", "2,4,65,112,1", "2,4,64,96,1",
           "1,2,200,96,0", "1,2,200,112,1", "40,4,257,192,0", "300,4,65,112,1"]


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


def reference(qkv, d_o, B, H, S, d, scale, mode, want_grad):
    hd = H * d
    x = qkv.float().requires_grad_(want_grad)
    q, k, v = [x[:, i * hd:(i + 1) * hd].reshape(B, S, H, d).permute(0, 2, 1, 3) for i in range(3)]
    if mode == 1:
        qq, kk = (q * q).sum(-1, keepdim=True), (k * k).sum(-1, keepdim=True)
        s = (qq + kk.transpose(-1, -2) - 2 * q @ k.transpose(-1, -2)).clamp_min(0).sqrt() * scale
    else:
        s = (q @ k.transpose(-1, -2)) * scale
    o = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * S, hd)
    lse = torch.logsumexp(s, -1).reshape(-1)
    g = None
    if want_grad:
        o.backward(d_o.float())
        g = x.grad
    return o.detach(), lse.detach(), g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="both")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--shapes", nargs="*", default=DEFAULT)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    res = []
    for spec in a.shapes:
        B, H, S, d, mode = [int(t) for t in spec.split(",")]
        hd = H * d
        g = torch.Generator("cpu").manual_seed(B + S + d + mode)
        qkv = (torch.randn(B * S, 3 * hd, generator=g) * (1.0 if mode else 0.7)).bfloat16().cuda()
        d_o = torch.randn(B * S, hd, generator=g).bfloat16().cuda()
        scale = 1.0 / math.sqrt(d if mode == 0 else hd)
        path = vb.lib.lib.vg_attention_path(1, mode, B, H, S, d)
        r = {"shape": spec, "path": path}
        big = B * H * S * S > 3e8
        try:
            o, lse = vb.ops.attention_fwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], B, H, S, d, scale, mode)
            torch.cuda.synchronize()
            if not big:
                oref, lref, gref = reference(qkv, d_o, B, H, S, d, scale, mode, a.what != "fwd")
                r["o"], r["lse"] = rel(o, oref), rel(lse, lref)
            if a.what != "fwd":
                dqkv = vb.ops.attention_bwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], o, d_o, lse, B, H, S, d, scale, mode)
                torch.cuda.synchronize()
                r["finite"] = bool(torch.isfinite(dqkv).all())
                if not big:
                    for i, n in enumerate(("dq", "dk", "dv")):
                        r[n] = rel(dqkv[:, i * hd:(i + 1) * hd], gref[:, i * hd:(i + 1) * hd])
            if a.time:
                flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
                for name, fn, fl in (("fwd", lambda: vb.ops.attention_fwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], B, H, S, d, scale, mode), 4),
                                     ("bwd", lambda: vb.ops.attention_bwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], o, d_o, lse, B, H, S, d, scale, mode), 8)):
                    if name == "bwd" and a.what == "fwd":
                        continue
                    ts = []
                    for it in range(8):
                        flush.zero_()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                        ts.append(e0.elapsed_time(e1) * 1e3)
                    us = sorted(ts[2:])[len(ts[2:]) // 2]
                    r[name + "_us"] = round(us, 1)
                    r[name + "_tflops"] = round(fl * B * H * S * S * d / us / 1e6, 1)
        except Exception as e:  # noqa: BLE001
            r["error"] = repr(e)[:300]
            res.append(r)
            print(json.dumps(r), flush=True)
            break
        res.append(r)
        print(json.dumps(r), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
