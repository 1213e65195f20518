"""Micro-benchmark + correctness check of the tcgen05 attention kernels at the C2 shape (B=512, H=4, S=65, d=32, bf16):
graph-batched launches over rotating >L2 buffers (bench.time_graph).  VG_ATTN_HP=0 selects the first-generation kernels."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitgan_b200 as vb
from bench import time_graph

bf = torch.bfloat16


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()


def check(B, H, S, d):
    g = torch.Generator().manual_seed(B + S + d)
    hd = H * d
    qkv = (torch.randn(B * S, 3 * hd, generator=g) * 0.8).bfloat16()
    d_o = torch.randn(B * S, hd, generator=g).bfloat16()
    scale = 1.0 / math.sqrt(d)
    ref_in = qkv.float().cuda().requires_grad_(True)
    q, k, v = [ref_in[:, i * hd:(i + 1) * hd].reshape(B, S, H, d).permute(0, 2, 1, 3) for i in range(3)]
    s = (q @ k.transpose(-1, -2)) * scale
    oref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * S, hd)
    lse_ref = torch.logsumexp(s, -1).reshape(-1)
    oref.backward(d_o.float().cuda())
    qc = qkv.cuda()
    o, lse = vb.ops.attention_fwd(qc[:, :hd], qc[:, hd:2 * hd], qc[:, 2 * hd:], B, H, S, d, scale, 0)
    dqkv = vb.ops.attention_bwd(qc[:, :hd], qc[:, hd:2 * hd], qc[:, 2 * hd:], o, d_o.cuda(), lse, B, H, S, d, scale, 0)
    torch.cuda.synchronize()
    errs = [rel(o, oref), rel(lse, lse_ref)] + [rel(dqkv[:, i * hd:(i + 1) * hd], ref_in.grad[:, i * hd:(i + 1) * hd]) for i in range(3)]
    ok = all(e < 2e-2 for e in errs) and bool(torch.isfinite(dqkv).all())
    print(f"check B={B} H={H} S={S} d={d}: o {errs[0]:.2e} lse {errs[1]:.2e} dq {errs[2]:.2e} dk {errs[3]:.2e} dv {errs[4]:.2e} {'OK' if ok else 'FAIL'}", flush=True)
    return ok


def timing(B=512, H=4, S=65, d=32):
    E, M, scale = H * d, B * S, d ** -0.5
    mk = lambda *shape: torch.randn(*shape, device="cuda").to(bf)

    def mk_f(i):
        qkv = mk(M, 3 * E)
        return lambda: vb.ops.attention_fwd(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], B, H, S, d, scale)

    def mk_b(i):
        qkv, d_o = mk(M, 3 * E), mk(M, E)
        o, lse = vb.ops.attention_fwd(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], B, H, S, d, scale)
        return lambda: vb.ops.attention_bwd(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], o, d_o, lse, B, H, S, d, scale)

    tf, tb = time_graph(mk_f, 10) * 1e3, time_graph(mk_b, 6) * 1e3
    bytes_f, bytes_b = 2.0 * 4 * M * E, 2.0 * 7 * M * E
    print(f"timing B={B} H={H} S={S} d={d} hp={os.environ.get('VG_ATTN_HP', '1')} ctas={os.environ.get('VG_ATTN_HP_CTAS', '2')}: "
          f"fwd {tf:.1f} us ({bytes_f / tf / 1e3:.0f} GB/s)  bwd {tb:.1f} us ({bytes_b / tb / 1e3:.0f} GB/s)", flush=True)


if __name__ == "__main__":
    if "--no-check" not in sys.argv:
        ok = True
        for shp in [(3, 4, 65, 32), (300, 4, 65, 32), (2, 8, 65, 32), (4, 2, 64, 64), (5, 4, 17, 32), (2, 4, 1, 32), (7, 2, 96, 32),
                    (3, 1, 80, 64), (2, 6, 33, 32), (3, 4, 128, 64)]:
            ok &= check(*shp)
        print("ALL OK" if ok else "SOME FAILED")
    timing()
