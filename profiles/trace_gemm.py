"""Per-CTA timeline of one tcgen05 GEMM launch (vg_gemm_set_trace): where the fixed cost of the skinny C2 GEMMs goes.
Launches the kernel after a dependent predecessor (as in the step graph) and prints, relative to the predecessor's end / the
first CTA's entry, the median and max over CTAs of every milestone.
Needs the library built with tracing: make -C vit-gan_b200/csrc clean && make -C vit-gan_b200/csrc TRACE=1 (all zeros otherwise)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitgan_b200 as vb
L, bf = vb.lib, torch.bfloat16
names = ["entry", "prologue", "pdl_wait", "tma0_issued", "full0", "mma0_commit", "tfull0", "store0", "drained", "exit"]

def run(M, N, K=128, label="", kind="plain"):
    x, w, b = torch.randn(M, K, device="cuda").to(bf), torch.randn(N, K, device="cuda").to(bf), torch.randn(N, device="cuda")
    o = torch.empty(M, N, device="cuda", dtype=bf)
    aux = torch.randn(M, N, device="cuda").to(bf)
    wt = torch.randn(K, N, device="cuda").to(bf)
    if kind == "fc1":
        call = lambda: vb.ops.gemm(x, w, bias=b, act=L.ACT_GELU, want_pre=True, path=L.GEMM_TCGEN05)
    elif kind == "dgrad_gelu":
        call = lambda: vb.ops.gemm(x, wt, trans_b=False, act=L.ACT_MUL_DGELU, aux=aux, out=o, path=L.GEMM_TCGEN05)
    else:
        call = lambda: vb.ops.gemm(x, w, bias=b, out=o, path=L.GEMM_TCGEN05)
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
    tr = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    for rep in range(3):
        flush.zero_()
        call()           # predecessor (same kernel, PDL edge)
        tr.zero_()
        torch.cuda.synchronize()
        vb.lib.lib.vg_gemm_set_trace(tr.data_ptr())
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        call()
        e.record()
        torch.cuda.synchronize()
        vb.lib.lib.vg_gemm_set_trace(None)
    t = tr.view(148, 16).cpu()
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    print(f"--- {label} M={M} N={N} K={K}: {t.shape[0]} CTAs, tiles/CTA {t[:, 10].min().item()}..{t[:, 10].max().item()}, event time {s.elapsed_time(e) * 1e3:.1f} us (cold L2, eager launch)")
    for k, n in enumerate(names):
        if n == "tfull0":
            continue
        v = (t[:, k] - t0).float() / 1e3
        print(f"    {n:12s} median {v.median().item():7.2f} us   min {v.min().item():7.2f}   max {v.max().item():7.2f}")

run(33280, 384, label="qkv fwd")
run(33280, 128, label="out-proj")
run(66560, 384, label="qkv fwd (merged D pass)")
run(33280, 256, label="fc1 + GELU (+pre)", kind="fc1")
run(33280, 256, label="dgrad fc2 x GELU'(aux)", kind="dgrad_gelu")
