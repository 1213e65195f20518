#!/usr/bin/env python
"""bench.py -- GAN training images/s for one full G+D step (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W                       (this repo's CUDA path)
  python bench.py --impl reference --gpus N --steps K --warmup W      (reference algorithm on the host CPU cores)

Default workload = BASELINE.json configs[3] ("c4"), the only config BASELINE.json quotes at 1/2/4/8 GPUs: scaled ViT-GAN 128x128,
patch 8, dim 768, depth 12 for G and D, GLOBAL batch 2048 strong-scaled over the N GPUs (2048/N images per GPU per step, run as
exact-gradient micro-batches of 256), bf16.  The same JSON line carries `secondary` records of c2 (configs[1]: v2 defaults 32x32,
512 per GPU) and c3 (configs[2]: v1 SLN generator / L2-attention spectral discriminator at 64x64, 128 per GPU).
Other workloads (--workload): c1 (v2 defaults B=64 fp32), c5 (generator-only sampling, B=4096).

One JSON line on stdout (rank 0).  `value` = whole-job images/s with inputs resident in HBM; `e2e` = the same
step driven from pinned HOST buffers (H2D of real+noise and D2H of the three losses inside the timed region);
`roofline` = the dominant kernel timed alone with CUDA events; `cpu_baseline` = the oracle on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--no-secondary", action="store_true", help="c4 only: skip the secondary c2 / c3 records")
    ap.add_argument("--no-dropin", action="store_true", help="skip the un-graphed reference-call-order torch-optimizer measurement")
    ap.add_argument("--precision", default=None, choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch override")
    ap.add_argument("--micro", type=int, default=None, help="micro-batches per step (exact gradient accumulation); default: auto")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of a whole-step CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the per-kernel timing section (used for ncu launch lists)")
    ap.add_argument("--torch-optim", action="store_true", help="torch.optim.AdamW instead of the fused flat Adam kernel")
    ap.add_argument("--no-pg-stream", action="store_true", help="keep the weight-gradient GEMMs on the main stream")
    ap.add_argument("--no-keep-g", action="store_true", help="micro-batched step: recompute every generator forward in the generator update")
    ap.add_argument("--separate-d-passes", action="store_true",
                    help="run D(real) and D(fake) of the discriminator update as two passes (reference call order) instead of one concatenated pass")
    ap.add_argument("--keep-unused-d-grads", action="store_true",
                    help="also compute D's parameter gradients in the G pass (the reference computes, then discards them)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p.get("bf16_tflops_sustained", p["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------------ workloads
def workload_spec(name, n_gpus, batch_override):
    """-> dict(kind, per_gpu_batch, global_batch, scaling, precision, cfg kwargs, description)."""
    if name in ("c1", "c2", "c5"):
        pb = {"c1": 64, "c2": 512, "c5": 4096}[name]
        spec = dict(kind="v2" if name != "c5" else "v2_sample", over={}, scaling="weak", precision="fp32" if name == "c1" else "bf16",
                    desc={"c1": "main-v2 default ViT-GAN 32x32, B=64/GPU, fp32 parity path",
                          "c2": "main-v2 default ViT-GAN 32x32 RGB (E128 L6 H4 P4 S65), B=512/GPU, bf16",
                          "c5": "v2 generator-only sampling (eval), B=4096/GPU, bf16"}[name])
    elif name == "c4":
        pb = 2048 // n_gpus
        spec = dict(kind="v2", over=dict(image_size=128, patch_size=8, embeddings_dimension=768, transformer_blocks_count=12),
                    scaling="strong", precision="bf16", desc="scaled ViT-GAN 128x128 P8 E768 L12 (H4 m2), global batch 2048, bf16")
    else:
        pb = 128
        spec = dict(kind="v1", over=dict(image_size=64), scaling="weak", precision="bf16",
                    desc="main-v1 ViTGAN variant (SLN generator, L2-attention spectral discriminator) 64x64, B=128/GPU, bf16")
    if batch_override:
        pb = batch_override
    spec.update(per_gpu_batch=pb, global_batch=pb * n_gpus, name=name)
    return spec


def step_flops_per_image(spec, skip_unused):
    """Algorithmic FLOPs per image of one G+D step: 9 F_D + 3 F_G, or 8 F_D + 3 F_G when the
    discarded D weight gradients of the third pass are not computed (their FLOPs leave the numerator too)."""
    from oracle import v1 as o1, v2 as o2
    if spec["kind"].startswith("v2"):
        cfg = o2.V2Config(**spec["over"])
        fd, fg = o2.flops_per_image_fwd(cfg, False), o2.flops_per_image_fwd(cfg, True)
    else:
        cfg = o1.V1Config(**spec["over"])
        fd, fg = o1.flops_per_image_fwd(cfg, False), o1.flops_per_image_fwd(cfg, True)
    if spec["kind"] == "v2_sample":
        return fg
    return (8 if skip_unused else 9) * fd + 3 * fg


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  In-process NVML (nvidia_ml_py) polled every ~2 ms -- the
    timed region of the default run is ~140 ms, too short for `nvidia-smi -lms`, whose first sample arrives after ~100 ms;
    nvidia-smi is the fallback when NVML cannot be loaded."""
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.rows, self.proc, self.nvml, self.stop_flag = [], None, None, False
        self.sm, self.mx, self.reasons = [], None, set()
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nvml, self.h = pynvml, h
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self.nvml, self.h
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(reasons_fn(h))
                for name, bit in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1.0)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx, "samples": len(self.sm),
                    "reasons": sorted(self.reasons), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------ reference arm / CPU baseline
# fixed per-step CPU sample batch per workload (no calibration: the same number of images in every run, so BENCH and SCALE agree);
# c1 / c2: 64 = the batch of BASELINE configs[0], the reference's own CPU-runnable case
CPU_SAMPLE_BATCH = {"c1": 64, "c2": 64, "c3": 16, "c4": 2, "c5": 64}


def cpu_reference_rate(spec, steps, warmup, batch=None):
    """Time the reference's CPU implementation of one G+D step on a bounded sample of the workload: same model / config, fixed
    small per-step batch, all host threads.  kind 'reference' = the reference's OWN modules and loop body (imported from
    /root/reference, or from the staged git-ignored copy oracle/_ref on the GPU box; external shims only), driven by
    torch Adam(W); kind 'port' = the oracle restatement (bit-exact to those modules, tests/test_oracle_vs_reference.py) when the
    reference sources are not staged."""
    from oracle import harness, refimport, v1 as o1, v2 as o2
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b = batch or CPU_SAMPLE_BATCH[spec["name"]]
    kind = "reference" if refimport.available() else "port"
    sample_kind = spec["kind"]
    if sample_kind.startswith("v2"):
        I = spec["over"].get("image_size", 32)
        cfg = o2.V2Config(**spec["over"], batch_size=3 * I * I)
        data = harness.synthetic_batches_v2(cfg, b, steps + warmup)
        if kind == "reference":
            gan, c = refimport.build_v2(seed=0, **spec["over"])
            go = torch.optim.AdamW(gan.generator.parameters(), lr=c.generator_learning_rate, weight_decay=1e-3)
            do = torch.optim.AdamW(gan.discriminator.parameters(), lr=c.discriminator_learning_rate, weight_decay=1e-3)
            gen, disc = gan.generator, gan.discriminator
            run = (lambda r, n: gen(n)) if sample_kind == "v2_sample" else (lambda r, n: harness.gan_step(gen, disc, go, do, r, n, "ce"))
            if sample_kind == "v2_sample":
                gan.eval()
        else:
            orc = harness.OracleV2(cfg, seed=0)
            run = (lambda r, n: orc.generator(n)) if sample_kind == "v2_sample" else (lambda r, n: orc.step(r, n))
    else:
        cfg = o1.V1Config(**spec["over"])
        data = harness.synthetic_batches_v1(cfg, b, steps + warmup)
        if kind == "reference":
            gen, disc = refimport.build_v1(image_size=spec["over"]["image_size"], seed=0)
            go = torch.optim.Adam(gen.parameters(), lr=2e-4, betas=(0.5, 0.999))
            do = torch.optim.Adam(disc.parameters(), lr=2e-4, betas=(0.5, 0.999))
            run = lambda r, n: harness.gan_step(gen, disc, go, do, r, n, "bce")
        else:
            orc = harness.OracleV1(cfg, seed=0)
            run = lambda r, n: orc.step(r, n)
    with torch.no_grad() if sample_kind == "v2_sample" else torch.enable_grad():
        for r, n in data[:warmup]:
            run(r, n)
        t0 = time.perf_counter()
        for r, n in data[warmup:]:
            run(r, n)
        dt = time.perf_counter() - t0
    return dict(value=b * steps / dt, batch=b, steps=steps, warmup=warmup, ms_per_step=1e3 * dt / steps, cores=torch.get_num_threads(), kind=kind)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec = workload_spec(args.workload, args.gpus, args.batch)
    # c4 on the CPU costs ~380 GFLOP per image: bound the number of steps so that the arm ends within a few minutes
    steps = args.steps if spec["name"] != "c4" else min(args.steps, 4)
    warmup = max(1, min(args.warmup, 2 if spec["name"] == "c4" else 5))
    res = cpu_reference_rate(spec, steps=steps, warmup=warmup)
    unit = "samples/s" if spec["kind"] == "v2_sample" else "img/s"
    what = ("the reference's own modules and loop body (src/, external shims only) under torch Adam(W)" if res["kind"] == "reference"
            else "oracle CPU port (bit-exact to the reference modules)")
    sample = (f"{what}, same model and config, fixed per-step batch {res['batch']} (the GPU arm: {spec['per_gpu_batch']} per GPU), "
              f"{res['warmup']} warm-up + {res['steps']} timed steps, {res['cores']} host threads")
    line = {
        "impl": "reference", "metric": "gan_train_images_per_sec" if unit == "img/s" else "generator_samples_per_sec",
        "value": res["value"], "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": spec["scaling"], "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{spec['name']}: {spec['desc']}", "per_gpu_batch": spec["per_gpu_batch"], "global_batch": spec["global_batch"],
                   "cpu_sample_batch": res["batch"], "cpu_steps_timed": res["steps"], "dropout": "p = 0 (parity protocol, SURVEY Q11)"},
        "cpu_baseline": {"value": res["value"], "unit": unit, "cores": res["cores"], "kind": res["kind"], "sample": sample},
        "e2e": {"value": res["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ kernel roofline (dominant kernel, timed alone)
def time_graph(make_call, n_sets, iters=48):
    """Average device time per launch: `iters` launches round-robin over `n_sets` disjoint buffer sets (footprint > L2,
    so every launch streams from HBM) captured into ONE CUDA graph -> no host launch overhead inside the event pair."""
    calls = [make_call(i) for i in range(n_sets)]
    for c in calls:
        c()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for c in calls:
            c()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            calls[i % n_sets]()
    g.replay()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record()
        e.synchronize()
        best = min(best, s.elapsed_time(e) / iters)
    return best


def kernel_rooflines(vb, spec, pk):
    """Per-launch algorithmic FLOPs/bytes over the CUDA-event time of each hot kernel, at this workload's shapes (c4: the shapes
    of one 256-image micro-batch, i.e. what every launch of the step sees at any N)."""
    from oracle import v1 as o1, v2 as o2
    dev, bf, L = "cuda", torch.bfloat16, vb.lib
    mk = lambda *shape: torch.randn(*shape, device=dev).to(bf)
    out = []

    def entry(name, make_call, set_bytes, flops, bytes_, iters=48):
        n_sets = max(2, int(300e6 // max(set_bytes, 1)) + 1)
        t = time_graph(make_call, n_sets, iters=iters) * 1e-3
        tf, gb = flops / t / 1e12, bytes_ / t / 1e9
        bound = "tensor" if flops / bytes_ > pk["tf_burst"] * 1e3 / pk["hbm"] else "hbm"
        out.append({"kernel": name, "bound": bound, "us": t * 1e6, "tflops": tf, "gbs": gb, "frac_tensor": tf / pk["tf_burst"],
                    "frac_hbm": gb / pk["hbm"], "alg_flops": flops, "alg_bytes": bytes_, "buffer_sets": n_sets})

    def attn_entries(tag, B, H, S, d, mode, scale, iters):
        hd, M = H * d, B * S
        path = {0: "CUDA-core", 1: "tcgen05 single-tile", 2: "tcgen05 multi-tile"}[L.lib.vg_attention_path(1, mode, B, H, S, d)]

        def mk_f(i):
            qkv = mk(M, 3 * hd)
            return lambda: vb.ops.attention_fwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], B, H, S, d, scale, mode)
        entry(f"attention fwd {tag} [{path}]", mk_f, 2 * 4 * M * hd, 4.0 * B * H * S * S * d, 2.0 * 4 * M * hd, iters)

        def mk_b(i):
            qkv, d_o = mk(M, 3 * hd), mk(M, hd)
            o, lse = vb.ops.attention_fwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], B, H, S, d, scale, mode)
            return lambda: vb.ops.attention_bwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], o, d_o, lse, B, H, S, d, scale, mode)
        entry(f"attention bwd {tag} [{path}]", mk_b, 2 * 8 * M * hd, 8.0 * B * H * S * S * d, 2.0 * 8 * M * hd, iters)

    if spec["kind"] == "v1":          # c3: the two attention shapes of the v1 GAN (the rest of its step is the same GEMM / LN kernels)
        c1 = o1.V1Config(**spec["over"])
        B = spec["per_gpu_batch"]
        attn_entries(f"v1 D L2-distance B{B} H4 S{c1.number_of_tokens + 1} d108->112", B, 4, c1.number_of_tokens + 1, 112, 1, 1.0 / (432 ** 0.5), 24)
        attn_entries(f"v1 G dot B{B} H4 S{c1.image_size} d96", B, 4, c1.image_size, 96, 0, 1.0 / (384 ** 0.5), 24)
        torch.cuda.empty_cache()
        return None, out
    if not spec["kind"].startswith("v2"):
        return None, []
    cfg = o2.V2Config(**spec["over"])
    big = spec["name"] == "c4"
    B = min(spec["per_gpu_batch"], 256) if big else spec["per_gpu_batch"]
    S, E, H, m = cfg.seq_len, cfg.embeddings_dimension, cfg.attention_heads_count, cfg.mlp_ratio
    M, d = B * S, cfg.embeddings_dimension // cfg.attention_heads_count
    it = 8 if big else 48
    wqkv, w1 = mk(3 * E, E), mk(m * E, E)
    bq, b1 = torch.randn(3 * E, device=dev), torch.randn(m * E, device=dev)
    gam, bet = torch.ones(E, device=dev), torch.zeros(E, device=dev)
    scale = d ** -0.5

    def mk_qkv(i):
        x, o = mk(M, E), torch.empty(M, 3 * E, device=dev, dtype=bf)
        return lambda: vb.ops.gemm(x, wqkv, bias=bq, out=o, path=L.GEMM_TCGEN05)
    entry("gemm_tc fwd qkv [M,E]x[E,3E]+bias", mk_qkv, 2 * (M * E + M * 3 * E), 2.0 * M * 3 * E * E, 2.0 * (M * E + 3 * E * E + M * 3 * E), it)

    def mk_fc1(i):
        x = mk(M, E)
        return lambda: vb.ops.gemm(x, w1, bias=b1, act=L.ACT_GELU, want_pre=True, path=L.GEMM_TCGEN05)
    entry("gemm_tc fwd fc1+gelu(+pre) [M,E]x[E,mE]", mk_fc1, 2 * (M * E + 2 * M * m * E), 2.0 * M * m * E * E, 2.0 * (M * E + m * E * E + 2 * M * m * E), it)

    def mk_dgrad(i):
        dy, o = mk(M, 3 * E), torch.empty(M, E, device=dev, dtype=bf)
        return lambda: vb.ops.gemm(dy, wqkv, trans_b=False, out=o, path=L.GEMM_TCGEN05)
    entry("gemm_tc dgrad qkv [M,3E]x[3E,E]", mk_dgrad, 2 * (M * 3 * E + M * E), 2.0 * M * 3 * E * E, 2.0 * (M * 3 * E + 3 * E * E + M * E), it)

    def mk_wgrad(i):
        dy, x, o = mk(M, 3 * E), mk(M, E), torch.zeros(3 * E, E, device=dev)
        return lambda: vb.ops.gemm(dy, x, trans_a=True, trans_b=False, accumulate=True, out=o, path=L.GEMM_TCGEN05)
    entry("gemm_tc wgrad qkv [3E,M]x[M,E] split-K", mk_wgrad, 2 * (M * 3 * E + M * E), 2.0 * M * 3 * E * E, 2.0 * (M * 3 * E + M * E) + 4.0 * 3 * E * E, it)

    if big:      # the other GEMM epilogues of the block at the compute-bound shapes (c2 fuses LayerNorm into these; see profiles/r02_*)
        w2, b2 = mk(E, m * E), torch.randn(E, device=dev)

        def mk_fc2(i):
            gact, x, o = mk(M, m * E), mk(M, E), torch.empty(M, E, device=dev, dtype=bf)
            return lambda: vb.ops.gemm(gact, w2, bias=b2, residual=x, out=o, path=L.GEMM_TCGEN05)
        entry("gemm_tc fwd fc2+bias+residual [M,mE]x[mE,E]", mk_fc2, 2 * (M * m * E + 2 * M * E), 2.0 * M * m * E * E,
              2.0 * (M * m * E + m * E * E + 2 * M * E), it)

        def mk_dgelu(i):
            dy, u, o = mk(M, E), mk(M, m * E), torch.empty(M, m * E, device=dev, dtype=bf)
            return lambda: vb.ops.gemm(dy, w2, trans_b=False, act=L.ACT_MUL_DGELU, aux=u, out=o, path=L.GEMM_TCGEN05)
        entry("gemm_tc dgrad fc2 x gelu'(u) [M,E]x[E,mE]", mk_dgelu, 2 * (M * E + 2 * M * m * E), 2.0 * M * m * E * E,
              2.0 * (M * E + m * E * E + 2 * M * m * E), it)

    attn_entries(f"B{B} H{H} S{S} d{d}", B, H, S, d, 0, scale, it)

    def mk_ln(i):
        x = mk(M, E)
        return lambda: vb.ops.layernorm_fwd(x, gam, bet)
    entry("layernorm fwd", mk_ln, 2 * 2 * M * E, 8.0 * M * E, 2.0 * 2 * M * E, it)

    def mk_lnb(i):
        x, dy, dr = mk(M, E), mk(M, E), mk(M, E)
        _, mean, rstd = vb.ops.layernorm_fwd(x, gam, bet)
        return lambda: vb.ops.layernorm_bwd(dy, x, mean, rstd, gam, dres=dr)
    entry("layernorm bwd (+residual grad)", mk_lnb, 2 * 4 * M * E, 12.0 * M * E, 2.0 * 4 * M * E, it)

    def mk_lnb_cs(i):
        x, dy, dr = mk(M, E), mk(M, E), mk(M, E)
        _, mean, rstd = vb.ops.layernorm_fwd(x, gam, bet)
        cr, cx = torch.zeros(E, device=dev), torch.zeros(E, device=dev)
        return lambda: vb.ops.layernorm_bwd(dy, x, mean, rstd, gam, dres=dr, dres_colsum=cr, dx_colsum=cx)
    entry("layernorm bwd (+residual grad, + the two neighbouring bias gradients)", mk_lnb_cs, 2 * 4 * M * E, 14.0 * M * E, 2.0 * 4 * M * E, it)
    torch.cuda.empty_cache()

    dom = out[0]                      # the top bucket of the step's launch list: profiles/r06_launches_c4_one_step.txt (c4), r02_* (c2)
    key = "frac_tensor" if dom["bound"] == "tensor" else "frac_hbm"
    traffic = None
    try:       # DRAM bytes per launch of the same kernel at the same shape from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(f"{dom['kernel']} M={M} E={E}")
        if t:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    except (OSError, ValueError, KeyError):
        traffic = None
    roof = {"kernel": dom["kernel"] + f" M={M} E={E}", "bound": dom["bound"],
            "achieved": dom["tflops"] if dom["bound"] == "tensor" else dom["gbs"],
            "peak": pk["tf_burst"] if dom["bound"] == "tensor" else pk["hbm"],
            "unit": "TFLOP/s" if dom["bound"] == "tensor" else "GB/s", "frac": dom[key], "traffic": traffic,
            "traffic_unit": "bytes per launch (dram read + write, ncu --set full; part of the output still sits in L2 when the kernel ends)",
            "peak_source": pk["src"], "us_per_launch": dom["us"],
            "note": "timed alone inside one CUDA graph over rotating >L2 buffer sets (burst peak); traffic from profiles/ ncu --set full"}
    return roof, out


# ------------------------------------------------------------------------------------------------ our arm
def measure(args, name, steps, warmup, full, rank, world, local):
    """One workload on the already-initialised process group: build the GAN, capture the step, time it (device-resident and end to
    end), and on rank 0 return the JSON record.  `full`: also the per-kernel rooflines, the CPU baseline and the drop-in number."""
    import gc
    import torch.distributed as dist
    import vitgan_b200 as vb
    from oracle import harness, v1 as o1, v2 as o2      # synthetic data generator + FLOP formulas + cpu_baseline leg only

    spec = workload_spec(name, world, args.batch if name == args.workload else None)
    prec = (args.precision if name == args.workload else None) or spec["precision"]
    vb.set_precision(prec)
    vb.set_dropout_policy("off")       # parity protocol: dropout p = 0 on both arms; reported in config
    pk = peaks()
    B = spec["per_gpu_batch"]
    skip_unused = not args.keep_unused_d_grads
    merge_d = not args.separate_d_passes
    torch.manual_seed(0)
    if spec["kind"].startswith("v2"):
        I = spec["over"].get("image_size", 32)
        cfg = vb.v2.Config(**spec["over"], batch_size=3 * I * I)
        gan = vb.v2.ViTGAN(cfg).cuda()
        gen, disc, loss_kind = gan.generator, gan.discriminator, "ce"
        ocfg = o2.V2Config(**spec["over"], batch_size=3 * I * I)
        mk = lambda n, seed: harness.synthetic_batches_v2(ocfg, B, n, seed=seed)
        opt = lambda net: vb.train.FusedAdam(net, 5e-4, weight_decay=1e-3, decoupled=True)
        topt = lambda ps: torch.optim.AdamW(ps, lr=5e-4, weight_decay=1e-3, capturable=True)
    else:
        I = spec["over"]["image_size"]
        gen = vb.v1.Generator(vb.v1.V1Config(**spec["over"])).cuda()
        disc = vb.v1.Discriminator(vb.v1.V1Config(**spec["over"])).cuda()
        loss_kind = "bce"
        ocfg = o1.V1Config(**spec["over"])
        mk = lambda n, seed: harness.synthetic_batches_v1(ocfg, B, n, seed=seed)
        opt = lambda net: vb.train.FusedAdam(net, 2e-4, betas=(0.5, 0.999))
        topt = lambda ps: torch.optim.Adam(ps, lr=2e-4, betas=(0.5, 0.999), capturable=True)

    n_data = 2 if spec["name"] == "c4" else 4           # c4 at N=1: 2 x 0.8 GB of pinned host data
    host = [(r.pin_memory(), n.pin_memory()) for r, n in mk(n_data, 1234 + rank)]     # each rank its own shard of the global batch
    devb = [(r.cuda(), n.cuda()) for r, n in host]
    unit = "img/s"
    gs = d_b = g_b = None
    n_micro, keep_g, keep_info = 1, 0, {}

    if spec["kind"] == "v2_sample":
        unit = "samples/s"
        gen.eval()

        def one(real, noise):
            # batched generator forward + fused de-normalise -> uint8 (the reference's sampling tail, generation.py:47-56)
            return (vb.v2.sample_uint8(gen, noise).sum(dtype=torch.int64).float(),)
        step_fn = one
        graph_used = False
    else:
        frozen = (lambda n, p: n.endswith((".q.weight", ".k.weight", ".v.weight"))) if spec["kind"] == "v1" else None
        if args.torch_optim:
            gopt = topt(list(gen.parameters()))
            dopt = topt([p for n, p in disc.named_parameters() if not (frozen and frozen(n, p))])
            assert world == 1, "--torch-optim is a single-GPU diagnostic"
        else:
            gnet, dnet = vb.train.FlatNet(gen), vb.train.FlatNet(disc, exclude=frozen)
            gopt, dopt = opt(gnet), opt(dnet)
            if world > 1:
                # one all-reduce per network when it is small (< 32 MB of fp32 gradients: latency-bound, and every bucket boundary
                # joins the parameter-gradient stream), two buckets (the first overlapping the rest of the backward) otherwise
                nb = lambda net: 1 if net.numel * 4 < (32 << 20) else 2
                d_b = vb.train.GradBuckets(dnet, n_buckets=nb(dnet), average_in_place=False)
                g_b = vb.train.GradBuckets(gnet, n_buckets=nb(gnet), average_in_place=False)
                gopt.grad_scale = dopt.grad_scale = 1.0 / world

        # micro-batching: c4's 2048/N images per GPU run as exact-gradient micro-batches of 256 (saved activations of one
        # micro-batch: ~50 GB of the 180 GB); every other workload is one batch
        n_micro = (args.micro if name == args.workload else None) or (max(1, B // 256) if spec["name"] == "c4" else 1)

        vb.functional.set_param_grad_stream(not args.no_pg_stream)     # eager steps use the same two-stream schedule the graph captures
        # generator graphs kept across the discriminator update (saves the second generator forward of those micro-batches):
        # as many as the measured free HBM holds, the same number on every rank
        keep_g, keep_info = 0, {}
        if n_micro > 1 and not args.no_keep_g:
            mbs = B // n_micro
            keep_g = vb.train.plan_keep_g_graphs(gen, disc, devb[0][0][:mbs], devb[0][1][:mbs], n_micro, info=keep_info)
            if world > 1:
                kt = torch.tensor([keep_g], device="cuda")
                dist.all_reduce(kt, op=dist.ReduceOp.MIN)
                keep_g = int(kt.item())

        def eager(real, noise):
            return vb.train.gan_step_microbatched(gen, disc, gopt, dopt, real, noise, loss_kind, n_micro=n_micro, d_buckets=d_b,
                                                  g_buckets=g_b, skip_unused_d_grads=skip_unused, merge_d_passes=merge_d, keep_g_graphs=keep_g)
        # launches per step, counted on one eager step (the graph replays exactly these)
        try:
            eager(*devb[0])
            torch.cuda.synchronize()
        except torch.cuda.OutOfMemoryError:      # the plan was too optimistic for this box: recompute every generator forward instead
            if keep_g == 0:
                raise
            if rank == 0:
                print(f"[bench] out of memory with {keep_g} kept generator graphs; falling back to 0", file=sys.stderr)
            keep_g = 0
            for opt_ in (gopt, dopt):
                opt_.zero_grad(set_to_none=False) if isinstance(opt_, vb.train.FusedAdam) else opt_.zero_grad(set_to_none=True)
            vb.functional.join_param_grad_stream()
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
            eager(*devb[0])
            torch.cuda.synchronize()
        vb.ops.launch_count = 0
        eager(*devb[1])
        torch.cuda.synchronize()
        launches_per_step = vb.ops.launch_count
        step_fn, graph_used = eager, False
        torch.cuda.empty_cache()
        if not args.no_graph:
            try:
                gs = vb.train.GraphedStep(gen, disc, gopt, dopt, devb[0][0], devb[0][1], loss_kind, warmup=2, d_buckets=d_b,
                                          g_buckets=g_b, skip_unused_d_grads=skip_unused, n_micro=n_micro, merge_d_passes=merge_d,
                                          param_grad_stream=not args.no_pg_stream, **({"keep_g_graphs": keep_g} if n_micro > 1 else {}))
                step_fn, graph_used = gs, True
            except Exception as ex:      # fall back to eager launches of the same kernels (never to another implementation)
                if rank == 0:
                    print(f"[bench] CUDA-graph capture failed ({type(ex).__name__}: {ex}); running eager", file=sys.stderr)
                torch.cuda.synchronize()
                vb.set_operand_cache(True)
    if spec["kind"] == "v2_sample":
        vb.ops.launch_count = 0
        step_fn(*devb[0]); torch.cuda.synchronize()
        launches_per_step = vb.ops.launch_count

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")

    def timed(n_steps):
        total_ms = 0.0
        last = None
        for i in range(n_steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            last = step_fn(*devb[i % n_data])
            e.record()
            e.synchronize()
            total_ms += s.elapsed_time(e)
        return total_ms, last

    def timed_e2e(n_steps):
        """End to end through the public step API, as an input pipeline runs it: every step's inputs are uploaded from pinned host
        memory (copy stream, double-buffered device staging: the upload of step i+1 overlaps the compute of step i) and every
        step's losses are read back to pinned host memory (asynchronous copy, checked after the region).  One CUDA-event pair
        around the whole region, the first upload included; ms/step = elapsed / n_steps."""
        main = torch.cuda.current_stream()
        copy_s = torch.cuda.Stream()
        stage = [tuple(torch.empty_like(t) for t in devb[0]) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        res_host = torch.zeros(n_steps, 8, dtype=torch.float32).pin_memory()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record(main)
        copy_s.wait_event(s)

        def upload(i):
            k = i % 2
            with torch.cuda.stream(copy_s):
                if i >= 2:
                    copy_s.wait_event(freed[k])                    # step i-2 has consumed this staging pair
                for dst, src in zip(stage[k], host[i % n_data]):
                    dst.copy_(src, non_blocking=True)
                ready[k].record(copy_s)
        upload(0)
        last, nres = None, 0
        for i in range(n_steps):
            if i + 1 < n_steps:
                upload(i + 1)
            main.wait_event(ready[i % 2])
            last = step_fn(*stage[i % 2])
            freed[i % 2].record(main)
            vals = torch.stack([t.float().reshape(()) for t in last])
            nres = vals.numel()
            res_host[i, :nres].copy_(vals, non_blocking=True)      # D2H read of the step's result
        e.record(main)
        e.synchronize()
        torch.cuda.synchronize()
        assert bool(torch.isfinite(res_host[:, :nres]).all()), "non-finite result in the end-to-end region"
        return s.elapsed_time(e), last

    # ---- warm-up, then the device-resident timed region
    warm = max(warmup, 3)
    timed(warm)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    t_wall0 = time.perf_counter()
    ms, last = timed(steps)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    # ---- end to end: pinned host buffers in, losses out, every step
    timed_e2e(2)
    barrier()
    ms_e2e, last_e2e = timed_e2e(steps)
    barrier()
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    losses = [float(x) for x in torch.stack([v.float().reshape(()) for v in last]).cpu()]
    finite = all(x == x and abs(x) != float("inf") for x in losses)
    n_res = len(last_e2e)
    h2d = sum(x.numel() * x.element_size() for x in host[0])

    # ---- release this workload (graph pools, flat buffers, NCCL bucket hooks) before anything else is measured
    torch.cuda.synchronize()
    for bk in (d_b, g_b):
        if bk is not None:
            bk.close()
    del step_fn, last, last_e2e
    gs = None
    if spec["kind"] != "v2_sample":
        del eager, gopt, dopt
        if not args.torch_optim:
            del gnet, dnet
    del gen, disc, devb, host, flush
    gc.collect()
    torch.cuda.empty_cache()
    vb.functional.set_param_grad_stream(False)

    line = None
    if rank == 0:
        imgs = spec["global_batch"] * steps
        value, e2e_value = imgs / (ms * 1e-3), imgs / (ms_e2e * 1e-3)
        fl = step_flops_per_image(spec, skip_unused)
        step_tf = value * fl / world / 1e12          # per GPU
        roof, all_k = kernel_rooflines(vb, spec, pk) if (full or spec["kind"] == "v1") and not args.no_roofline else (None, [])   # rank 0; the others wait at the barrier
        cpu = dropin = None
        if full and world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_rate(spec, steps=2, warmup=1)
            cpu = {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": r["kind"],
                   "sample": f"{'reference modules + loop body (staged src/)' if r['kind'] == 'reference' else 'oracle CPU port of the reference step'}, same model, fixed batch {r['batch']}, {r['warmup']} warm-up + {r['steps']} timed steps, {r['cores']} host threads"}
        if full and world == 1 and not args.no_dropin and spec["kind"] != "v2_sample":
            dropin = dropin_rate(vb, spec, min(B, 256))
        line = {
            "metric": "gan_train_images_per_sec" if unit == "img/s" else "generator_samples_per_sec",
            "value": value, "unit": unit, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": spec["scaling"], "vs_baseline": None,
            "dtype": prec, "data": "synthetic",
            "config": {"workload": f"{spec['name']}: {spec['desc']}", "per_gpu_batch": B, "global_batch": spec["global_batch"],
                       "parallelism": f"dp{world}", "cuda_graph": graph_used, "micro_batches": (n_micro if spec["kind"] != "v2_sample" else 1),
                       "micro_batch_images": B // max(1, n_micro), "kept_generator_graphs": keep_g, "kept_generator_graphs_plan": keep_info, "optimizer": "torch" if args.torch_optim else "fused flat Adam(W) kernel",
                       "d_param_grads_in_g_pass": not skip_unused,
                       "d_update_passes": "D(real) and D(fake) as one concatenated pass (same gradient sum)" if (merge_d and spec["kind"] != "v2_sample" and n_micro == 1) else "separate",
                       "dropout": "p = 0 (parity protocol, SURVEY Q11)",
                       "l2": "192 MiB buffer written between timed steps (L2 flush); step working set is >1 GB anyway"},
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * n_res,
                    "how": "public step API fed from pinned host buffers: per step one H2D upload of its inputs (copy stream, double-buffered staging, overlapping the previous step) and one async D2H read of its losses; one event pair around all steps, no per-step synchronisation and no L2 flush (each step streams >1 GB through the 126 MB L2 and its inputs arrive from the host) -- which is why it can read a little above `value`, whose steps are timed one by one with a flush in between",
                    "ms_per_step": ms_e2e / steps},
            "gpu_launches": launches_per_step * steps, "launches_per_step": launches_per_step,
            "clocks": clocks, "wall_s_timed_region": t_wall, "losses_last_step": losses, "losses_finite": finite,
            "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
            "step_flops_per_image": fl, "step_tflops_per_gpu": step_tf, "step_frac_of_bf16_sustained": step_tf / pk["tf_sust"],
            "step_frac_of_bf16_burst": step_tf / pk["tf_burst"],
            "peaks": pk, "roofline": roof, "kernels": all_k, "cpu_baseline": cpu, "dropin": dropin,
        }
    return line


def dropin_rate(vb, spec, B):
    """What a maintainer following INTEGRATION.md gets in the reference's own loop: the same CUDA-backed modules, UN-graphed, in the
    reference call order (three separate D passes, D's parameter gradients also computed in the G pass) with stock torch Adam(W),
    parameters and gradients left where torch put them.  Batch = one micro-batch of the workload."""
    import gc
    from oracle import harness, v1 as o1, v2 as o2
    torch.manual_seed(0)
    if spec["kind"] == "v2":
        I = spec["over"].get("image_size", 32)
        gan = vb.v2.ViTGAN(vb.v2.Config(**spec["over"], batch_size=3 * I * I)).cuda()
        gen, disc, kind = gan.generator, gan.discriminator, "ce"
        data = harness.synthetic_batches_v2(o2.V2Config(**spec["over"], batch_size=3 * I * I), B, 2)
        go = torch.optim.AdamW(gen.parameters(), lr=5e-4, weight_decay=1e-3)
        do = torch.optim.AdamW(disc.parameters(), lr=5e-4, weight_decay=1e-3)
    else:
        gen = vb.v1.Generator(vb.v1.V1Config(**spec["over"])).cuda()
        disc = vb.v1.Discriminator(vb.v1.V1Config(**spec["over"])).cuda()
        kind = "bce"
        data = harness.synthetic_batches_v1(o1.V1Config(**spec["over"]), B, 2)
        go = torch.optim.Adam(gen.parameters(), lr=2e-4, betas=(0.5, 0.999))
        do = torch.optim.Adam(disc.parameters(), lr=2e-4, betas=(0.5, 0.999))
    data = [(r.cuda(), n.cuda()) for r, n in data]
    vb.set_operand_cache(True)
    n_steps = 6
    for i in range(3):
        vb.train.gan_step(gen, disc, go, do, *data[i % 2], kind)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(n_steps):
        out = vb.train.gan_step(gen, disc, go, do, *data[i % 2], kind)
    e.record()
    e.synchronize()
    ms = s.elapsed_time(e) / n_steps
    ok = all(bool(torch.isfinite(t).all()) for t in out)
    del gen, disc, go, do, data
    gc.collect()
    torch.cuda.empty_cache()
    return {"value": B / (ms * 1e-3), "unit": "img/s", "ms_per_step": ms, "batch": B, "losses_finite": ok,
            "how": "vitgan_b200 modules in the reference's loop body (train.gan_step defaults = src/v2/training.py:177-211 call order), "
                   "eager launches (no CUDA graph), stock torch.optim Adam(W), no flat buffers / fused optimizer / merged passes"}


def run_ours(args):
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("WARN", "VERSION"):
            os.environ.pop("NCCL_DEBUG")          # NCCL would print its version banner on stdout, ahead of the JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = measure(args, args.workload, args.steps, args.warmup, True, rank, world, local)
    secondary = []
    if args.workload == "c4" and not args.no_secondary:
        for name in ("c2", "c3"):
            try:
                rec = measure(args, name, min(args.steps, 20), 3, False, rank, world, local)
                if rank == 0:
                    keep = ("value", "unit", "ms_per_step", "scaling", "dtype", "config", "e2e", "launches_per_step", "clocks", "losses_finite",
                            "step_flops_per_image", "step_tflops_per_gpu", "step_frac_of_bf16_sustained", "kernels")
                    secondary.append({k: rec[k] for k in keep})
            except Exception as ex:      # a failed secondary record never costs the primary line
                if rank == 0:
                    secondary.append({"config": {"workload": name}, "error": f"{type(ex).__name__}: {ex}"[:300]})
    if rank == 0:
        line["secondary"] = secondary
        print(json.dumps(line), flush=True)
    if world > 1:
        # Every captured graph (they hold NCCL work) has been destroyed in measure(); tear the communicator down in order.
        # A watchdog guards the driver's run against a teardown that does not return: everything is already printed.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        dist.destroy_process_group()
        os._exit(0)


if __name__ == "__main__":
    import faulthandler
    if os.environ.get("VG_BENCH_WATCHDOG"):      # debugging aid: dump all Python stacks if the run stalls
        faulthandler.dump_traceback_later(int(os.environ["VG_BENCH_WATCHDOG"]), repeat=False, exit=True)
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
