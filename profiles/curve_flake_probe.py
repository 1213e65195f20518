"""How close does the 200-step loss-curve test (tests/test_gpu_models.py::test_v2_200_step_loss_curve) run to its bounds?
The CUDA path sums gradients in a run-dependent order (TMA reduce-add / atomics), so every run is a different trajectory of a
chaotic map; this prints, per repetition, each quantity the test asserts next to its bound."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vitgan_b200 as vb
from oracle import harness, v2 as o2

vb.set_dropout_policy("off")
fx = torch.load(os.path.join(ROOT, "tests", "golden", "curves_200.pt"))
steps, B = fx["steps"], fx["v2_batch"]
batches = harness.synthetic_batches_v2(o2.V2Config(batch_size=3 * 32 * 32), B, steps)
rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()


def run(prec, n):
    vb.set_precision(prec)
    torch.manual_seed(fx["seed"])
    gan = vb.v2.ViTGAN(vb.v2.Config(batch_size=3 * 32 * 32)).cuda()
    go = torch.optim.AdamW(gan.generator.parameters(), lr=5e-4, weight_decay=1e-3)
    do = torch.optim.AdamW(gan.discriminator.parameters(), lr=5e-4, weight_decay=1e-3)
    return torch.stack([torch.stack(vb.train.gan_step(gan.generator, gan.discriminator, go, do, r.cuda(), n_.cuda(), "ce")).cpu()
                        for r, n_ in batches[:n]]).double()


def stats(cand, first, factor, noise_ratio, floor):
    n = cand.shape[0]
    f32, f64 = fx["v2_f32"][:n], fx["v2_f64"][:n]
    ref_dev = (f32 - f64).abs().amax(1).cummax(0).values
    dev = (cand - f64).abs().amax(1)
    horizon = int((ref_dev * noise_ratio < (1e-2 if noise_ratio == 1.0 else 1.0)).sum())
    bound = factor * noise_ratio * ref_dev[:horizon] + floor
    out = dict(first=rel(cand[:first], f32[:first]), horizon=horizon, env=float((dev[:horizon] / bound).max()))
    if n >= 160:
        out["ratio"] = [round(float(x), 3) for x in (cand[-80:].mean(0) / f64[-80:].mean(0))]
    return out


def run_v1(prec, n):
    from oracle import v1 as o1
    vb.set_precision(prec)
    torch.manual_seed(fx["seed"])
    G, D = vb.v1.Generator(vb.v1.V1Config(image_size=32)).cuda(), vb.v1.Discriminator(vb.v1.V1Config(image_size=32)).cuda()
    go = torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.5, 0.999))
    do = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.5, 0.999))
    b1 = harness.synthetic_batches_v1(o1.V1Config(image_size=32), fx["v1_batch"], fx["steps"])
    return torch.stack([torch.stack(vb.train.gan_step(G, D, go, do, r.cuda(), z.cuda(), "bce")).cpu() for r, z in b1[:n]]).double()


def stats_v1(cand, first, factor, noise_ratio, floor):
    global fx
    keep = fx["v2_f32"], fx["v2_f64"]
    fx["v2_f32"], fx["v2_f64"] = fx["v1_f32"], fx["v1_f64"]
    try:
        return stats(cand, first, factor, noise_ratio, floor)
    finally:
        fx["v2_f32"], fx["v2_f64"] = keep


which = sys.argv[2] if len(sys.argv) > 2 else "v2"
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    if which == "v2":
        print("fp32", stats(run("fp32", steps), 30, 30, 1.0, 1e-5), flush=True)
        print("bf16", stats(run("bf16", 60), 10, 1, 2.0 ** 15, 2e-2), flush=True)
    else:      # the v1 test's parameters: first 8 steps, factor 100 (fp32) / noise ratio 2^15 (bf16)
        print("v1 fp32", stats_v1(run_v1("fp32", steps), 8, 100, 1.0, 1e-5), flush=True)
        print("v1 bf16", stats_v1(run_v1("bf16", 40), 8, 1, 2.0 ** 15, 2e-2), flush=True)
