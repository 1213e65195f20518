"""Generate tests/golden/*.pt from the REAL reference (imported from /root/reference).  Test infrastructure.

Run in the build container only:  ``python -m oracle.make_golden``
The fixtures are small (tiny configs carry their parameters; default-size configs carry only the
seed and the outputs, because the oracle's seeded init is bit-identical to the reference's).
They pin the oracle (tests/test_oracle_golden.py) and are also what the GPU parity tests compare
the CUDA path against, since /root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F

from . import refimport
from .harness import gan_step, synthetic_batches_v1, synthetic_batches_v2
from . import v1 as o1, v2 as o2

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _grads(module):
    return {k: v.grad.detach().clone() for k, v in module.named_parameters() if v.grad is not None}


def v2_tiny():
    """Whole tiny v2 GAN: params, D/G outputs, all grads of one D pass and one G pass, 3-step losses."""
    over = dict(embeddings_dimension=32, attention_heads_count=2, transformer_blocks_count=2,
                image_size=16, patch_size=4, mlp_ratio=2, classes_count=10)
    gan, c = refimport.build_v2(seed=7, **over)
    params = {k: v.detach().clone() for k, v in gan.state_dict().items()}
    g = torch.Generator().manual_seed(11)
    x = torch.rand(3, 3, 16, 16, generator=g) * 2 - 1
    z = torch.randn(3, 3, 16, 16, generator=g)
    xg = x.clone().requires_grad_(True)
    d_out = gan.discriminator(xg)
    F.cross_entropy(d_out, torch.ones(3, dtype=torch.long)).backward()
    d_grads, dx = _grads(gan.discriminator), xg.grad.clone()
    gan.zero_grad(set_to_none=True)
    g_out = gan.generator(z)
    F.cross_entropy(gan.discriminator(g_out), torch.ones(3, dtype=torch.long)).backward()
    g_grads = _grads(gan.generator)
    # 3 steps of the real loop on a fresh copy
    gan2, _ = refimport.build_v2(seed=7, **over)
    go = torch.optim.AdamW(gan2.generator.parameters(), lr=c.generator_learning_rate, weight_decay=1e-3)
    do = torch.optim.AdamW(gan2.discriminator.parameters(), lr=c.discriminator_learning_rate, weight_decay=1e-3)
    cfg = o2.V2Config(**over, batch_size=c.batch_size)
    losses = [torch.stack(gan_step(gan2.generator, gan2.discriminator, go, do, r, n, "ce"))
              for r, n in synthetic_batches_v2(cfg, 3, 3)]
    return dict(config={**over, "batch_size": c.batch_size}, seed=7, params=params, x=x, z=z,
                d_out=d_out.detach(), d_grads=d_grads, d_dx=dx, g_out=g_out.detach(), g_grads=g_grads,
                losses=torch.stack(losses), params_after={k: v.detach().clone() for k, v in gan2.state_dict().items()})


def v2_blocks():
    """Per-block fixtures (inputs, params, outputs, grads) for EmbedLayer, SelfAttention, Encoder, Classifier."""
    m = refimport.v2_modules()
    torch.manual_seed(3)
    g = torch.Generator().manual_seed(5)
    out = {}

    def run(name, mod, x):
        for prm in mod.parameters():          # spread the zero-initialised biases / pos / cls
            if prm.abs().max() == 0:
                prm.data.normal_(0, 0.05, generator=g)
        xg = x.clone().requires_grad_(True)
        y = mod(xg)
        dy = torch.randn(y.shape, generator=g)
        y.backward(dy)
        out[name] = dict(params={k: v.detach().clone() for k, v in mod.state_dict().items()}, x=x, y=y.detach(),
                         dy=dy, dx=xg.grad.clone(), grads=_grads(mod))

    run("embed", m.EmbedLayer(3, 64, 16, 4, dropout=0.0), torch.randn(2, 3, 16, 16, generator=g))
    run("attention", m.SelfAttention(64, 4), torch.randn(2, 17, 64, generator=g))
    run("encoder", refimport.zero_dropout(m.Encoder(64, 4, 2, dropout=0.1)), torch.randn(2, 17, 64, generator=g))
    run("classifier", m.Classifier(64, 10), torch.randn(2, 17, 64, generator=g))
    return out


def v2_default():
    """Default main-v2 config (C1 shape at B=8): seed + outputs only (init is reproducible from the seed)."""
    gan, c = refimport.build_v2(seed=0)
    cfg = o2.V2Config(batch_size=c.batch_size)
    (real, noise), = synthetic_batches_v2(cfg, 8, 1, seed=99)
    d_out = gan.discriminator(real)
    g_out = gan.generator(noise)
    gan2, _ = refimport.build_v2(seed=0)
    go = torch.optim.AdamW(gan2.generator.parameters(), lr=5e-4, weight_decay=1e-3)
    do = torch.optim.AdamW(gan2.discriminator.parameters(), lr=5e-4, weight_decay=1e-3)
    losses = [torch.stack(gan_step(gan2.generator, gan2.discriminator, go, do, r, n, "ce"))
              for r, n in synthetic_batches_v2(cfg, 8, 5)]
    return dict(seed=0, data_seed=99, batch=8, d_out=d_out.detach(), g_out_slice=g_out.detach()[:, :, :4, :4].clone(),
                g_out_mean=g_out.detach().mean(), losses=torch.stack(losses))


def v1_blocks():
    """v1 per-block fixtures at reduced width: SLN, L1/L2 Attention head, MHSA, Transformer(SLN), SIREN, PatchEncoder."""
    mods = refimport.v1_modules()
    cfgm = mods["config"]
    cfgm.config.image_size = 32
    g = torch.Generator().manual_seed(17)
    torch.manual_seed(19)
    out = {}

    def grads_of(mod):
        return _grads(mod)

    # SLN
    s = mods["spectral_layer_norm"].SLN(48)
    h, w = torch.randn(2, 9, 48, generator=g, requires_grad=True), torch.randn(2, 9, 48, generator=g, requires_grad=True)
    y = s(h, w); dy = torch.randn(y.shape, generator=g); y.backward(dy)
    out["sln"] = dict(params=dict(s.state_dict()), h=h.detach(), w=w.detach(), y=y.detach(), dy=dy,
                      dh=h.grad.clone(), dw=w.grad.clone(), grads=grads_of(s))
    # attention heads (lp=1 dot / lp=2 L2-distance with S>25 so cdist takes the matmul path)
    for lp, S in ((1, 12), (2, 30)):
        tp = cfgm.TransformerParameters(input_features=48, spectral_scaling=(lp == 2), lp=lp)
        msa = mods["attention"].MultiHeadSelfAttention(tp, output_size=48, head_dimension=12)
        x = torch.randn(2, S, 48, generator=g, requires_grad=True)
        pre = {k: v.detach().clone() for k, v in msa.state_dict().items()}
        spectra = [[float(v) for v in hd.init_spectrum] for hd in msa.attention_heads] if lp == 2 else None
        y = msa(x); dy = torch.randn(y.shape, generator=g); y.backward(dy)
        # NB (Q4): after forward the q/k/v weights are fresh Parameters; their grads live on those
        out[f"msha_lp{lp}"] = dict(params=pre, x=x.detach(), y=y.detach(), dy=dy, dx=x.grad.clone(),
                                   grads=grads_of(msa), init_spectrum=spectra)
    # D transformer block / G transformer block (dropout zeroed)
    tpd = cfgm.TransformerParameters(input_features=48, spectral_scaling=True, lp=2)
    t = refimport.zero_dropout(mods["transformer"].Transformer(tpd))
    x = torch.randn(2, 30, 48, generator=g, requires_grad=True)
    pre = {k: v.detach().clone() for k, v in t.state_dict().items()}
    y = t(x); dy = torch.randn(y.shape, generator=g); y.backward(dy)
    out["transformer_d"] = dict(params=pre, x=x.detach(), y=y.detach(), dy=dy, dx=x.grad.clone(), grads=grads_of(t))
    tpg = cfgm.TransformerParameters(input_features=48, spectral_scaling=False, lp=1)
    t = refimport.zero_dropout(mods["transformer"].TransformerSLN(tpg))
    h = torch.randn(9, 48, generator=g, requires_grad=True)       # (S,F): first-layer broadcast case
    w = torch.randn(2, 9, 48, generator=g, requires_grad=True)
    _, hf = t(h, w); dy = torch.randn(hf.shape, generator=g); hf.backward(dy)
    out["transformer_sln"] = dict(params=dict(t.state_dict()), h=h.detach(), w=w.detach(), hf=hf.detach(), dy=dy,
                                  dh=h.grad.clone(), dw=w.grad.clone(), grads=grads_of(t))
    # SIREN
    sp = mods["siren"].SIRENParameters(input_features=48, output_features=40, is_first=True)
    sr = mods["siren"].SIREN(sp)
    x = torch.randn(2, 9, 48, generator=g, requires_grad=True)
    y = sr(x); dy = torch.randn(y.shape, generator=g); y.backward(dy)
    out["siren"] = dict(params=dict(sr.state_dict()), x=x.detach(), y=y.detach(), dy=dy, dx=x.grad.clone(), grads=grads_of(sr))
    # PatchEncoder (scrambled token layout) at 32 px
    mods["patch_encoder"].PatchEncoder.projection_output_size = 432
    pe = mods["patch_encoder"].PatchEncoder(cfgm.EncoderParameters())
    x = torch.randn(2, 3, 32, 32, generator=g, requires_grad=True)
    tok = pe._get_tokens(x.detach())
    y = pe(x); dy = torch.randn(y.shape, generator=g) * 0.1; y.backward(dy)
    keep = slice(0, 24)   # fixture stays small: keep 24 output features
    out["patch_encoder"] = dict(x=x.detach(), tokens_checksum=tok.double().sum(), tokens_t5=tok[:, 5, :].clone(),
                                seed_note="params re-created in test from state below",
                                proj_w=pe.projection_matrix.weight.detach()[keep].clone(),
                                cls=pe.cls_token.detach()[..., keep].clone(),
                                pos=pe.positional_embedding.detach()[:, keep].clone(),
                                y=y.detach()[..., keep].clone())
    return out


def v1_default(image_size=32):
    """Default v1 G and D at 32 px, B=2: seed + outputs + 2-step losses (params reproducible from the seed)."""
    G, D = refimport.build_v1(image_size, seed=0)
    cfg = o1.V1Config(image_size=image_size)
    (real, z), = synthetic_batches_v1(cfg, 2, 1, seed=77)
    d_out = D(real)
    g_out = G(z)
    G2, D2 = refimport.build_v1(image_size, seed=0)
    go = torch.optim.Adam(G2.parameters(), lr=2e-4, betas=(0.5, 0.999))
    do = torch.optim.Adam(D2.parameters(), lr=2e-4, betas=(0.5, 0.999))      # built BEFORE any forward (Q4)
    losses = [torch.stack(gan_step(G2, D2, go, do, r, zz, "bce")) for r, zz in synthetic_batches_v1(cfg, 2, 2)]
    return dict(seed=0, data_seed=77, batch=2, image_size=image_size, d_out=d_out.detach(),
                g_out_slice=g_out.detach()[:, :, :4, :4].clone(), g_out_mean=g_out.detach().mean(),
                losses=torch.stack(losses))


def curves_200():
    """200-step loss curves of the REAL reference modules under the reference's own step sequence, in fp32 and in fp64 (the same
    modules after .double()): the pair calibrates the loss-curve parity tests (the fp32 reference's own drift from fp64 is the
    yardstick).  v2: default model, B = 8, AdamW(5e-4, wd 1e-3); v1: 32 px, B = 2, Adam(2e-4, (0.5, 0.999))."""
    out = {}
    steps = 200
    c2 = o2.V2Config(batch_size=3 * 32 * 32)
    b2 = synthetic_batches_v2(c2, 8, steps)
    for name, dt in (("v2_f32", torch.float32), ("v2_f64", torch.float64)):
        gan, c = refimport.build_v2(seed=0)
        gan = gan.to(dt)
        go = torch.optim.AdamW(gan.generator.parameters(), lr=c.generator_learning_rate, weight_decay=1e-3)
        do = torch.optim.AdamW(gan.discriminator.parameters(), lr=c.discriminator_learning_rate, weight_decay=1e-3)
        out[name] = torch.stack([torch.stack(gan_step(gan.generator, gan.discriminator, go, do, r.to(dt), n.to(dt), "ce")) for r, n in b2]).double()
    c1 = o1.V1Config(image_size=32)
    b1 = synthetic_batches_v1(c1, 2, steps)
    for name, dt in (("v1_f32", torch.float32), ("v1_f64", torch.float64)):
        G, D = refimport.build_v1(32, seed=0)
        G, D = G.to(dt), D.to(dt)
        for m in D.modules():                                   # the construction-time spectra are python floats of fp32 SVDs; keep them
            if hasattr(m, "init_spectrum"):
                m.init_spectrum = [torch.as_tensor(s, dtype=dt) if not torch.is_tensor(s) else s.to(dt) for s in m.init_spectrum]
        go = torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.5, 0.999))
        do = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.5, 0.999))
        out[name] = torch.stack([torch.stack(gan_step(G, D, go, do, r.to(dt), z.to(dt), "bce")) for r, z in b1]).double()
    out["steps"], out["v2_batch"], out["v1_batch"], out["seed"], out["data_seed"] = steps, 8, 2, 0, 1234
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    assert refimport.available(), "needs /root/reference"
    import sys
    only = sys.argv[1:]
    for name, fn in (("v2_tiny", v2_tiny), ("v2_blocks", v2_blocks), ("v2_default", v2_default),
                     ("v1_blocks", v1_blocks), ("v1_default", v1_default), ("curves_200", curves_200)):
        if only and name not in only:
            continue
        obj = fn()
        path = os.path.join(OUT, name + ".pt")
        torch.save(obj, path)
        print(f"wrote {path}  {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
