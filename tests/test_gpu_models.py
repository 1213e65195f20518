"""GPU parity of the CUDA-backed module mirrors (vitgan_b200.v2 / .v1) against
  (a) the golden fixtures generated from the REAL reference (tests/golden/*.pt, oracle/make_golden.py), and
  (b) the CPU oracle run live on the same seeded inputs at the default configs.
fp32 path: 1e-4; bf16 path: 2e-2 (BASELINE.json tolerances, max|a-b|/max|b| per tensor)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import harness, v1 as o1, v2 as o2

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 2e-2}
# gradients pass through more roundings than outputs; bf16 gradient tensors are compared at 3x the block tolerance
GTOL = {"fp32": 2e-4, "bf16": 6e-2}


@pytest.fixture(scope="module")
def vb():
    import vitgan_b200
    return vitgan_b200


@pytest.fixture(params=["fp32", "bf16"])
def prec(request, vb):
    vb.set_precision(request.param)
    yield request.param
    vb.set_precision("bf16")


def rel(a, b):
    return harness.rel_err(a, b)


def load_into(module, params, prefix=""):
    sd = {k[len(prefix):]: v for k, v in params.items() if k.startswith(prefix)}
    missing, unexpected = module.load_state_dict(sd, strict=True)
    return module.cuda()


def check_block(vb, prec, mod, fx, fwd=None):
    x = fx["x"].cuda().requires_grad_(True)
    y = (fwd or mod)(x)
    assert rel(y, fx["y"]) < TOL[prec]
    y.backward(fx["dy"].cuda().to(y.dtype))
    assert rel(x.grad, fx["dx"]) < GTOL[prec]
    grads = dict(mod.named_parameters())
    for k, g in fx["grads"].items():
        assert grads[k].grad is not None, k
        assert rel(grads[k].grad, g) < GTOL[prec], k


def test_v2_blocks_vs_reference_golden(vb, prec, golden):
    fx = golden("v2_blocks")
    v2 = vb.v2
    check_block(vb, prec, load_into(v2.EmbedLayer(3, 64, 16, 4), fx["embed"]["params"]), fx["embed"])
    check_block(vb, prec, load_into(v2.SelfAttention(64, 4), fx["attention"]["params"]), fx["attention"])
    check_block(vb, prec, load_into(v2.Encoder(64, 4, 2), fx["encoder"]["params"]), fx["encoder"])
    check_block(vb, prec, load_into(v2.Classifier(64, 10), fx["classifier"]["params"]), fx["classifier"])


def test_v2_tiny_gan_vs_reference_golden(vb, prec, golden):
    fx = golden("v2_tiny")
    cfg = vb.v2.Config(**fx["config"])
    gan = vb.v2.ViTGAN(cfg)
    assert set(gan.state_dict()) == set(fx["params"])           # same state_dict keys as the reference
    gan.load_state_dict(fx["params"])
    gan = gan.cuda()
    x = fx["x"].cuda().requires_grad_(True)
    d_out = gan.discriminator(x)
    assert d_out.dtype == torch.float32 and rel(d_out, fx["d_out"]) < TOL[prec]
    F.cross_entropy(d_out, torch.ones(3, dtype=torch.long, device="cuda")).backward()
    assert rel(x.grad, fx["d_dx"]) < GTOL[prec]
    named = dict(gan.discriminator.named_parameters())
    for k, g in fx["d_grads"].items():
        assert rel(named[k].grad, g) < GTOL[prec], k
    gan.zero_grad(set_to_none=True)
    g_out = gan.generator(fx["z"].cuda())
    assert g_out.shape == fx["g_out"].shape and rel(g_out, fx["g_out"]) < TOL[prec]
    F.cross_entropy(gan.discriminator(g_out), torch.ones(3, dtype=torch.long, device="cuda")).backward()
    named = dict(gan.generator.named_parameters())
    for k, g in fx["g_grads"].items():
        assert rel(named[k].grad, g) < GTOL[prec], k


def test_v2_tiny_three_steps_vs_reference_golden(vb, prec, golden):
    """The reference's own step sequence (src/v2/training.py:177-211) with torch AdamW on the patched-in kernels."""
    fx = golden("v2_tiny")
    cfg = vb.v2.Config(**fx["config"])
    gan = vb.v2.ViTGAN(cfg)
    gan.load_state_dict(fx["params"])
    gan = gan.cuda()
    go = torch.optim.AdamW(gan.generator.parameters(), lr=cfg.generator_learning_rate, weight_decay=1e-3)
    do = torch.optim.AdamW(gan.discriminator.parameters(), lr=cfg.discriminator_learning_rate, weight_decay=1e-3)
    ocfg = o2.V2Config(**{k: v for k, v in fx["config"].items()})
    losses = []
    for real, noise in harness.synthetic_batches_v2(ocfg, 3, 3):
        losses.append(torch.stack(vb.train.gan_step(gan.generator, gan.discriminator, go, do, real.cuda(), noise.cuda(), "ce")))
    losses = torch.stack(losses)
    assert rel(losses, fx["losses"]) < TOL[prec]
    if prec == "fp32":
        for k, v in fx["params_after"].items():
            assert rel(gan.state_dict()[k], v) < 5e-3, k     # Adam's g/sqrt(v) amplifies 1e-6 gradient noise


def test_v2_default_config_vs_oracle_live(vb, prec, golden):
    """main-v2 defaults (E128 L6 H4 S65), B=8: outputs vs golden + 5-step loss curve vs golden and the live oracle."""
    fx = golden("v2_default")
    cfg = vb.v2.Config(batch_size=3 * 32 * 32)
    torch.manual_seed(fx["seed"])
    gan = vb.v2.ViTGAN(cfg)                                   # seeded init == reference init (checked on CPU)
    ocfg = o2.V2Config(batch_size=cfg.batch_size)
    gan = gan.cuda()
    (real, noise), = harness.synthetic_batches_v2(ocfg, fx["batch"], 1, seed=fx["data_seed"])
    with torch.no_grad():
        assert rel(gan.discriminator(real.cuda()), fx["d_out"]) < TOL[prec]
        g = gan.generator(noise.cuda())
    assert rel(g[:, :, :4, :4], fx["g_out_slice"]) < TOL[prec]
    go = torch.optim.AdamW(gan.generator.parameters(), lr=5e-4, weight_decay=1e-3)
    do = torch.optim.AdamW(gan.discriminator.parameters(), lr=5e-4, weight_decay=1e-3)
    losses = torch.stack([torch.stack(vb.train.gan_step(gan.generator, gan.discriminator, go, do, r.cuda(), n.cuda(), "ce"))
                          for r, n in harness.synthetic_batches_v2(ocfg, fx["batch"], 5)])
    assert rel(losses, fx["losses"]) < TOL[prec]


def test_v2_fused_adam_flatnet_and_skip_grads(vb, golden):
    """FlatNet + FusedAdam (one launch per network) and skip_unused_d_grads reproduce the torch-AdamW loss curve."""
    vb.set_precision("fp32")
    fx = golden("v2_tiny")
    cfg = vb.v2.Config(**fx["config"])
    gan = vb.v2.ViTGAN(cfg)
    gan.load_state_dict(fx["params"])
    gan = gan.cuda()
    keys = list(gan.state_dict().keys())
    gnet, dnet = vb.train.FlatNet(gan.generator), vb.train.FlatNet(gan.discriminator)
    assert list(gan.state_dict().keys()) == keys
    go = vb.train.FusedAdam(gnet, cfg.generator_learning_rate, weight_decay=1e-3, decoupled=True)
    do = vb.train.FusedAdam(dnet, cfg.discriminator_learning_rate, weight_decay=1e-3, decoupled=True)
    ocfg = o2.V2Config(**fx["config"])
    losses = torch.stack([torch.stack(vb.train.gan_step(gan.generator, gan.discriminator, go, do, r.cuda(), n.cuda(), "ce",
                                                         skip_unused_d_grads=True))
                          for r, n in harness.synthetic_batches_v2(ocfg, 3, 3)])
    assert rel(losses, fx["losses"]) < 1e-4
    vb.set_precision("bf16")


def test_v2_graphed_step_matches_eager(vb, golden):
    vb.set_precision("bf16")
    fx = golden("v2_tiny")
    cfg = vb.v2.Config(**fx["config"])
    ocfg = o2.V2Config(**fx["config"])
    batches = harness.synthetic_batches_v2(ocfg, 3, 6)

    def build():
        gan = vb.v2.ViTGAN(cfg)
        gan.load_state_dict(fx["params"])
        gan = gan.cuda()
        gnet, dnet = vb.train.FlatNet(gan.generator), vb.train.FlatNet(gan.discriminator)
        go = vb.train.FusedAdam(gnet, 5e-4, weight_decay=1e-3, decoupled=True)
        do = vb.train.FusedAdam(dnet, 5e-4, weight_decay=1e-3, decoupled=True)
        return gan, go, do

    gan, go, do = build()
    eager = [torch.stack(vb.train.gan_step(gan.generator, gan.discriminator, go, do, r.cuda(), n.cuda(), "ce")).cpu() for r, n in batches]
    gan, go, do = build()
    r0, n0 = batches[0]
    step = vb.train.GraphedStep(gan.generator, gan.discriminator, go, do, r0.cuda(), n0.cuda(), "ce", warmup=0)
    # the capture itself does not execute; replay all six batches
    graphed = [torch.stack(step(r.cuda(), n.cuda())).cpu().clone() for r, n in batches]
    vb.set_operand_cache(True)
    assert rel(torch.stack(graphed), torch.stack(eager)) < 2e-2


# ------------------------------------------------------------------------------------------------ v1
def test_v1_blocks_vs_reference_golden(vb, prec, golden):
    fx = golden("v1_blocks")
    v1 = vb.v1
    # SLN
    b = fx["sln"]
    m = load_into(v1.SLN(48), b["params"])
    h, w = b["h"].cuda().requires_grad_(True), b["w"].cuda().requires_grad_(True)
    y = m(h, w)
    assert rel(y, b["y"]) < TOL[prec]
    y.backward(b["dy"].cuda().to(y.dtype))
    assert rel(h.grad, b["dh"]) < GTOL[prec] and rel(w.grad, b["dw"]) < GTOL[prec]
    for k, g in b["grads"].items():
        assert rel(dict(m.named_parameters())[k].grad, g) < GTOL[prec], k
    # multi-head attention: dot (G) and L2 + spectral rescale (D)
    for lp in (1, 2):
        b = fx[f"msha_lp{lp}"]
        tp = v1.TransformerParameters(input_features=48, spectral_scaling=(lp == 2), lp=lp)
        m = v1.MultiHeadSelfAttention(tp, output_size=48, head_dimension=12)
        m.load_state_dict(b["params"])
        if lp == 2:
            for hd, sp in zip(m.attention_heads, b["init_spectrum"]):
                hd.init_spectrum = [torch.tensor(s) for s in sp]       # the reference's construction-time SVD values
            m.train_qkv = True                                           # expose dL/dW_eff for the parity check (Q4)
        m = m.cuda()
        check_block(vb, prec, m, b)
    # discriminator block
    b = fx["transformer_d"]
    tp = v1.TransformerParameters(input_features=48, spectral_scaling=True, lp=2)
    m = v1.Transformer(tp)
    m.load_state_dict(b["params"])
    for hd in m.msha.attention_heads:                                   # init spectrum of the loaded weights
        hd.init_spectrum = [torch.linalg.svdvals(w.weight.detach()).max() for w in (hd.q, hd.k, hd.v)]
    m.msha.train_qkv = True
    check_block(vb, prec, m.cuda(), b)
    # generator block with the first-layer (S,F) broadcast
    b = fx["transformer_sln"]
    tp = v1.TransformerParameters(input_features=48, spectral_scaling=False, lp=1)
    m = load_into(v1.TransformerSLN(tp), b["params"])
    h, w = b["h"].cuda().requires_grad_(True), b["w"].cuda().requires_grad_(True)
    _, hf = m(h, w)
    assert rel(hf, b["hf"]) < TOL[prec]
    hf.backward(b["dy"].cuda().to(hf.dtype))
    assert rel(h.grad, b["dh"]) < GTOL[prec] and rel(w.grad, b["dw"]) < GTOL[prec]
    for k, g in b["grads"].items():
        assert rel(dict(m.named_parameters())[k].grad, g) < GTOL[prec], k
    # SIREN
    b = fx["siren"]
    m = load_into(v1.SIREN(v1.SIRENParameters(48, 40, is_first=True)), b["params"])
    check_block(vb, prec, m, b)
    # patch encoder (scrambled layout), 24 of 432 output features kept in the fixture
    b = fx["patch_encoder"]
    pe = v1.PatchEncoder(v1.V1Config(image_size=32), projection_output_size=24)
    with torch.no_grad():
        pe.projection_matrix.weight.copy_(b["proj_w"]); pe.cls_token.copy_(b["cls"]); pe.positional_embedding.copy_(b["pos"])
    assert rel(pe.cuda()(b["x"].cuda()), b["y"]) < TOL[prec]


def test_v1_default_vs_reference_golden_and_oracle(vb, prec, golden):
    fx = golden("v1_default")
    I = fx["image_size"]
    torch.manual_seed(fx["seed"])
    G = vb.v1.Generator(vb.v1.V1Config(image_size=I))
    D = vb.v1.Discriminator(vb.v1.V1Config(image_size=I))      # same RNG stream order as the reference (G then D)
    G, D = G.cuda(), D.cuda()
    ocfg = o1.V1Config(image_size=I)
    (real, z), = harness.synthetic_batches_v1(ocfg, fx["batch"], 1, seed=fx["data_seed"])
    with torch.no_grad():
        assert rel(D(real.cuda()), fx["d_out"]) < TOL[prec]
        g = G(z.cuda())
    assert g.shape == (fx["batch"], 3, I, I)
    # sin(30*(.)) twice: the bf16 path's activation rounding is amplified by omega_0 = 30 per SIREN layer
    assert rel(g[:, :, :4, :4], fx["g_out_slice"]) < (TOL[prec] if prec == "fp32" else 0.25)
    go = torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.5, 0.999))
    do = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.5, 0.999))
    losses = torch.stack([torch.stack(vb.train.gan_step(G, D, go, do, r.cuda(), zz.cuda(), "bce"))
                          for r, zz in harness.synthetic_batches_v1(ocfg, fx["batch"], 2)])
    assert rel(losses, fx["losses"]) < (5e-4 if prec == "fp32" else 5e-2)
    # Q4: the discriminator's q/k/v never receive gradients / updates
    assert all(p.grad is None for n, p in D.named_parameters() if n.endswith((".q.weight", ".k.weight", ".v.weight")))


def test_v1_full_grads_vs_oracle_64px(vb):
    """Config C3 geometry (64 px: G S=64 F=384, D S=65 F=432 d=108), fp32 path: every parameter gradient of one
    D pass and one G pass against the CPU oracle."""
    vb.set_precision("fp32")
    cfg = o1.V1Config(image_size=64)
    orc = harness.OracleV1(cfg, seed=3)
    G = vb.v1.Generator(vb.v1.V1Config(image_size=64)); D = vb.v1.Discriminator(vb.v1.V1Config(image_size=64))
    G.load_state_dict({k[len("generator."):]: v.detach() for k, v in orc.p.items() if k.startswith("generator.")})
    D.load_state_dict({k[len("discriminator."):]: v.detach() for k, v in orc.p.items() if k.startswith("discriminator.")})
    for blk in D.transformer_layers:
        for hd in blk.msha.attention_heads:
            hd.init_spectrum = [torch.linalg.svdvals(w.weight.detach()).max() for w in (hd.q, hd.k, hd.v)]
        blk.msha.train_qkv = True
    G, D = G.cuda(), D.cuda()
    (real, z), = harness.synthetic_batches_v1(cfg, 2, 1, seed=5)
    out_o = orc.discriminator(real)
    fake_o = orc.generator(z)
    loss_o = F.binary_cross_entropy(out_o, torch.ones(2, 1)) + F.binary_cross_entropy(orc.discriminator(fake_o), torch.ones(2, 1))
    loss_o.backward()
    out = D(real.cuda())
    fake = G(z.cuda())
    assert rel(out, out_o) < 1e-4 and rel(fake, fake_o) < 1e-4
    loss = F.binary_cross_entropy(out, torch.ones(2, 1, device="cuda")) + F.binary_cross_entropy(D(fake), torch.ones(2, 1, device="cuda"))
    loss.backward()
    for name, mod in (("generator.", G), ("discriminator.", D)):
        for k, p in mod.named_parameters():
            assert p.grad is not None, k
            assert rel(p.grad, orc.p[name + k].grad) < 5e-4, k
    vb.set_precision("bf16")
