"""The oracle (CPU restatement) must reproduce the fixtures produced by the REAL reference
(oracle/make_golden.py).  CPU-only; this is what pins the oracle."""
import torch
import torch.nn.functional as F

from oracle import harness, v1 as o1, v2 as o2

TOL = 1e-6  # same torch build -> observed bit-exact; tolerance only guards BLAS thread-count effects


def close(a, b, tol=TOL):
    return harness.rel_err(a, b) <= tol


def test_v2_tiny_outputs_grads_and_losses(golden):
    fx = golden("v2_tiny")
    cfg = o2.V2Config(**fx["config"])
    # seeded oracle init == reference init
    p0 = o2.init_vitgan(cfg, seed=fx["seed"])
    assert set(p0) == set(fx["params"])
    assert all(torch.equal(p0[k], fx["params"][k]) for k in p0)
    p = {k: v.clone().requires_grad_(True) for k, v in fx["params"].items()}
    x = fx["x"].clone().requires_grad_(True)
    d_out = o2.vit_discriminator(p, "discriminator.", x, cfg)
    assert close(d_out, fx["d_out"])
    F.cross_entropy(d_out, torch.ones(3, dtype=torch.long)).backward()
    assert close(x.grad, fx["d_dx"])
    for k, g in fx["d_grads"].items():
        assert close(p["discriminator." + k].grad, g), k
    for v in p.values():
        v.grad = None
    g_out = o2.vit_generator(p, "generator.", fx["z"], cfg)
    assert close(g_out, fx["g_out"])
    F.cross_entropy(o2.vit_discriminator(p, "discriminator.", g_out, cfg), torch.ones(3, dtype=torch.long)).backward()
    for k, g in fx["g_grads"].items():
        assert close(p["generator." + k].grad, g), k
    orc = harness.OracleV2(cfg, seed=fx["seed"])
    losses = torch.stack([torch.stack(orc.step(r, n)) for r, n in harness.synthetic_batches_v2(cfg, 3, 3)])
    assert close(losses, fx["losses"])
    for k, v in fx["params_after"].items():
        assert close(orc.p[k], v, 1e-5), k


def test_v2_blocks(golden):
    fx = golden("v2_blocks")
    fns = {
        "embed": lambda p, x: o2.embed_layer(p, "", x, 4),
        "attention": lambda p, x: o2.self_attention(p, "", x, 4),
        "encoder": lambda p, x: o2.encoder(p, "", x, 4),
        "classifier": lambda p, x: o2.classifier(p, "", x),
    }
    for name, fn in fns.items():
        b = fx[name]
        p = {k: v.clone().requires_grad_(True) for k, v in b["params"].items()}
        x = b["x"].clone().requires_grad_(True)
        y = fn(p, x)
        assert close(y, b["y"]), name
        y.backward(b["dy"])
        assert close(x.grad, b["dx"]), name
        for k, g in b["grads"].items():
            assert close(p[k].grad, g), (name, k)


def test_v2_default_config(golden):
    fx = golden("v2_default")
    cfg = o2.V2Config(batch_size=3 * 32 * 32)
    orc = harness.OracleV2(cfg, seed=fx["seed"])
    (real, noise), = harness.synthetic_batches_v2(cfg, fx["batch"], 1, seed=fx["data_seed"])
    with torch.no_grad():
        assert close(orc.discriminator(real), fx["d_out"])
        g = orc.generator(noise)
    assert close(g[:, :, :4, :4], fx["g_out_slice"]) and close(g.mean(), fx["g_out_mean"])
    losses = torch.stack([torch.stack(orc.step(r, n)) for r, n in harness.synthetic_batches_v2(cfg, fx["batch"], 5)])
    assert close(losses, fx["losses"], 1e-5)


def test_v1_blocks(golden):
    fx = golden("v1_blocks")
    b = fx["sln"]
    p = {k: v.clone().requires_grad_(True) for k, v in b["params"].items()}
    h, w = b["h"].clone().requires_grad_(True), b["w"].clone().requires_grad_(True)
    y = o1.sln(p, "", h, w)
    assert close(y, b["y"])
    y.backward(b["dy"])
    assert close(h.grad, b["dh"]) and close(w.grad, b["dw"])
    for k, g in b["grads"].items():
        assert close(p[k].grad, g), k

    for lp in (1, 2):
        b = fx[f"msha_lp{lp}"]
        p = {k: v.clone().requires_grad_(True) for k, v in b["params"].items()}
        spectra = None
        if lp == 2:
            spectra = {f"attention_heads.{i}.": tuple(torch.tensor(s) for s in sp) for i, sp in enumerate(b["init_spectrum"])}
        x = b["x"].clone().requires_grad_(True)
        y = o1.multi_head_self_attention(p, "", x, 4, lp, spectra)
        assert close(y, b["y"]), lp
        y.backward(b["dy"])
        assert close(x.grad, b["dx"], 1e-5), lp
        for k, g in b["grads"].items():
            assert close(p[k].grad, g, 1e-5), (lp, k)

    b = fx["transformer_d"]
    p = {k: v.clone().requires_grad_(True) for k, v in b["params"].items()}
    cfg = o1.V1Config(d_layers=1)
    spectra = {"msha.attention_heads.%d." % i: tuple(o1.sigma_max(b["params"]["msha.attention_heads.%d.%s.weight" % (i, n)])
                                                     for n in "qkv") for i in range(4)}
    x = b["x"].clone().requires_grad_(True)
    y = o1.transformer(p, "", x, 4, spectra)
    assert close(y, b["y"])
    y.backward(b["dy"])
    assert close(x.grad, b["dx"], 1e-5)
    for k, g in b["grads"].items():
        assert close(p[k].grad, g, 1e-5), k

    b = fx["transformer_sln"]
    p = {k: v.clone().requires_grad_(True) for k, v in b["params"].items()}
    h, w = b["h"].clone().requires_grad_(True), b["w"].clone().requires_grad_(True)
    _, hf = o1.transformer_sln(p, "", h, w, 4)
    assert close(hf, b["hf"])
    hf.backward(b["dy"])
    assert close(h.grad, b["dh"], 1e-5) and close(w.grad, b["dw"], 1e-5)
    for k, g in b["grads"].items():
        assert close(p[k].grad, g, 1e-5), k

    b = fx["siren"]
    p = {k: v.clone().requires_grad_(True) for k, v in b["params"].items()}
    x = b["x"].clone().requires_grad_(True)
    y = o1.siren(p, "", x, 30)
    assert close(y, b["y"])
    y.backward(b["dy"])
    assert close(x.grad, b["dx"])

    b = fx["patch_encoder"]
    cfg = o1.V1Config(image_size=32)
    tok = o1.get_tokens(b["x"], cfg)
    assert tok.shape == (2, 49, 432)
    assert torch.equal(tok[:, 5, :], b["tokens_t5"]) and tok.double().sum() == b["tokens_checksum"]
    p = {"projection_matrix.weight": b["proj_w"], "cls_token": b["cls"], "positional_embedding": b["pos"]}
    assert close(o1.patch_encoder(p, "", b["x"], cfg), b["y"])


def test_v1_default(golden):
    fx = golden("v1_default")
    cfg = o1.V1Config(image_size=fx["image_size"])
    orc = harness.OracleV1(cfg, seed=fx["seed"])
    (real, z), = harness.synthetic_batches_v1(cfg, fx["batch"], 1, seed=fx["data_seed"])
    with torch.no_grad():
        assert close(orc.discriminator(real), fx["d_out"])
        g = orc.generator(z)
    assert close(g[:, :, :4, :4], fx["g_out_slice"]) and close(g.mean(), fx["g_out_mean"])
    losses = torch.stack([torch.stack(orc.step(r, zz)) for r, zz in harness.synthetic_batches_v1(cfg, fx["batch"], 2)])
    assert close(losses, fx["losses"], 1e-5)


def test_oracle_convert_to_uint8_known_answers():
    """utils.convert_to_uint8 (src/v2/utils.py:194-196): known answers at the clamp edges and the mid-points."""
    from oracle import v2 as o2
    x = torch.tensor([-2.0, -1.0, -0.5, 0.0, 0.5, 1.0, 2.0, 0.999, -0.999])
    assert o2.convert_to_uint8(x).tolist() == [0, 0, 63, 127, 191, 255, 255, 254, 0]


def test_oracle_reproduces_reference_loss_curves(golden):
    """tests/golden/curves_200.pt holds 200-step loss curves of the REAL reference modules (fp32 and .double()).  The oracle's
    step must reproduce their heads bit for bit in fp32 (v2: 12 steps, v1: 3 steps -- the oracle is a restatement, so the whole
    curve is identical; the CPU suite only re-runs the head to stay within its time budget) and to 1e-12 in fp64."""
    fx = golden("curves_200")
    c2 = o2.V2Config(batch_size=3 * 32 * 32)
    for dt, key, tol in ((torch.float32, "v2_f32", 0.0), (torch.float64, "v2_f64", 1e-12)):
        orc = harness.OracleV2(c2, seed=fx["seed"], dtype=dt)
        got = torch.stack([torch.stack(orc.step(r.to(dt), n.to(dt))) for r, n in harness.synthetic_batches_v2(c2, fx["v2_batch"], 12)]).double()
        assert (got - fx[key][:12]).abs().max() <= tol, key
    c1 = o1.V1Config(image_size=32)
    orc = harness.OracleV1(c1, seed=fx["seed"])
    got = torch.stack([torch.stack(orc.step(r, z)) for r, z in harness.synthetic_batches_v1(c1, fx["v1_batch"], 3)]).double()
    assert (got - fx["v1_f32"][:3]).abs().max() == 0.0
