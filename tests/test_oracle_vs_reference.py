"""Live check of the oracle against the REAL reference modules (bit-exact on CPU).
Runs only where /root/reference exists (the build container); skipped on the GPU box."""
import pytest
import torch

from oracle import harness, refimport, v1 as o1, v2 as o2

pytestmark = pytest.mark.skipif(not refimport.available(), reason="/root/reference not present")


def test_v2_init_forward_and_steps_bit_exact():
    gan, c = refimport.build_v2(seed=0)
    cfg = o2.V2Config(batch_size=c.batch_size)
    p = o2.init_vitgan(cfg, seed=0)
    sd = gan.state_dict()
    assert set(sd) == set(p) and all(torch.equal(sd[k], p[k]) for k in sd)
    (real, noise), = harness.synthetic_batches_v2(cfg, 4, 1, seed=5)
    with torch.no_grad():
        assert torch.equal(gan.discriminator(real), o2.vit_discriminator(p, "discriminator.", real, cfg))
        assert torch.equal(gan.generator(noise), o2.vit_generator(p, "generator.", noise, cfg))
    go = torch.optim.AdamW(gan.generator.parameters(), lr=5e-4, weight_decay=1e-3)
    do = torch.optim.AdamW(gan.discriminator.parameters(), lr=5e-4, weight_decay=1e-3)
    orc = harness.OracleV2(cfg, seed=0)
    for real, noise in harness.synthetic_batches_v2(cfg, 4, 3):
        a = harness.gan_step(gan.generator, gan.discriminator, go, do, real, noise, "ce")
        b = orc.step(real, noise)
        assert all(torch.equal(u, v) for u, v in zip(a, b))


@pytest.mark.parametrize("image_size", [32, 64])
def test_v1_init_forward_and_steps(image_size):
    G, D = refimport.build_v1(image_size, seed=0)
    cfg = o1.V1Config(image_size=image_size)
    orc = harness.OracleV1(cfg, seed=0)
    for k, v in G.state_dict().items():
        assert torch.equal(v, orc.p["generator." + k]), k
    for k, v in D.state_dict().items():
        assert torch.equal(v, orc.p["discriminator." + k]), k
    assert cfg.number_of_tokens == D.patch_encoder.number_of_tokens and cfg.stride == D.patch_encoder.stride
    (real, z), = harness.synthetic_batches_v1(cfg, 2, 1, seed=5)
    with torch.no_grad():
        assert torch.equal(G(z), orc.generator(z))
        assert harness.rel_err(orc.discriminator(real), D(real)) < 1e-6
    go = torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.5, 0.999))
    do = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.5, 0.999))
    for real, z in harness.synthetic_batches_v1(cfg, 2, 2):
        a = harness.gan_step(G, D, go, do, real, z, "bce")
        b = orc.step(real, z)
        assert all(harness.rel_err(v, u) < 1e-5 for u, v in zip(a, b))
