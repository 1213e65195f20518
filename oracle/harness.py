"""Oracle harness: synthetic data + the G+D training step, restated on the CPU.  Test infrastructure.

Restates the active step of the reference loops:
  v2  src/v2/training.py:177-211 (AdamW lr 5e-4 wd 1e-3 :150-157, CrossEntropyLoss :159)
  v1  src/v1/gan.py:222-252      (Adam lr 2e-4 betas (0.5,0.999) :301-328, BCELoss :16-20)
with the external shims of SURVEY.md section 3.5: class-index CE targets (Q2), dropout p=0 (Q11),
explicit seeds and CPU-generated inputs (Q12).

The same step functions drive BOTH the oracle parameters and the CUDA-backed modules (anything
exposing ``generator(x)`` / ``discriminator(x)`` callables and optimizers), so that parity of loss
curves is well defined.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import v1 as o1
from . import v2 as o2


# ------------------------------------------------------------------------------------------
# synthetic data: always generated on the CPU in fp32
# ------------------------------------------------------------------------------------------

def synthetic_batches_v2(cfg: o2.V2Config, batch: int, steps: int, seed: int = 1234):
    """[(real, noise_for_G)] * steps; real in [-1,1] like Normalize(0.5,0.5) CIFAR, noise ~ N(0,1) (training.py:35-42)."""
    g = torch.Generator("cpu").manual_seed(seed)
    shape = (batch, cfg.input_channels, cfg.image_size, cfg.image_size)
    return [(torch.rand(shape, generator=g) * 2 - 1, torch.randn(shape, generator=g)) for _ in range(steps)]


def synthetic_batches_v1(cfg: o1.V1Config, batch: int, steps: int, seed: int = 1234):
    """[(real, z)] * steps; z ~ N(0,1) of size lattent_space_size (gan.py:231-232)."""
    g = torch.Generator("cpu").manual_seed(seed)
    shape = (batch, cfg.number_of_channels, cfg.image_size, cfg.image_size)
    return [(torch.rand(shape, generator=g) * 2 - 1, torch.randn(batch, cfg.lattent_space_size, generator=g))
            for _ in range(steps)]


# ------------------------------------------------------------------------------------------
# the step, written once against "anything callable"
# ------------------------------------------------------------------------------------------

def gan_step(gen_fwd, disc_fwd, gen_opt, disc_opt, real, noise, loss_kind: str):
    """One G+D iteration: 3 D forwards/backwards, 1 G forward/backward, 2 optimizer steps.

    loss_kind 'ce'  : v2, training.py:177-211 with class-index targets (label 1 = real, 0 = fake)
    loss_kind 'bce' : v1, gan.py:222-252 with float targets of shape (B,1)
    Returns (loss_real, loss_fake, loss_g) as detached tensors (no host sync here).
    """
    b = real.shape[0]
    dev = real.device
    if loss_kind == "ce":
        ones = torch.ones(b, dtype=torch.long, device=dev)
        zeros = torch.zeros(b, dtype=torch.long, device=dev)
        crit = lambda out, tgt: F.cross_entropy(out.float(), tgt)
    else:
        ones = torch.ones(b, 1, device=dev)
        zeros = torch.zeros(b, 1, device=dev)
        crit = lambda out, tgt: F.binary_cross_entropy(out.float(), tgt)

    disc_opt.zero_grad(set_to_none=True)                      # gan.discriminator.zero_grad()   :177 / :222
    loss_real = crit(disc_fwd(real), ones)                    # :182-183 / :226-227
    loss_real.backward()                                      # :184 / :228
    fake = gen_fwd(noise)                                     # :187 / :233
    loss_fake = crit(disc_fwd(fake.detach()), zeros)          # :190-193 / :235-238
    loss_fake.backward()                                      # :194 / :239
    disc_opt.step()                                           # :197 / :242
    gen_opt.zero_grad(set_to_none=True)                       # :199 / :245
    loss_g = crit(disc_fwd(fake), ones)                       # :204-209 / :247-250
    loss_g.backward()                                         # :210 / :251
    gen_opt.step()                                            # :211 / :252
    return loss_real.detach(), loss_fake.detach(), loss_g.detach()


# ------------------------------------------------------------------------------------------
# oracle-side model containers (leaf tensors + functional forwards)
# ------------------------------------------------------------------------------------------

class OracleV2:
    """Oracle parameters + optimizers for the v2 GAN."""

    def __init__(self, cfg: o2.V2Config, seed: int = 0, dtype=torch.float32, params=None):
        self.cfg = cfg
        src = params if params is not None else o2.init_vitgan(cfg, seed, dtype)
        self.p = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in src.items()}
        g = [v for k, v in self.p.items() if k.startswith("generator.")]
        d = [v for k, v in self.p.items() if k.startswith("discriminator.")]
        self.gen_opt = torch.optim.AdamW(g, lr=cfg.generator_learning_rate, weight_decay=1e-3)      # training.py:150-152
        self.disc_opt = torch.optim.AdamW(d, lr=cfg.discriminator_learning_rate, weight_decay=1e-3)  # :153-157

    def generator(self, x):
        return o2.vit_generator(self.p, "generator.", x, self.cfg)

    def discriminator(self, x):
        return o2.vit_discriminator(self.p, "discriminator.", x, self.cfg)

    def step(self, real, noise):
        return gan_step(self.generator, self.discriminator, self.gen_opt, self.disc_opt, real, noise, "ce")


class OracleV1:
    """Oracle parameters + optimizers for the v1 GAN.

    Reproduces SURVEY Q4: the reference re-wraps D's q/k/v weights in fresh nn.Parameters on every
    forward, so the optimizer (built from the original objects) never updates them.  Here they are
    simply left out of the D optimizer; after construction sigma_init/sigma_now == 1 exactly.
    """

    def __init__(self, cfg: o1.V1Config, seed: int = 0, dtype=torch.float32, params=None):
        self.cfg = cfg
        if params is None:
            gp = o1.init_generator(cfg, "generator.", seed, dtype)
            dp = o1.init_discriminator(cfg, "discriminator.", None, dtype)   # same RNG stream, G first (vitgan.py:9-13)
            params = {**gp, **dp}
        self.p = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in params.items()}
        self.spectra = o1.initial_spectra(self.p, "discriminator.", cfg)
        g = [v for k, v in self.p.items() if k.startswith("generator.")]
        frozen = (".q.weight", ".k.weight", ".v.weight")
        d = [v for k, v in self.p.items() if k.startswith("discriminator.") and not k.endswith(frozen)]
        self.gen_opt = torch.optim.Adam(g, lr=2e-4, betas=(0.5, 0.999))      # gan.py:316-321
        self.disc_opt = torch.optim.Adam(d, lr=2e-4, betas=(0.5, 0.999))     # gan.py:322-327

    def generator(self, z):
        return o1.generator(self.p, "generator.", z, self.cfg)

    def discriminator(self, x):
        return o1.discriminator(self.p, "discriminator.", x, self.cfg, self.spectra)

    def step(self, real, z):
        return gan_step(self.generator, self.discriminator, self.gen_opt, self.disc_opt, real, z, "bce")


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b|  -- the per-tensor relative error BASELINE.json's tolerances refer to."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = b.abs().max().item()
    return (a - b).abs().max().item() / (den if den > 0 else 1.0)
