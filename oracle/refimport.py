"""Import the REAL reference.  Test infrastructure.

In the build container the reference is imported from ``/root/reference``; on the GPU box (where that path does not exist) from
the git-ignored staged copy ``oracle/_ref`` made by ``oracle/stage_ref.py``.  Used by ``make_golden.py`` and
``tests/test_oracle_vs_reference.py`` to pin the oracle against the reference itself, by the GPU tests that patch instances of
the real reference classes, and by ``bench.py --impl reference``.
No reference file is edited; the shims are the external ones listed in SURVEY.md section 3.5/8c.
"""
from __future__ import annotations

import os
import sys
from unittest.mock import MagicMock

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _resolve_root() -> str:
    env = os.environ.get("VITGAN_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/src/v2"):
        return "/root/reference"
    return _STAGED


REFERENCE_ROOT = _resolve_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "v2"))


def _prepare():
    # packages the reference imports at module top but that do no hot-path arithmetic (absent here)
    for name in ("matplotlib", "matplotlib.pyplot", "torchmetrics", "torchmetrics.image",
                 "torchmetrics.image.fid", "ray", "ray.tune"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = MagicMock()
    os.environ.setdefault("SCRATCH", "/tmp/vitgan_scratch")      # src/v1/config.py:9,11 needs it
    os.makedirs(os.environ["SCRATCH"], exist_ok=True)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def zero_dropout(module):
    """parity runs set every nn.Dropout.p = 0 on both sides."""
    import torch.nn as nn
    for m in module.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    return module


def build_v2(seed=0, **cfg_overrides):
    """-> (reference ViTGAN instance, reference Config).  Shim Q1: batch_size = C*I*I unless given."""
    _prepare()
    import torch
    from src.v2 import modules
    from src.v2.utils import Config
    c = Config(**cfg_overrides)
    if "batch_size" not in cfg_overrides:
        c = Config(**{**cfg_overrides, "batch_size": c.input_channels * c.image_size ** 2})
    torch.manual_seed(seed)
    gan = modules.ViTGAN(c)
    return zero_dropout(gan), c


def v2_modules():
    _prepare()
    from src.v2 import modules
    return modules


def build_v1(image_size=32, seed=0):
    """-> (reference Generator, reference Discriminator) with shims Q3 and Q13 applied."""
    _prepare()
    import torch
    from src.v1.config import config
    config.image_size = image_size                                           # Q13: before construction
    config.discriminator_params.mapping_mlp_params.output_features = 1      # Q3
    from src.v1.patch_encoder import PatchEncoder
    from src.v1.transformer import Transformer
    PatchEncoder.projection_output_size = 432                                # Q3
    Transformer.input_features = 432                                         # Q3
    from src.v1.generator import Generator
    from src.v1.discriminatorViT import Discriminator
    torch.manual_seed(seed)
    g = Generator()
    d = Discriminator()
    return zero_dropout(g), zero_dropout(d)


def v1_modules():
    _prepare()
    import importlib
    names = ("attention", "transformer", "spectral_layer_norm", "muilti_layer_perceptron",
             "patch_encoder", "siren", "generator", "discriminatorViT", "config")
    return {n: importlib.import_module("src.v1." + n) for n in names}
