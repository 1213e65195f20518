"""Oracle (CPU, torch) restatement of the reference's v2 ViT path.  Test infrastructure.

Functional style: every function takes ``p`` -- a flat ``dict[str, Tensor]`` keyed
exactly like the reference ``state_dict()`` -- plus a key ``prefix``.  Gradients come
from torch autograd over these forwards, which is also what the reference relies on.

Reference: /root/reference/src/v2/modules.py (line numbers cited per function).
"""
from __future__ import annotations

import dataclasses
import math

import torch
import torch.nn.functional as F


@dataclasses.dataclass
class V2Config:
    """Mirror of the reference pydantic ``Config`` (src/v2/utils.py:25-40), same field names."""

    attention_heads_count: int = 4
    batch_size: int = 64  # sizes the G head Linear(classes_count -> batch_size), modules.py:361-364
    classes_count: int = 10
    discriminator_learning_rate: float = 5e-4
    dropout_rate: float = 0.1
    embeddings_dimension: int = 128
    generator_learning_rate: float = 5e-4
    image_size: int = 32
    input_channels: int = 3
    mlp_ratio: int = 2
    patch_size: int = 4
    transformer_blocks_count: int = 6

    @property
    def n_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    @property
    def seq_len(self) -> int:
        return self.n_patches + 1


# ------------------------------------------------------------------------------------------
# forward blocks
# ------------------------------------------------------------------------------------------

def embed_layer(p, pre, x, patch_size):
    """EmbedLayer.forward, modules.py:82-100 (dropout omitted: parity runs use p=0)."""
    tok = F.conv2d(x, p[pre + "conv1.weight"], p[pre + "conv1.bias"], stride=patch_size)  # :84
    b, e = tok.shape[0], tok.shape[1]
    tok = tok.reshape(b, e, -1).permute(0, 2, 1)                                           # :87-92
    tok = tok + p[pre + "pos_embedding"]                                                   # :93-95 (CLS gets no pos)
    cls = torch.repeat_interleave(p[pre + "cls_token"], b, 0)                              # :96-98
    return torch.cat((cls, tok), dim=1)


def self_attention(p, pre, x, n_heads):
    """SelfAttention.forward, modules.py:123-162."""
    b, s, e = x.shape
    d = e // n_heads

    def heads(name):
        y = F.linear(x, p[pre + name + ".weight"], p[pre + name + ".bias"])                # :128-139
        return y.reshape(b, s, n_heads, d).permute(0, 2, 1, 3)

    q, k, v = heads("queries"), heads("keys"), heads("values")
    att = torch.matmul(q, k.permute(0, 1, 3, 2))                                           # :142-145
    att = att / (float(d) ** 0.5)                                                          # :147-149 (scale AFTER QK^T)
    att = torch.softmax(att, dim=-1)                                                       # :151
    o = torch.matmul(att, v)                                                               # :153-155
    o = o.permute(0, 2, 1, 3).reshape(b, s, e)                                             # :158-159
    return F.linear(o, p[pre + "out_projection.weight"], p[pre + "out_projection.bias"])   # :161


def encoder(p, pre, x, n_heads):
    """Encoder.forward, modules.py:178-183 (pre-LN, eps 1e-5, exact-erf GELU)."""
    e = x.shape[-1]
    h = F.layer_norm(x, (e,), p[pre + "norm1.weight"], p[pre + "norm1.bias"], 1e-5)
    x = x + self_attention(p, pre + "attention.", h, n_heads)                              # :179
    h = F.layer_norm(x, (e,), p[pre + "norm2.weight"], p[pre + "norm2.bias"], 1e-5)
    h = F.gelu(F.linear(h, p[pre + "fc1.weight"], p[pre + "fc1.bias"]))                    # :173-174,181
    x = x + F.linear(h, p[pre + "fc2.weight"], p[pre + "fc2.bias"])                        # :180-182
    return x


def classifier(p, pre, x):
    """Classifier.forward, modules.py:194-199."""
    c = x[:, 0, :]
    c = torch.tanh(F.linear(c, p[pre + "fc1.weight"], p[pre + "fc1.bias"]))
    return F.linear(c, p[pre + "fc2.weight"], p[pre + "fc2.bias"])


def vision_transformer(p, pre, x, cfg: V2Config):
    """VisionTransformer.forward, modules.py:232-238."""
    x = embed_layer(p, pre + "embedding.", x, cfg.patch_size)
    for i in range(cfg.transformer_blocks_count):
        x = encoder(p, f"{pre}encoder.{i}.", x, cfg.attention_heads_count)
    e = x.shape[-1]
    x = F.layer_norm(x, (e,), p[pre + "norm.weight"], p[pre + "norm.bias"], 1e-5)          # :236
    return classifier(p, pre + "classifier.", x)


def vit_discriminator(p, pre, x, cfg: V2Config):
    """ViTDiscriminator.forward, modules.py:393-395."""
    return vision_transformer(p, pre + "vit.", x, cfg)


def vit_generator(p, pre, x, cfg: V2Config):
    """ViTGenerator.forward, modules.py:368-372 ("unpatchify" = Linear + view)."""
    y = vision_transformer(p, pre + "vit.", x, cfg)
    y = F.linear(y, p[pre + "linear.weight"], p[pre + "linear.bias"])
    return y.view(-1, cfg.input_channels, cfg.image_size, cfg.image_size)


# ------------------------------------------------------------------------------------------
# random init: reproduces the RNG consumption order of the reference constructors, so the
# same torch.manual_seed gives bit-identical weights (checked in tests/test_oracle_vs_reference.py)
# ------------------------------------------------------------------------------------------

def _default_linear_init(p, name, out_f, in_f, fan_in=None, wshape=None):
    """nn.Linear / nn.Conv2d.reset_parameters: kaiming_uniform(a=sqrt5) == U(+-1/sqrt(fan_in)) for W and b."""
    fan_in = in_f if fan_in is None else fan_in
    bound = 1.0 / math.sqrt(fan_in)
    # kaiming_uniform_(a=sqrt(5)): gain = sqrt(2/(1+a^2)); std = gain/sqrt(fan_in); bound = sqrt(3)*std
    a = math.sqrt(5)
    bound_w = math.sqrt(3.0) * (math.sqrt(2.0 / (1 + a ** 2)) / math.sqrt(fan_in))
    p[name + ".weight"] = torch.empty(wshape or (out_f, in_f)).uniform_(-bound_w, bound_w)
    p[name + ".bias"] = torch.empty(out_f).uniform_(-bound, bound)


def _vit_reinit(p, name):
    """vit_init_weights for Conv2d/Linear, modules.py:242-245."""
    torch.nn.init.trunc_normal_(p[name + ".weight"], mean=0.0, std=0.02)
    p[name + ".bias"].zero_()


def init_vision_transformer(p, pre, cfg: V2Config):
    """VisionTransformer.__init__ + self.apply(vit_init_weights), modules.py:202-230,241-253."""
    e, c, ps, m = cfg.embeddings_dimension, cfg.input_channels, cfg.patch_size, cfg.mlp_ratio
    lin = []  # (name) in construction order == nn.Module.apply (children-first) order
    # ---- construction (default inits consume the RNG first) ----
    _default_linear_init(p, pre + "embedding.conv1", e, None, fan_in=c * ps * ps, wshape=(e, c, ps, ps))
    p[pre + "embedding.pos_embedding"] = torch.zeros(1, cfg.n_patches, e)
    p[pre + "embedding.cls_token"] = torch.zeros(1, 1, e)
    lin.append(pre + "embedding.conv1")
    for i in range(cfg.transformer_blocks_count):
        b = f"{pre}encoder.{i}."
        for nm in ("norm1", "norm2"):
            p[b + nm + ".weight"] = torch.ones(e)
            p[b + nm + ".bias"] = torch.zeros(e)
        for nm in ("queries", "keys", "values", "out_projection"):
            _default_linear_init(p, b + "attention." + nm, e, e)
            lin.append(b + "attention." + nm)
        _default_linear_init(p, b + "fc1", e * m, e)
        _default_linear_init(p, b + "fc2", e, e * m)
        lin += [b + "fc1", b + "fc2"]
    p[pre + "norm.weight"] = torch.ones(e)
    p[pre + "norm.bias"] = torch.zeros(e)
    _default_linear_init(p, pre + "classifier.fc1", e, e)
    _default_linear_init(p, pre + "classifier.fc2", cfg.classes_count, e)
    lin += [pre + "classifier.fc1", pre + "classifier.fc2"]
    # ---- self.apply(vit_init_weights): children first, registration order ----
    _vit_reinit(p, lin[0])                                                     # conv1
    torch.nn.init.trunc_normal_(p[pre + "embedding.cls_token"], mean=0.0, std=0.02)   # :251 (cls first)
    torch.nn.init.trunc_normal_(p[pre + "embedding.pos_embedding"], mean=0.0, std=0.02)
    for name in lin[1:]:
        _vit_reinit(p, name)


def init_vitgan(cfg: V2Config, seed: int | None = 0, dtype=torch.float32):
    """ViTGAN.__init__, modules.py:398-405: generator (ViT + head Linear) then discriminator."""
    if seed is not None:
        torch.manual_seed(seed)
    p: dict[str, torch.Tensor] = {}
    init_vision_transformer(p, "generator.vit.", cfg)
    _default_linear_init(p, "generator.linear", cfg.batch_size, cfg.classes_count)   # :361-364 (keeps default init)
    init_vision_transformer(p, "discriminator.vit.", cfg)
    return {k: v.to(dtype) for k, v in p.items()}


def flops_per_image_fwd(cfg: V2Config, generator: bool) -> float:
    """Algorithmic forward FLOPs per image (SURVEY.md section 8d formula; multiply-add = 2, unpadded)."""
    e, n, s, l, m = (cfg.embeddings_dimension, cfg.n_patches, cfg.seq_len,
                     cfg.transformer_blocks_count, cfg.mlp_ratio)
    c, ps = cfg.input_channels, cfg.patch_size
    f = 2 * n * e * c * ps * ps + l * (2 * s * e * e * (4 + 2 * m) + 4 * s * s * e) + 2 * e * e + 2 * e * cfg.classes_count
    if generator:
        f += 2 * cfg.classes_count * c * cfg.image_size ** 2
    return float(f)


def convert_to_uint8(images):
    """utils.convert_to_uint8 (src/v2/utils.py:194-196): de-normalise generator output to bytes."""
    return (images * 127.5 + 127.5).clamp(0, 255).to(torch.uint8)
