"""Host-side mirror of the reference's v2 ViT-GAN modules (src/v2/modules.py:67-410).

Same class names, constructor arguments, parameter names / shapes / init and ``state_dict`` keys as the
reference, so checkpoints and the training-loop call sites (src/v2/training.py:177-211) work unchanged;
only the body of each ``forward`` differs: it is one ``torch.autograd.Function`` backed by the sm_100a
kernels of libvitgan_b200.  ``patch.patch_v2(gan)`` applies the same forwards to instances of the real
reference classes.  Dropout is not implemented inside the fused blocks: a block in training mode with p > 0 raises
(functional.check_dropout) unless the caller opts into the p = 0 parity protocol (set_dropout_policy('off')).
"""
from __future__ import annotations

import dataclasses

import torch
import torch.nn as nn

from . import functional as Fn
from . import lib as L
from . import ops


@dataclasses.dataclass
class Config:
    """Field-for-field mirror of the reference pydantic Config (src/v2/utils.py:25-40)."""

    attention_heads_count: int = 4
    batch_size: int = 64
    classes_count: int = 10
    discriminator_learning_rate: float = 5e-4
    dropout_rate: float = 0.1
    embeddings_dimension: int = 128
    epochs: int = 500
    generator_learning_rate: float = 5e-4
    image_size: int = 32
    input_channels: int = 3
    mlp_ratio: int = 2
    optimizer_beta1: float = 0.5
    optimizer_beta2: float = 0.999
    patch_size: int = 4
    transformer_blocks_count: int = 6


# ---- forward bodies, written against duck-typed modules so patch.py can bind them to reference instances

def embed_forward(self, x):
    """EmbedLayer.forward (modules.py:82-100)."""
    Fn.check_dropout(self)
    return Fn.EmbedV2Fn.apply(x, self.conv1.weight, self.conv1.bias, self.pos_embedding, self.cls_token,
                              self.conv1.kernel_size[0])


def self_attention_forward(self, x):
    """SelfAttention.forward (modules.py:123-162)."""
    return Fn.SelfAttentionFn.apply(x, self.n_attention_heads, self.queries.weight, self.queries.bias, self.keys.weight,
                                    self.keys.bias, self.values.weight, self.values.bias, self.out_projection.weight,
                                    self.out_projection.bias)


def encoder_forward_chained(self, x, ln1=None, next_norm=None):
    """Encoder.forward (modules.py:178-183) as one autograd Function.  ln1 = (xn, mean, rstd) of this block's first LayerNorm if
    the previous block's fc2 epilogue already produced it; next_norm = the next block's norm1 module, whose LayerNorm is then
    computed by this block's fc2 epilogue.  -> (y, ln1-of-next-block | None)."""
    Fn.check_dropout(self)
    a = self.attention
    c = ln1 if ln1 is not None else (None, None, None)
    ng, nb = (next_norm.weight, next_norm.bias) if next_norm is not None else (None, None)
    out = Fn.EncoderFn.apply(x, c[0], c[1], c[2], ng, nb, a.n_attention_heads, self.norm1.weight, self.norm1.bias, a.queries.weight,
                             a.queries.bias, a.keys.weight, a.keys.bias, a.values.weight, a.values.bias, a.out_projection.weight,
                             a.out_projection.bias, self.norm2.weight, self.norm2.bias, self.fc1.weight, self.fc1.bias,
                             self.fc2.weight, self.fc2.bias)
    return out[0], (out[1:] if out[1] is not None else None)


def encoder_forward(self, x):
    """Encoder.forward (modules.py:178-183)."""
    return encoder_forward_chained(self, x)[0]


def _chainable(block, nxt):
    return (type(nxt).__name__ == type(block).__name__ and getattr(nxt.norm1, "eps", None) == 1e-5 and getattr(block.norm2, "eps", None) == 1e-5
            and tuple(nxt.norm1.normalized_shape) == tuple(block.norm1.normalized_shape))


def _head(fc1, fc2, c):
    h = Fn.LinearFn.apply(c, fc1.weight, fc1.bias, L.ACT_TANH, 0.0, "act", None)
    return Fn.LinearFn.apply(h, fc2.weight, fc2.bias, L.ACT_NONE, 0.0, "act", torch.float32)


def classifier_forward(self, x):
    """Classifier.forward (modules.py:194-199): CLS row -> fc1 -> tanh -> fc2 (logits in fp32)."""
    return _head(self.fc1, self.fc2, Fn.ClsRowFn.apply(x))


def vit_forward(self, x):
    """VisionTransformer.forward (modules.py:232-238).  The final LayerNorm is applied to the CLS rows only:
    the reference normalises all S rows and then discards all but row 0 (modules.py:195,236)."""
    x = embed_forward(self.embedding, x)
    blocks, ln1 = list(self.encoder), None
    for i, block in enumerate(blocks):       # with the fused-LayerNorm epilogue enabled, block i's fc2 GEMM also normalises for block i+1
        nxt = blocks[i + 1].norm1 if Fn.fused_layernorm_epilogue() and i + 1 < len(blocks) and _chainable(block, blocks[i + 1]) else None
        x, ln1 = encoder_forward_chained(block, x, ln1, nxt)
    c = Fn.ClsRowFn.apply(x)
    c = Fn.LayerNormFn.apply(c, self.norm.weight, self.norm.bias, self.norm.eps)
    return _head(self.classifier.fc1, self.classifier.fc2, c)


def generator_forward(self, x):
    """ViTGenerator.forward (modules.py:368-372): ViT -> Linear(classes -> C*I*I) -> view ("unpatchify")."""
    y = vit_forward(self.vit, x)
    y = Fn.LinearFn.apply(y, self.linear.weight, self.linear.bias, L.ACT_NONE, 0.0, "fp32", torch.float32)
    return y.view(-1, self.input_channels, self.image_size, self.image_size)


def discriminator_forward(self, x):
    """ViTDiscriminator.forward (modules.py:393-395)."""
    return vit_forward(self.vit, x)


# ---- module classes (mirror) -------------------------------------------------------------------

class EmbedLayer(nn.Module):
    def __init__(self, n_channels, embed_dim, image_size, patch_size, dropout=0.0):
        super().__init__()
        self.conv1 = nn.Conv2d(n_channels, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.pos_embedding = nn.Parameter(torch.zeros(1, (image_size // patch_size) ** 2, embed_dim), requires_grad=True)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim), requires_grad=True)
        self.dropout = nn.Dropout(dropout)

    forward = embed_forward


class SelfAttention(nn.Module):
    def __init__(self, embed_dim, n_attention_heads):
        super().__init__()
        self.embed_dim = embed_dim
        self.n_attention_heads = n_attention_heads
        self.head_embed_dim = embed_dim // n_attention_heads
        self.queries = nn.Linear(self.embed_dim, self.head_embed_dim * self.n_attention_heads)
        self.keys = nn.Linear(self.embed_dim, self.head_embed_dim * self.n_attention_heads)
        self.values = nn.Linear(self.embed_dim, self.head_embed_dim * self.n_attention_heads)
        self.out_projection = nn.Linear(self.head_embed_dim * self.n_attention_heads, self.embed_dim)

    forward = self_attention_forward


class Encoder(nn.Module):
    def __init__(self, embed_dim, n_attention_heads, forward_mul, dropout=0.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(embed_dim)
        self.attention = SelfAttention(embed_dim, n_attention_heads)
        self.dropout1 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(embed_dim)
        self.fc1 = nn.Linear(embed_dim, embed_dim * forward_mul)
        self.activation = nn.GELU()
        self.fc2 = nn.Linear(embed_dim * forward_mul, embed_dim)
        self.dropout2 = nn.Dropout(dropout)

    forward = encoder_forward


class Classifier(nn.Module):
    def __init__(self, embed_dim, n_classes):
        super().__init__()
        self.fc1 = nn.Linear(embed_dim, embed_dim)
        self.activation = nn.Tanh()
        self.fc2 = nn.Linear(embed_dim, n_classes)

    forward = classifier_forward


def vit_init_weights(m):
    """modules.py:241-253 (type checks by class NAME so that patched reference instances also match)."""
    if isinstance(m, (nn.Conv2d, nn.Linear)):
        nn.init.trunc_normal_(m.weight, mean=0.0, std=0.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.weight, 1)
        nn.init.constant_(m.bias, 0)
    elif isinstance(m, EmbedLayer):
        nn.init.trunc_normal_(m.cls_token, mean=0.0, std=0.02)
        nn.init.trunc_normal_(m.pos_embedding, mean=0.0, std=0.02)


class VisionTransformer(nn.Module):
    def __init__(self, n_channels, embed_dim, n_layers, n_attention_heads, forward_mul, image_size, patch_size, n_classes,
                 dropout=0.1):
        super().__init__()
        self.embedding = EmbedLayer(n_channels, embed_dim, image_size, patch_size, dropout=dropout)
        self.encoder = nn.ModuleList([Encoder(embed_dim, n_attention_heads, forward_mul, dropout=dropout) for _ in range(n_layers)])
        self.norm = nn.LayerNorm(embed_dim)
        self.classifier = Classifier(embed_dim, n_classes)
        self.apply(vit_init_weights)

    forward = vit_forward


def _vit_from_config(config):
    return VisionTransformer(n_channels=config.input_channels, embed_dim=config.embeddings_dimension,
                             n_layers=config.transformer_blocks_count, n_attention_heads=config.attention_heads_count,
                             forward_mul=config.mlp_ratio, image_size=config.image_size, patch_size=config.patch_size,
                             n_classes=config.classes_count, dropout=config.dropout_rate)


class ViTGenerator(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.vit = _vit_from_config(config)
        self.linear = nn.Linear(config.classes_count, config.batch_size)
        self.image_size = config.image_size
        self.input_channels = config.input_channels

    forward = generator_forward


class ViTDiscriminator(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.vit = _vit_from_config(config)

    forward = discriminator_forward


class ViTGAN(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.generator = ViTGenerator(config)
        self.discriminator = ViTDiscriminator(config)

    def forward(self, z):
        generated_images = self.generator(z)
        discriminator_output = self.discriminator(generated_images)
        return generated_images, discriminator_output


@torch.no_grad()
def sample_uint8(generator, noise):
    """Batched sampling path: the reference re-runs the generator per image and de-normalises on the
    host (src/v2/generation.py:47-56, utils.convert_to_uint8 utils.py:194-196); here one batched forward through the CUDA
    kernels and a fused `* 127.5 + 127.5 -> clamp -> uint8` kernel.  noise: (B, C, I, I) -> uint8 (B, C, I, I)."""
    was_training = generator.training
    generator.eval()
    try:
        img = generator(noise)
    finally:
        generator.train(was_training)
    return ops.denorm_u8(img.float() if img.dtype != torch.float32 else img)
