"""Oracle (CPU, torch) restatement of the reference's v1 ViTGAN path.  Test infrastructure.

Generator  = mapping Linear -> 4 x TransformerSLN -> SLN -> 2 x SIREN -> view
Discriminator = overlapping-patch encoder -> 4 x Transformer (L2 attention, spectral rescale)
                -> Linear(432 -> 1) -> sigmoid

``p`` is a flat dict keyed like the reference ``Generator().state_dict()`` /
``Discriminator().state_dict()`` (with an optional prefix).  Reference files:
/root/reference/src/v1/{attention,transformer,spectral_layer_norm,muilti_layer_perceptron,
patch_encoder,siren,generator,discriminatorViT}.py -- cited per function.

Shims the reference needs to run at all are baked in as config values:
``projection_output_size = 432``, ``Transformer.input_features = 432``, D head out = 1.
"""
from __future__ import annotations

import dataclasses
import math

import torch
import torch.nn.functional as F


@dataclasses.dataclass
class V1Config:
    """Defaults of src/v1/config.py:20-70 that reach the hot path."""

    image_size: int = 32
    number_of_channels: int = 3
    lattent_space_size: int = 1024            # (sic) config.py:66
    feature_hidden_size: int = 384            # GeneratorParameters
    g_layers: int = 4
    output_hidden_dimension: int = 768
    d_layers: int = 4
    number_of_heads: int = 4
    patch_size: int = 8                       # EncoderParameters
    overlap: int = 2
    omega_0: int = 30                         # siren.py:12
    d_out_features: int = 1                   # shim Q3

    @property
    def window(self):
        return self.patch_size + 2 * self.overlap

    @property
    def token_size(self):                     # patch_encoder.py:17-19
        return self.number_of_channels * self.window ** 2

    @property
    def stride(self):                         # patch_encoder.py:20-22
        return (self.image_size - self.patch_size - 2 * self.overlap) // self.patch_size + 1

    @property
    def number_of_tokens(self):               # patch_encoder.py:23-27
        return ((self.image_size - (self.window - 1) - 1) // self.stride + 1) ** 2

    @property
    def d_features(self):                     # discriminatorViT.py:24 (= token_size = 432)
        return self.token_size


# ------------------------------------------------------------------------------------------
# shared blocks
# ------------------------------------------------------------------------------------------

def sigma_max(w):
    """max singular value as the reference gets it: full SVD then python max(), attention.py:39,54-58."""
    return torch.linalg.svdvals(w).max()


def attention_head(p, pre, x, scale, lp, init_spectrum=None):
    """Attention.forward, attention.py:43-52.

    lp == 1: dot-product scores (:69-70); lp == 2: +||q_i - k_j||_2 via torch.cdist (:66-67).
    ``init_spectrum`` (sq0, sk0, sv0) enables _weight_spectral_rescale (:60-64):
    W <- sigma_init / sigma_now * W before the projections.
    """
    wq, wk, wv = p[pre + "q.weight"], p[pre + "k.weight"], p[pre + "v.weight"]
    if init_spectrum is not None:
        # the reference re-wraps the product in a fresh leaf nn.Parameter (:62-64): no gradient flows
        # through sigma_max, and the leaf's grad is dL/dW_eff -> the factor is a detached constant here
        with torch.no_grad():
            fq = init_spectrum[0] / sigma_max(wq)
            fk = init_spectrum[1] / sigma_max(wk)
            fv = init_spectrum[2] / sigma_max(wv)
        wq, wk, wv = fq * wq, fk * wk, fv * wv
    q, k, v = F.linear(x, wq), F.linear(x, wk), F.linear(x, wv)                 # :46-48 (bias=False)
    if lp == 2:
        att = torch.cdist(q, k, p=2)                                            # :67
    else:
        att = torch.einsum("...id,...jd->...ij", q, k)                          # :70
    return torch.softmax(att / (scale ** (1 / 2)), dim=-1) @ v                  # :51  (scale = H*d, Q7)


def multi_head_self_attention(p, pre, x, n_heads, lp, spectra=None):
    """MultiHeadSelfAttention.forward, attention.py:97-103 (python loop over heads, cat, out Linear)."""
    d = p[pre + "attention_heads.0.q.weight"].shape[0]
    scale = n_heads * d                                                          # :82,:90
    outs = []
    for h in range(n_heads):
        hp = f"{pre}attention_heads.{h}."
        outs.append(attention_head(p, hp, x, scale, lp, None if spectra is None else spectra[hp]))
    o = torch.cat(outs, dim=-1)
    return F.linear(o, p[pre + "output_linear.weight"], p[pre + "output_linear.bias"])


def mlp_single(p, pre, x):
    """MLP.forward with the default ``layers=[]``: one Linear, no activation (muilti_layer_perceptron.py:37-42, Q8)."""
    return F.linear(x, p[pre + "model.0.0.weight"], p[pre + "model.0.0.bias"])


def sln(p, pre, h, w):
    """SLN.forward, spectral_layer_norm.py:19-20: gamma*w*LN(h) + beta*w, scalar gamma/beta (Q9)."""
    f = h.shape[-1]
    ln = F.layer_norm(h, (f,), p[pre + "layer_norm.weight"], p[pre + "layer_norm.bias"], 1e-5)
    return p[pre + "gamma"] * w * ln + p[pre + "beta"] * w


def siren(p, pre, x, omega_0):
    """SIREN.forward, siren.py:44-45."""
    return torch.sin(omega_0 * F.linear(x, p[pre + "linear.weight"], p[pre + "linear.bias"]))


# ------------------------------------------------------------------------------------------
# generator
# ------------------------------------------------------------------------------------------

def transformer_sln(p, pre, h, w, n_heads):
    """TransformerSLN.forward, transformer.py:85-88 -> (w, h')."""
    a = multi_head_self_attention(p, pre + "msha.", sln(p, pre + "layer_norm_1.", h, w), n_heads, lp=1)
    htmp = a + h                                                                 # :86 (h (S,F) broadcasts on layer 0)
    hf = mlp_single(p, pre + "mlp.", sln(p, pre + "layer_norm_2.", htmp, w)) + htmp   # :87
    return w, hf


def generator(p, pre, z, cfg: V1Config):
    """Generator.forward, generator.py:58-69."""
    i, f = cfg.image_size, cfg.feature_hidden_size
    w = mlp_single(p, pre + "mapping_mlp.", z).view(-1, i, f)                    # :59-61 (seq len = image_size)
    h = p[pre + "embedding"]                                                     # :62
    for l in range(cfg.g_layers):
        w, h = transformer_sln(p, f"{pre}transformer_layers.{l}.", h, w, cfg.number_of_heads)
    w = sln(p, pre + "sln.", h, w)                                               # :65
    y = siren(p, pre + "output_network.0.", w, cfg.omega_0)
    y = siren(p, pre + "output_network.1.", y, cfg.omega_0)                      # :66
    return y.view(z.shape[0], cfg.number_of_channels, i, i)                      # :66-68 (raw reinterpretation)


# ------------------------------------------------------------------------------------------
# discriminator
# ------------------------------------------------------------------------------------------

def get_tokens(images, cfg: V1Config):
    """PatchEncoder._get_tokens, patch_encoder.py:54-73.

    NOTE the reference flattens (C, n_h, n_w, win, win) WITHOUT a permute, so token t is the flat
    range [432 t, 432 t + 432) of that tensor ("scrambled" layout) -- reproduced as is.
    """
    pt = images.unfold(2, cfg.window, cfg.stride).unfold(3, cfg.window, cfg.stride)
    pt = pt.contiguous()
    return pt.view(pt.shape[0], pt.shape[2] * pt.shape[3], pt.shape[1] * pt.shape[4] * pt.shape[5])


def patch_encoder(p, pre, images, cfg: V1Config):
    """PatchEncoder.forward, patch_encoder.py:39-52 (dropout p=0.0)."""
    tok = F.linear(get_tokens(images, cfg), p[pre + "projection_matrix.weight"])          # :44 (no bias)
    cls = p[pre + "cls_token"].expand(tok.shape[0], 1, tok.shape[-1])                    # :45-47
    return torch.cat((cls, tok), dim=1) + p[pre + "positional_embedding"]                # :49-50 (CLS gets pos)


def transformer(p, pre, x, n_heads, spectra):
    """Transformer.forward, transformer.py:40-45 (D block: LN, L2-attention MSA, +res, LN, Linear, +res)."""
    f = x.shape[-1]
    x1 = F.layer_norm(x, (f,), p[pre + "layer_norm_1.weight"], p[pre + "layer_norm_1.bias"], 1e-5)
    x = x + multi_head_self_attention(p, pre + "msha.", x1, n_heads, lp=2, spectra=spectra)
    x2 = F.layer_norm(x, (f,), p[pre + "layer_norm_2.weight"], p[pre + "layer_norm_2.bias"], 1e-5)
    return x + mlp_single(p, pre + "mlp.", x2)


def discriminator(p, pre, images, cfg: V1Config, spectra):
    """Discriminator.forward, discriminatorViT.py:44-51."""
    t = patch_encoder(p, pre + "patch_encoder.", images, cfg)
    for l in range(cfg.d_layers):
        t = transformer(p, f"{pre}transformer_layers.{l}.", t, cfg.number_of_heads, spectra)
    return torch.sigmoid(mlp_single(p, pre + "mlp.", t[:, 0, :]))                         # :48-51


def initial_spectra(p, pre, cfg: V1Config):
    """Attention.__init__ spectral bookkeeping, attention.py:37-39: sigma_max of q/k/v at construction."""
    out = {}
    with torch.no_grad():
        for l in range(cfg.d_layers):
            for h in range(cfg.number_of_heads):
                hp = f"{pre}transformer_layers.{l}.msha.attention_heads.{h}."
                out[hp] = tuple(sigma_max(p[hp + n + ".weight"]).clone() for n in ("q", "k", "v"))
    return out


# ------------------------------------------------------------------------------------------
# random init in the reference's RNG consumption order
# ------------------------------------------------------------------------------------------

def _linear_init(p, name, out_f, in_f, bias=True):
    a = math.sqrt(5)
    bw = math.sqrt(3.0) * (math.sqrt(2.0 / (1 + a ** 2)) / math.sqrt(in_f))
    p[name + ".weight"] = torch.empty(out_f, in_f).uniform_(-bw, bw)
    if bias:
        b = 1.0 / math.sqrt(in_f)
        p[name + ".bias"] = torch.empty(out_f).uniform_(-b, b)


def _msha_init(p, pre, f, n_heads):
    d = f // n_heads
    for h in range(n_heads):
        for n in ("q", "k", "v"):
            _linear_init(p, f"{pre}attention_heads.{h}.{n}", d, f, bias=False)
    _linear_init(p, pre + "output_linear", f, n_heads * d)


def init_generator(cfg: V1Config, pre="", seed: int | None = 0, dtype=torch.float32):
    """Generator.__init__, generator.py:13-56 (+ TransformerSLN/SLN/SIREN constructors)."""
    if seed is not None:
        torch.manual_seed(seed)
    p = {}
    i, f = cfg.image_size, cfg.feature_hidden_size
    _linear_init(p, pre + "mapping_mlp.model.0.0", i * f, cfg.lattent_space_size)
    p[pre + "embedding"] = torch.randn(i, f)

    def sln_init(name):
        p[name + "layer_norm.weight"] = torch.ones(f)
        p[name + "layer_norm.bias"] = torch.zeros(f)
        p[name + "beta"] = torch.randn(1, 1, 1)            # spectral_layer_norm.py:16 (beta first)
        p[name + "gamma"] = torch.randn(1, 1, 1)

    for l in range(cfg.g_layers):
        b = f"{pre}transformer_layers.{l}."
        sln_init(b + "layer_norm_1.")
        sln_init(b + "layer_norm_2.")
        _msha_init(p, b + "msha.", f, cfg.number_of_heads)
        _linear_init(p, b + "mlp.model.0.0", f, f)
    sln_init(pre + "sln.")
    h = cfg.output_hidden_dimension
    _linear_init(p, pre + "output_network.0.linear", h, f)
    p[pre + "output_network.0.linear.weight"].uniform_(-1 / f, 1 / f)                      # siren.py:31-35
    _linear_init(p, pre + "output_network.1.linear", cfg.number_of_channels * i, h)
    bnd = float(__import__("numpy").sqrt(6 / h) / cfg.omega_0)                             # siren.py:37-42
    p[pre + "output_network.1.linear.weight"].uniform_(-bnd, bnd)
    return {k: v.to(dtype) for k, v in p.items()}


def init_discriminator(cfg: V1Config, pre="", seed: int | None = None, dtype=torch.float32):
    """Discriminator.__init__, discriminatorViT.py:17-42 (+ PatchEncoder/Transformer constructors)."""
    if seed is not None:
        torch.manual_seed(seed)
    p = {}
    f = cfg.d_features
    _linear_init(p, pre + "patch_encoder.projection_matrix", f, cfg.token_size, bias=False)
    p[pre + "patch_encoder.cls_token"] = torch.randn(1, 1, f)
    p[pre + "patch_encoder.positional_embedding"] = torch.randn(cfg.number_of_tokens + 1, f)
    for l in range(cfg.d_layers):
        b = f"{pre}transformer_layers.{l}."
        for nm in ("layer_norm_1", "layer_norm_2"):
            p[b + nm + ".weight"] = torch.ones(f)
            p[b + nm + ".bias"] = torch.zeros(f)
        _msha_init(p, b + "msha.", f, cfg.number_of_heads)
        _linear_init(p, b + "mlp.model.0.0", f, f)
    _linear_init(p, pre + "mlp.model.0.0", cfg.d_out_features, f)
    return {k: v.to(dtype) for k, v in p.items()}


def flops_per_image_fwd(cfg: V1Config, generator_net: bool) -> float:
    """SURVEY.md section 8d, v1 formulas (sigma_max excluded, cdist counted as one S^2 d contraction)."""
    if generator_net:
        i, f, s = cfg.image_size, cfg.feature_hidden_size, cfg.image_size
        return float(2 * cfg.lattent_space_size * i * f + cfg.g_layers * (2 * s * f * f * 5 + 4 * s * s * f)
                     + 2 * s * f * cfg.output_hidden_dimension + 2 * s * cfg.output_hidden_dimension * 3 * i)
    f, n = cfg.d_features, cfg.number_of_tokens
    s = n + 1
    return float(2 * n * f * f + cfg.d_layers * (2 * s * f * f * 5 + 4 * s * s * f) + 2 * f)
