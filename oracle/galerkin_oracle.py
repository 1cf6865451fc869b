"""CPU oracle: a functional, pure-torch restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``galerkin_transformer_b200/`` may import
this file; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs
(``cpu_baseline`` / ``--impl reference``) use it, and only as the checker / baseline.

Every function takes a flat ``state_dict`` (``sd``) with the reference's parameter
names plus a ``prefix`` and recomputes what the reference ``nn.Module`` computes, in
whatever dtype/device the tensors are in (fp32 for the baseline, fp64 for tight
checks).  It is written against the behaviour of scaomath/galerkin-transformer
@ f9e7d6ab; each function cites the reference lines it follows
(paths relative to the reference root).

Pinning: the reference ships no golden vectors, so this oracle is
pinned against the reference modules themselves, imported in the build container by
``tests/golden/make_golden.py``; the resulting fixtures are committed under
``tests/golden/*.pt`` and ``tests/test_oracle_golden.py`` checks this file against
them (fp32, rel-L2 <= 2e-6 forward, <= 2e-5 gradients).

Dropout handling.  The reference passes attention weights through
``F.dropout(p_attn)`` with p=0.5, training=True *unconditionally*
(libs/layers.py:700-701, 730-731).  Here that is an explicit argument:
``attn_mask`` (a 0/1 keep-mask, scaled by 2 when applied), ``attn_dropout=True``
(draw a fresh mask) or neither (identity; what the parity tests use on both sides).
All ``nn.Dropout`` modules are treated as identity (p=0 / eval).
"""
import math

import torch
import torch.nn.functional as F

GALERKIN_TYPES = ("galerkin",)
FOURIER_TYPES = ("fourier", "integral", "local")


def _j(prefix, name):
    """Join state_dict key parts; an empty prefix means the module is the root."""
    return prefix + "." + name if prefix else name


def _lin(sd, prefix, x):
    """nn.Linear: x @ W^T + b."""
    b = sd.get(_j(prefix, "bias"))
    return F.linear(x, sd[_j(prefix, "weight")], b)


def _act(name, x):
    if name == "silu":
        return F.silu(x)
    if name == "gelu":
        return F.gelu(x)
    if name == "identity":
        return x
    return F.relu(x)


def _head_layernorm(sd, prefix, t, eps):
    """Per-head LayerNorm over the last d_k entries, separate affine per head.

    libs/layers.py:846-851 / 859-864: ``stack([norm_h(x[:, h]) for h in heads], 1)``.
    t: (B, H, n, d_k)."""
    n_head, d_k = t.shape[1], t.shape[-1]
    out = []
    for h in range(n_head):
        out.append(F.layer_norm(t[:, h], (d_k,), sd[_j(prefix, f"{h}.weight")],
                                sd[_j(prefix, f"{h}.bias")], eps))
    return torch.stack(out, dim=1)


def simple_attention(sd, prefix, query, key, value, pos=None, *, n_head,
                     attention_type="galerkin", norm=True, eps=1e-5, pos_dim=None,
                     attn_mask=None, attn_dropout=False):
    """SimpleAttention.forward, libs/layers.py:829-899.

    Returns (out (B, n, d_model), attn_weight).  ``attn_weight`` is the
    post-dropout (B,H,d,d) matrix for Galerkin and (B,H,n,n) for Fourier."""
    bsz, d_model = query.shape[0], query.shape[-1]
    d_k = d_model // n_head
    p = f"{prefix}." if prefix else ""
    q, k, v = [_lin(sd, f"{p}linears.{i}", x).view(bsz, -1, n_head, d_k).transpose(1, 2)
               for i, x in enumerate((query, key, value))]            # :837-839
    if norm:
        if attention_type in GALERKIN_TYPES:                           # :842-851
            k = _head_layernorm(sd, p + "norm_K", k, eps)
            v = _head_layernorm(sd, p + "norm_V", v, eps)
        else:                                                          # :859-864
            k = _head_layernorm(sd, p + "norm_K", k, eps)
            q = _head_layernorm(sd, p + "norm_Q", q, eps)
    use_pos = pos is not None and (pos_dim is None or pos_dim > 0)
    if use_pos:                                                        # :869-874, pos FIRST
        pp = pos.unsqueeze(1).expand(-1, n_head, -1, -1)
        q, k, v = [torch.cat([pp, t], dim=-1) for t in (q, k, v)]
    n = q.shape[-2]
    d = q.shape[-1]

    def _drop(w):
        if attn_mask is not None:
            return w * (attn_mask.to(w.dtype) * 2.0)
        if attn_dropout:
            return F.dropout(w)                                        # p=.5, always on
        return w

    if attention_type in GALERKIN_TYPES:                               # :708-734
        w = _drop(torch.matmul(k.transpose(-2, -1), v) / n)
        x = torch.matmul(q, w)
    elif attention_type in FOURIER_TYPES:                              # :687-703
        w = _drop(torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(d) / n)
        x = torch.matmul(w, v)
    else:
        raise NotImplementedError(attention_type)
    out = x.transpose(1, 2).reshape(bsz, -1, n_head * d)               # :892-894
    if use_pos:
        out = _lin(sd, p + "fc", out)                                  # :896-897
    return out, w


def feed_forward(sd, prefix, x, activation="relu"):
    """FeedForward.forward, libs/layers.py:979-987 (no batch norm, dropout off)."""
    return _lin(sd, _j(prefix, "lr2"), _act(activation, _lin(sd, _j(prefix, "lr1"), x)))


def encoder_layer(sd, prefix, x, pos=None, *, n_head, attention_type="galerkin",
                  layer_norm=False, attn_norm=None, norm_eps=1e-5, pos_dim=1,
                  residual_type="add", activation_type="relu", attn_mask=None,
                  attn_dropout=False, return_attn=False):
    """SimpleTransformerEncoderLayer.forward, libs/model.py:104-140."""
    if attn_norm is None:
        attn_norm = not layer_norm                                     # model.py:63-65
    if (not layer_norm) and (not attn_norm):
        attn_norm = True
    norm_eps = 1e-5 if norm_eps is None else norm_eps
    d_model = x.shape[-1]
    att, w = simple_attention(sd, _j(prefix, "attn"), x, x, x,
                              pos if (pos is not None and pos_dim > 0) else None,
                              n_head=n_head, attention_type=attention_type, norm=attn_norm,
                              eps=norm_eps, pos_dim=pos_dim, attn_mask=attn_mask,
                              attn_dropout=attn_dropout)
    x = x + att if residual_type in ("add", "plus", None) else x - att  # :124-127
    if layer_norm:
        x = F.layer_norm(x, (d_model,), sd[_j(prefix, "layer_norm1.weight")],
                         sd[_j(prefix, "layer_norm1.bias")], norm_eps)
    x = x + feed_forward(sd, _j(prefix, "ff"), x, activation_type or "relu")  # :131-132
    if layer_norm:
        x = F.layer_norm(x, (d_model,), sd[_j(prefix, "layer_norm2.weight")],
                         sd[_j(prefix, "layer_norm2.bias")], norm_eps)
    return (x, w) if return_attn else x


def spectral_conv1d(sd, prefix, x, *, modes, activation="silu", return_freq=False):
    """SpectralConv1d.forward, libs/layers.py:1077-1106.  x: (B, n, C_in)."""
    n = x.shape[1]
    res = _lin(sd, _j(prefix, "linear"), x)
    w = torch.view_as_complex(sd[_j(prefix, "fourier_weight")].contiguous())  # (in,out,m)
    x_ft = torch.fft.rfft(x.permute(0, 2, 1), n=n, norm="ortho")
    out_modes = torch.einsum("bix,iox->box", x_ft[:, :, :modes], w)    # :1068-1075
    out_ft = x_ft.new_zeros(x.shape[0], w.shape[1], n // 2 + 1)
    out_ft[:, :, :modes] = out_modes                                   # :1093-1094
    y = torch.fft.irfft(out_ft, n=n, norm="ortho").permute(0, 2, 1)
    y = _act(activation, y + res)
    return (y, out_ft) if return_freq else y


def spectral_conv2d(sd, prefix, x, *, modes, activation="silu", return_freq=False):
    """SpectralConv2d.forward, libs/layers.py:1153-1197.

    x: (B, n, n, C_in) or (B, n*n, C_in)."""
    bsz, three_d = x.shape[0], x.dim() == 3
    n = int(x.shape[1] ** 0.5) if three_d else x.shape[1]
    c_in = x.shape[-1]
    x = x.reshape(-1, n, n, c_in)
    res = _lin(sd, _j(prefix, "linear"), x)
    w0 = torch.view_as_complex(sd[_j(prefix, "fourier_weight.0")].contiguous())
    w1 = torch.view_as_complex(sd[_j(prefix, "fourier_weight.1")].contiguous())
    x_ft = torch.fft.rfft2(x.permute(0, 3, 1, 2), s=(n, n), norm="ortho")
    out_ft = x_ft.new_zeros(bsz, w0.shape[1], n, n // 2 + 1)
    out_ft[:, :, :modes, :modes] = torch.einsum(
        "bixy,ioxy->boxy", x_ft[:, :, :modes, :modes], w0)             # :1181-1182
    out_ft[:, :, -modes:, :modes] = torch.einsum(
        "bixy,ioxy->boxy", x_ft[:, :, -modes:, :modes], w1)            # :1183-1184
    y = torch.fft.irfft2(out_ft, s=(n, n), norm="ortho").permute(0, 2, 3, 1)
    y = _act(activation, y + res)
    if three_d:
        y = y.reshape(bsz, n * n, -1)
    return (y, out_ft) if return_freq else y


def spectral_regressor(sd, prefix, x, grid=None, *, modes, num_spectral_layers=2,
                       spacial_dim=2, spacial_fc=False, activation="silu",
                       last_activation=True, normalizer=None):
    """SpectralRegressor.forward, libs/model.py:603-637."""
    conv = spectral_conv2d if spacial_dim == 2 else spectral_conv1d
    if spacial_fc:
        x = _lin(sd, _j(prefix, "fc"), torch.cat([x, grid], dim=-1))     # :615-617
    for i in range(num_spectral_layers):
        act = activation
        if i == num_spectral_layers - 1 and not last_activation:       # :588-589
            act = "identity"
        x = conv(sd, _j(prefix, f"spectral_conv.{i}"), x, modes=modes, activation=act)
    x = _lin(sd, _j(prefix, "regressor.0"), x)
    x = _lin(sd, _j(prefix, "regressor.2"), _act(activation, x))
    if normalizer is not None:
        x = normalizer.inverse_transform(x)
    return x


def pointwise_regressor(sd, prefix, x, grid=None, *, num_layers=2, spacial_fc=False,
                        activation="silu"):
    """PointwiseRegressor.forward, libs/model.py:507-529 (dropout off)."""
    if spacial_fc:
        x = _lin(sd, _j(prefix, "fc"), torch.cat([x, grid], dim=-1))
    for i in range(num_layers):
        x = _act(activation, _lin(sd, _j(prefix, f"ff.{i}.0"), x))
    return _lin(sd, _j(prefix, "out"), x)


# ---------------------------------------------------------------------------------
# CNN down/up-scalers (libs/layers.py:88-150, 431-512, 624-670; libs/model.py:640-749).
# Not hot-path operators, but part of the model the metric is quoted on.
# ---------------------------------------------------------------------------------
def _conv_block(sd, prefix, x, activation, padding=1):
    """Conv2dResBlock with residual=False, basic_block=False: act(conv3x3(x)), no bias."""
    return _act(activation, F.conv2d(x, sd[_j(prefix, "conv.0.weight")], None, padding=padding))


def _interp(x, size_or_scale):
    """layers.py:485-493: a float is a scale factor (recomputed), a pair is a size."""
    if isinstance(size_or_scale, float):
        return F.interpolate(x, scale_factor=size_or_scale, mode="bilinear",
                             recompute_scale_factor=True, align_corners=True)
    return F.interpolate(x, size=tuple(size_or_scale), mode="bilinear", align_corners=True)


def interp_downscaler(sd, prefix, x, interp_size, activation="relu"):
    """DownScaler(downsample_mode='interp') -> Interp2dEncoder.forward,
    libs/model.py:675-687 + libs/layers.py:483-512.  x: (B, n, n, C) channel-last."""
    p = _j(prefix, "downsample")
    x = x.permute(0, 3, 1, 2)
    x = _conv_block(sd, p + ".conv0", x, activation)
    x = _act(activation, _interp(x, interp_size[0]))
    x1 = _conv_block(sd, p + ".conv1", x, activation)
    x2 = _conv_block(sd, p + ".conv2", x1, activation)
    x3 = _conv_block(sd, p + ".conv3", x2, activation)
    out = torch.cat([x1, x2, x3], dim=1)
    out = _act(activation, _interp(out, interp_size[1]))
    return out.permute(0, 2, 3, 1)


def interp_upscaler(sd, prefix, x, interp_size, activation="silu"):
    """UpScaler(upsample_mode='interp') -> Interp2dUpsample.forward,
    libs/model.py:740-749 + libs/layers.py:661-670: interp -> act(act(conv)) -> interp."""
    p = _j(prefix, "upsample")
    x = x.permute(0, 3, 1, 2)
    x = _interp(x, interp_size[0])
    x = _act(activation, _conv_block(sd, p + ".conv.0", x, activation))
    x = _interp(x, interp_size[1])
    return x.permute(0, 2, 3, 1)


# ---------------------------------------------------------------------------------
# Model assemblies
# ---------------------------------------------------------------------------------
def _encoder_kwargs(cfg):
    return dict(n_head=cfg["n_head"], attention_type=cfg["attention_type"],
                layer_norm=bool(cfg.get("layer_norm")), attn_norm=cfg.get("attn_norm"),
                norm_eps=cfg.get("norm_eps"), pos_dim=cfg["pos_dim"])


def fourier_transformer_2d(sd, cfg, node, pos, grid, *, attn_masks=None, attn_dropout=False):
    """FourierTransformer2D.forward, libs/model.py:953-1017 (no GCN/GAT, dropout off)."""
    bsz = node.shape[0]
    n_s = int(pos.shape[1] ** 0.5)
    n_hidden = cfg["n_hidden"]
    if cfg.get("downscaler_size"):
        x = interp_downscaler(sd, "downscaler", node, cfg["downscaler_size"],
                              cfg.get("downscaler_activation") or "silu")
    else:                                                              # model.py:967-969, 1113
        x = _lin(sd, "downscaler.id", torch.cat([node, pos.reshape(bsz, n_s, n_s, -1)], -1))
    x = x.reshape(bsz, -1, n_hidden)
    ek = _encoder_kwargs(cfg)
    for i in range(cfg["num_encoder_layers"]):                         # :976-981
        x = encoder_layer(sd, f"encoder_layers.{i}", x, pos, **ek,
                          attn_mask=None if attn_masks is None else attn_masks[i],
                          attn_dropout=attn_dropout)
    x = x.reshape(bsz, n_s, n_s, n_hidden)
    if cfg.get("upscaler_size"):
        x = interp_upscaler(sd, "upscaler", x, cfg["upscaler_size"],
                            cfg.get("upscaler_activation") or "silu")
    if cfg["decoder_type"] == "ifft2":
        x = spectral_regressor(sd, "regressor", x, grid, modes=cfg["fourier_modes"],
                               num_spectral_layers=cfg["num_regressor_layers"],
                               spacial_dim=cfg["spacial_dim"], spacial_fc=cfg["spacial_fc"],
                               activation=cfg.get("regressor_activation") or "silu",
                               last_activation=cfg.get("last_activation", True))
    else:
        x = pointwise_regressor(sd, "regressor", x, grid,
                                num_layers=cfg["num_regressor_layers"],
                                spacial_fc=cfg["spacial_fc"],
                                activation=cfg.get("regressor_activation") or "silu")
    if cfg.get("boundary_condition") == "dirichlet":                   # :1008-1010
        x = F.pad(x[:, 1:-1, 1:-1], (0, 0, 1, 1, 1, 1), "constant", 0)
    return x


def simple_transformer(sd, cfg, node, pos, grid=None, *, attn_masks=None, attn_dropout=False):
    """SimpleTransformer.forward, libs/model.py:760-807 (Identity feature extractor,
    'ifft' SpectralRegressor decoder with dim_feedforward=freq_dim)."""
    x = _lin(sd, "feat_extract.id", node)
    ek = _encoder_kwargs(cfg)
    ek["residual_type"] = cfg.get("residual_type") or "add"
    ek["activation_type"] = cfg.get("attn_activation") or "relu"
    for i in range(cfg["num_encoder_layers"]):
        x = encoder_layer(sd, f"encoder_layers.{i}", x, pos, **ek,
                          attn_mask=None if attn_masks is None else attn_masks[i],
                          attn_dropout=attn_dropout)
    return spectral_regressor(sd, "regressor", x, grid, modes=cfg["fourier_modes"],
                              num_spectral_layers=cfg["num_regressor_layers"],
                              spacial_dim=cfg.get("spacial_dim") or cfg["pos_dim"],
                              spacial_fc=bool(cfg.get("spacial_fc")),
                              activation=cfg.get("regressor_activation") or "silu")


def fourier_transformer_2d_lite(sd, cfg, node, pos, grid, *, attn_masks=None,
                                attn_dropout=False):
    """FourierTransformer2DLite.forward, libs/model.py:1196-1226."""
    bsz, n_grid = node.shape[0], grid.shape[1]
    x = torch.cat([node.reshape(bsz, -1, node.shape[-1]), pos], dim=-1)
    x = _lin(sd, "feat_extract.id", x)
    ek = _encoder_kwargs(cfg)
    for i in range(cfg["num_encoder_layers"]):
        x = encoder_layer(sd, f"encoder_layers.{i}", x, pos, **ek,
                          attn_mask=None if attn_masks is None else attn_masks[i],
                          attn_dropout=attn_dropout)
    x = x.reshape(bsz, n_grid, n_grid, -1)
    return spectral_regressor(sd, "regressor", x, grid, modes=cfg["fourier_modes"],
                              num_spectral_layers=cfg["num_regressor_layers"],
                              spacial_dim=cfg.get("spacial_dim") or cfg["pos_dim"],
                              spacial_fc=bool(cfg.get("spacial_fc")),
                              activation=cfg.get("regressor_activation") or "silu")
