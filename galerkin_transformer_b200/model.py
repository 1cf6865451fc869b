"""Model assemblies around the B200 operators, with the reference's constructor / forward
signatures and state_dict layout (scaomath/galerkin-transformer `libs/model.py`):

  SimpleTransformerEncoderLayer  model.py:33-140      SpectralRegressor   model.py:532-637
  PointwiseRegressor             model.py:472-529     DownScaler/UpScaler model.py:640-749 ('interp')
  SimpleTransformer              model.py:752-942     FourierTransformer2D model.py:945-1184
  FourierTransformer2DLite       model.py:1186-1283

The encoder layers, regressors and every nn.Linear on the path run on the CUDA library.  Of the
interpolation-CNN down/up-scalers (the first "next" row) the bilinear resizes run on the
library's streaming kernel (csrc/interp.cu); their 3x3 convolutions stay stock cuDNN, kept in channels-last so no
permute copies surround them.
"""
import copy
import math
from collections import defaultdict

import torch
import torch.nn.functional as F
from torch import nn

from . import functional as GF
from .layers import (FeedForward, Identity, SimpleAttention, SpectralConv1d, SpectralConv2d)

ADDITIONAL_ATTR = ['normalizer', 'raw_laplacian', 'return_latent', 'residual_type', 'norm_type',
                   'norm_eps', 'boundary_condition', 'upscaler_size', 'downscaler_size', 'spacial_dim',
                   'spacial_fc', 'regressor_activation', 'attn_activation', 'downscaler_activation',
                   'upscaler_activation', 'encoder_dropout', 'decoder_dropout', 'ffn_dropout']


def _default(value, d):
    return d if value is None else value


def _activation(name):
    return nn.SiLU() if name == 'silu' else nn.ReLU()


def _act_name(module):
    return 'silu' if isinstance(module, nn.SiLU) else 'relu'


class SimpleTransformerEncoderLayer(nn.Module):
    def __init__(self, d_model=96, pos_dim=1, n_head=2, dim_feedforward=512, attention_type='fourier',
                 pos_emb=False, layer_norm=True, attn_norm=None, norm_type='layer', norm_eps=None,
                 batch_norm=False, attn_weight=False, xavier_init: float = 1e-2,
                 diagonal_weight: float = 1e-2, symmetric_init=False, residual_type='add',
                 activation_type='relu', dropout=0.1, ffn_dropout=None, debug=False):
        super().__init__()
        if pos_emb:
            raise NotImplementedError("sinusoidal pos_emb (off in every shipped config)")
        dropout = _default(dropout, 0.05)
        if attention_type in ['linear', 'softmax']:
            dropout = 0.1
        ffn_dropout = _default(ffn_dropout, dropout)
        norm_eps = _default(norm_eps, 1e-5)
        attn_norm = _default(attn_norm, not layer_norm)
        if (not layer_norm) and (not attn_norm):
            attn_norm = True
        norm_type = _default(norm_type, 'layer')
        self.attn = SimpleAttention(n_head=n_head, d_model=d_model, attention_type=attention_type,
                                    diagonal_weight=diagonal_weight, xavier_init=xavier_init,
                                    symmetric_init=symmetric_init, pos_dim=pos_dim, norm=attn_norm,
                                    norm_type=norm_type, eps=norm_eps, dropout=dropout)
        self.d_model = d_model
        self.n_head = n_head
        self.pos_dim = pos_dim
        self.add_layer_norm = layer_norm
        if layer_norm:
            self.layer_norm1 = nn.LayerNorm(d_model, eps=norm_eps)
            self.layer_norm2 = nn.LayerNorm(d_model, eps=norm_eps)
        dim_feedforward = _default(dim_feedforward, 2 * d_model)
        self.ff = FeedForward(in_dim=d_model, dim_feedforward=dim_feedforward, batch_norm=batch_norm,
                              activation=activation_type, dropout=ffn_dropout)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.residual_type = residual_type
        self.add_pos_emb = pos_emb
        self.debug = debug
        self.attn_weight = attn_weight
        self.attn.materialize_attn = bool(attn_weight)
        self.__name__ = attention_type.capitalize() + 'TransformerEncoderLayer'

    def forward(self, x, pos=None, weight=None):
        """x: (B, n, d_model); pos: (B, n, pos_dim) joined to every head."""
        a = self.attn
        use_pos = pos is not None and self.pos_dim > 0
        p1 = self.dropout1.p if self.training else 0.0
        p2 = self.dropout2.p if self.training else 0.0
        sign = 1.0 if (self.residual_type in ['add', 'plus'] or self.residual_type is None) else -1.0
        if self._fused_ok(x, pos, weight):
            return self._forward_fused(x, pos, p1, p2, sign)
        if use_pos:
            # attention core, then  x +/- dropout1(fc(heads))  in the fc GEMM's epilogue
            heads, attn_weight = a.forward_heads(x, x, x, pos=pos, weight=weight)
            x = GF.linear(heads, a.fc.weight, a.fc.bias, residual=x, rscale=sign, drop_p=p1)
        else:
            att_output, attn_weight = a(x, x, x, weight=weight)
            att_output = F.dropout(att_output, p1, True) if p1 > 0 else att_output
            x = x + sign * att_output
        if self.add_layer_norm:
            x = GF.layer_norm(x, self.layer_norm1.weight, self.layer_norm1.bias, self.layer_norm1.eps)
        # x + dropout2(ff(x)) in the second FFN GEMM's epilogue
        x = self.ff(x, residual=x, rscale=1.0, out_drop_p=p2)
        if self.add_layer_norm:
            x = GF.layer_norm(x, self.layer_norm2.weight, self.layer_norm2.bias, self.layer_norm2.eps)
        if self.attn_weight:
            return x, attn_weight
        return x


    # ---- fused path: the whole layer as three tcgen05 kernels (csrc/encoder_fwd.cu) ----
    def _fused_static_ok(self):
        a = self.attn
        if GF.get_precision() != 'x3' or self.add_layer_norm or a.attention_type != 'galerkin':
            return False
        if a.pos_dim < 1:                      # without position columns the reference layer has no `fc` at all
            return False
        if self.ff.act_name != 'relu' or self.ff.lr2.out_features != self.d_model:
            return False
        return GF.encoder_fused_supported(self.d_model, self.n_head, a.pos_dim, self.ff.lr1.out_features)

    def _fused_ok(self, x, pos, weight):
        if weight is not None or pos is None or x.dim() != 3 or not x.is_cuda:
            return False
        return self._fused_static_ok()

    def _fused_params(self):
        a = self.attn
        ps = [lin.weight for lin in a.linears] + [lin.bias for lin in a.linears]
        if a.add_norm:
            for mods, attr in ((a.norm_K, "weight"), (a.norm_K, "bias"), (a.norm_V, "weight"), (a.norm_V, "bias")):
                ps += [getattr(m, attr) for m in mods]
        ps += [a.fc.weight, a.fc.bias]
        return ps + [self.ff.lr1.weight, self.ff.lr1.bias, self.ff.lr2.weight, self.ff.lr2.bias]

    def prepack(self):
        """Pack this layer's parameters for the fused kernels now (consumed by the next forward); lets a model pack
        all its layers up front, off the critical path."""
        a = self.attn
        self._prepacked = GF.encoder_pack(self._fused_params(), a.n_head, self.d_model, a.pos_dim, self.ff.lr1.out_features)

    def _forward_fused(self, x, pos, p1, p2, sign):
        a = self.attn
        bsz, n = x.size(0), x.size(1)
        assert pos.size(-1) == a.pos_dim
        d = a.d_k + a.pos_dim
        keep = a._next_mask
        a._next_mask = None
        mask_p = 0.0
        if keep is None and a.attn_dropout == 'reference':
            mask_p = 0.5
        elif keep is not None:
            assert tuple(keep.shape) == (bsz, a.n_head, d, d), f"keep-mask shape {tuple(keep.shape)}"
            keep = keep.to(device=x.device, dtype=torch.uint8).contiguous()
        pf = self.ff.dropout.p if self.training else 0.0
        y, attn_weight = GF.encoder_layer(x, pos, self._fused_params(), n_head=a.n_head, pos_dim=a.pos_dim, eps=a.eps,
                                          attention_scale=1.0 / n, keep_mask=keep, mask_p=mask_p, p_attn_out=p1,
                                          res_sign=sign, p_ffn=pf, p_out=p2, packed=self.__dict__.pop('_prepacked', None))
        a.attn_weight = attn_weight
        if self.attn_weight:
            return y, attn_weight
        return y


def prepack_encoder_layers(layers, like, pos, weight=None):
    """Pack the parameters of every fused encoder layer on an auxiliary stream, so the ten pack launches overlap
    whatever precedes the encoder stack (the down-scaler); returns the stream the caller must wait on before the first
    layer runs, or None when nothing was packed."""
    if weight is not None or pos is None or not like.is_cuda:
        return None
    todo = [l for l in layers if isinstance(l, SimpleTransformerEncoderLayer) and l._fused_static_ok()]
    if not todo:
        return None
    main = torch.cuda.current_stream(like.device)
    side = GF.aux_stream(like.device)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        for l in todo:
            l.prepack()
    return side


class PointwiseRegressor(nn.Module):
    def __init__(self, in_dim, n_hidden, out_dim, num_layers: int = 2, spacial_fc: bool = False,
                 spacial_dim=1, dropout=0.1, activation='silu', return_latent=False, debug=False):
        super().__init__()
        dropout = _default(dropout, 0.1)
        self.spacial_fc = spacial_fc
        activ = _activation(activation)
        if self.spacial_fc:
            in_dim = in_dim + spacial_dim
            self.fc = nn.Linear(in_dim, n_hidden)
        self.ff = nn.ModuleList([nn.Sequential(nn.Linear(n_hidden, n_hidden), activ)
                                 for _ in range(num_layers)])
        self.dropout = nn.Dropout(dropout)
        self.out = nn.Linear(n_hidden, out_dim)
        self.return_latent = return_latent
        self.debug = debug

    def forward(self, x, grid=None):
        if self.spacial_fc:
            x = GF.linear_cat(x, grid, self.fc.weight, self.fc.bias)
        p = self.dropout.p if self.training else 0.0
        for layer in self.ff:
            x = GF.linear(x, layer[0].weight, layer[0].bias, act=_act_name(layer[1]), drop_p=p)
        x = GF.linear(x, self.out.weight, self.out.bias)
        return (x, None) if self.return_latent else x


class SpectralRegressor(nn.Module):
    def __init__(self, in_dim, n_hidden, freq_dim, out_dim, modes: int, num_spectral_layers: int = 2,
                 n_grid=None, dim_feedforward=None, spacial_fc=False, spacial_dim=2, return_freq=False,
                 return_latent=False, normalizer=None, activation='silu', last_activation=True,
                 dropout=0.1, debug=False):
        super().__init__()
        if spacial_dim == 2:
            conv = SpectralConv2d
        elif spacial_dim == 1:
            conv = SpectralConv1d
        else:
            raise NotImplementedError("3D not implemented.")
        activation = _default(activation, 'silu')
        self.activation = _activation(activation)
        dropout = _default(dropout, 0.1)
        self.spacial_fc = spacial_fc
        if self.spacial_fc:
            self.fc = nn.Linear(in_dim + spacial_dim, n_hidden)
        dims = [n_hidden] + [freq_dim] * num_spectral_layers
        self.spectral_conv = nn.ModuleList(
            [conv(in_dim=dims[i], out_dim=dims[i + 1], n_grid=n_grid, modes=modes, dropout=dropout,
                  activation=activation, return_freq=return_freq, debug=debug)
             for i in range(num_spectral_layers)])
        if not last_activation:
            self.spectral_conv[-1].activation = Identity()
        self.n_grid = n_grid
        self.dim_feedforward = _default(dim_feedforward, 2 * spacial_dim * freq_dim)
        self.regressor = nn.Sequential(nn.Linear(freq_dim, self.dim_feedforward), self.activation,
                                       nn.Linear(self.dim_feedforward, out_dim))
        self.normalizer = normalizer
        self.return_freq = return_freq
        self.return_latent = return_latent
        self.debug = debug

    def forward(self, x, edge=None, pos=None, grid=None):
        x_latent, x_fts = [], []
        if self.spacial_fc:
            x = GF.linear_cat(x, grid, self.fc.weight, self.fc.bias)
        for layer in self.spectral_conv:
            if self.return_freq:
                x, x_ft = layer(x)
                x_fts.append(x_ft.contiguous())
            else:
                x = layer(x)
            if self.return_latent:
                x_latent.append(x.contiguous())
        r0, r2 = self.regressor[0], self.regressor[2]
        x = GF.mlp2(x, r0.weight, r0.bias, r2.weight, r2.bias, act=_act_name(self.regressor[1]))
        if self.normalizer:
            x = self.normalizer.inverse_transform(x)
        if self.return_freq or self.return_latent:
            return x, dict(preds_freq=x_fts, preds_latent=x_latent)
        return x


# -------------------------------------------------------------------------------------------
# interpolation-CNN scalers: stock PyTorch / cuDNN, channels-last
# -------------------------------------------------------------------------------------------
class Conv2dResBlock(nn.Module):
    """conv3x3 (no bias) -> dropout -> activation; the residual / basic-block variants of
    libs/layers.py:88-150 are never enabled by the interp scalers and are not implemented."""

    def __init__(self, in_dim, out_dim, kernel_size=3, padding=1, dilation=1, dropout=0.1, stride=1,
                 bias=False, residual=False, basic_block=False, activation_type='silu'):
        super().__init__()
        if residual or basic_block:
            raise NotImplementedError("Conv2dResBlock(residual/basic_block)")
        self.activation = _activation(_default(activation_type, 'silu'))
        self.conv = nn.Sequential(nn.Conv2d(in_dim, out_dim, kernel_size=kernel_size, padding=padding,
                                            dilation=dilation, stride=stride, bias=bias),
                                  nn.Dropout(dropout))
        self.add_res = residual

    def forward(self, x):
        conv, drop = self.conv[0], self.conv[1]
        act = _act_name(self.activation)
        # 'x3' mode: the library's bf16x3 tensor-core convolution (csrc/conv.cu) with dropout + activation in its epilogue;
        # 'fp32' / 'tf32' modes: stock cuDNN at that precision
        if (GF.get_precision() == 'x3' and x.is_cuda and conv.bias is None and conv.kernel_size == (3, 3)
                and conv.padding == (1, 1) and conv.stride == (1, 1) and conv.dilation == (1, 1) and conv.groups == 1
                and isinstance(self.activation, (nn.ReLU, nn.SiLU))
                and (conv.in_channels > 1 or (act == 'relu' and conv.out_channels % 4 == 0))
                and GF.conv3x3_supported(conv.in_channels, conv.out_channels)):
            p = drop.p if self.training else 0.0
            y = GF.conv3x3_block(x.permute(0, 2, 3, 1), conv.weight, act=act, drop_p=p)
            return y.permute(0, 3, 1, 2)
        return self.activation(self.conv(x))


def _resize(x, spec):
    """F.interpolate(x, size=... | scale_factor=..., mode='bilinear', align_corners=True[, recompute_scale_factor=True])
    for an NCHW view of channel-last memory, on the library's streaming resize kernel (csrc/interp.cu)."""
    H, W = x.shape[-2:]
    if isinstance(spec, float):       # recompute_scale_factor=True: only the floor()ed output size reaches the kernel
        size = (int(math.floor(float(H) * spec)), int(math.floor(float(W) * spec)))
    else:
        size = tuple(int(s) for s in spec)
    y = GF.interp_bilinear(x.permute(0, 2, 3, 1), size[0], size[1])      # a view when x is channels-last
    return y.permute(0, 3, 1, 2)


class Interp2dEncoder(nn.Module):
    """libs/layers.py:431-512 with residual=False."""

    def __init__(self, in_dim, out_dim, kernel_size=3, stride=1, padding=1, dilation=1, interp_size=None,
                 residual=False, activation_type='silu', dropout=0.1, debug=False):
        super().__init__()
        if residual:
            raise NotImplementedError("Interp2dEncoder(residual=True)")
        c0 = out_dim // 3
        c2 = out_dim - 2 * c0
        p1 = max(padding // 2, 1)
        p2 = max(padding // 4, 1)
        activation_type = _default(activation_type, 'silu')
        self.interp_size = interp_size
        blk = dict(kernel_size=kernel_size, dropout=dropout, activation_type=activation_type)
        self.conv0 = Conv2dResBlock(in_dim, out_dim, padding=padding, **blk)
        self.conv1 = Conv2dResBlock(out_dim, c0, padding=p1, stride=stride, **blk)
        self.conv2 = Conv2dResBlock(c0, c0, dilation=dilation, padding=p2, **blk)
        self.conv3 = Conv2dResBlock(c0, c2, **blk)
        self.activation = _activation(activation_type)
        self.add_res = residual

    def forward(self, x):
        x = self.activation(_resize(self.conv0(x), self.interp_size[0]))
        x1 = self.conv1(x)
        x2 = self.conv2(x1)
        x3 = self.conv3(x2)
        out = torch.cat([x1, x2, x3], dim=1)
        return self.activation(_resize(out, self.interp_size[1]))


class Interp2dUpsample(nn.Module):
    """libs/layers.py:624-670: interp -> [conv block -> dropout -> act] -> interp."""

    def __init__(self, in_dim, out_dim, kernel_size=3, padding=1, residual=False, conv_block=True,
                 interp_mode='bilinear', interp_size=None, activation_type='silu', dropout=0.1,
                 debug=False):
        super().__init__()
        activation_type = _default(activation_type, 'silu')
        self.activation = _activation(activation_type)
        self.dropout = nn.Dropout(dropout)
        if conv_block:
            self.conv = nn.Sequential(Conv2dResBlock(in_dim, out_dim, kernel_size=kernel_size,
                                                     padding=padding, residual=residual, dropout=dropout,
                                                     activation_type=activation_type),
                                      self.dropout, self.activation)
        self.conv_block = conv_block
        self.interp_size = interp_size
        self.interp_mode = interp_mode

    def forward(self, x):
        if self.interp_mode != 'bilinear':
            raise NotImplementedError(f"Interp2dUpsample(interp_mode={self.interp_mode!r})")
        x = _resize(x, self.interp_size[0])
        if self.conv_block:
            x = self.conv(x)
        return _resize(x, self.interp_size[1])


class DownScaler(nn.Module):
    def __init__(self, in_dim, out_dim, dropout=0.1, padding=5, downsample_mode='conv',
                 activation_type='silu', interp_size=None, debug=False):
        super().__init__()
        if downsample_mode != 'interp':
            raise NotImplementedError("DownScaler: only downsample_mode='interp' (the shipped configs)")
        self.downsample = Interp2dEncoder(in_dim=in_dim, out_dim=out_dim, interp_size=interp_size,
                                          activation_type=activation_type, dropout=dropout, debug=debug)
        self.in_dim = in_dim
        self.out_dim = out_dim

    def forward(self, x):
        """(B, n, n, in_dim) -> (B, n_s, n_s, out_dim)"""
        bsz, n = x.size(0), x.size(1)
        x = x.view(bsz, n, n, self.in_dim).permute(0, 3, 1, 2)          # NCHW view of NHWC memory
        x = self.downsample(x.contiguous(memory_format=torch.channels_last))
        return x.permute(0, 2, 3, 1)                                     # contiguous if channels-last


class UpScaler(nn.Module):
    def __init__(self, in_dim: int, out_dim: int, hidden_dim=None, padding=2, output_padding=0,
                 dropout=0.1, upsample_mode='conv', activation_type='silu', interp_mode='bilinear',
                 interp_size=None, debug=False):
        super().__init__()
        if upsample_mode != 'interp':
            raise NotImplementedError("UpScaler: only upsample_mode='interp' (the shipped configs)")
        self.upsample = Interp2dUpsample(in_dim=in_dim, out_dim=out_dim, interp_mode=interp_mode,
                                         interp_size=interp_size, dropout=dropout,
                                         activation_type=activation_type, debug=debug)
        self.in_dim = in_dim
        self.out_dim = out_dim

    def forward(self, x):
        """(B, n_s, n_s, in_dim) -> (B, n, n, out_dim)"""
        x = x.permute(0, 3, 1, 2)
        x = self.upsample(x.contiguous(memory_format=torch.channels_last))
        return x.permute(0, 2, 3, 1)


# -------------------------------------------------------------------------------------------
# models
# -------------------------------------------------------------------------------------------
class _ConfiguredModel(nn.Module):
    """kwargs -> attributes, missing keys read as None (libs/model.py:25-30, 832-835)."""

    KNOWN_KEYS = ['node_feats', 'edge_feats', 'pos_dim', 'n_targets', 'n_hidden', 'num_feat_layers',
                  'num_encoder_layers', 'n_head', 'pred_len', 'n_freq_targets', 'dim_feedforward',
                  'feat_extract_type', 'graph_activation', 'attention_type', 'xavier_init', 'diagonal_weight',
                  'symmetric_init', 'layer_norm', 'attn_norm', 'batch_norm', 'spacial_residual',
                  'return_attn_weight', 'seq_len', 'bulk_regression', 'decoder_type', 'freq_dim',
                  'num_regressor_layers', 'fourier_modes', 'dropout', 'downscaler_dropout',
                  'upscaler_dropout', 'upsample_mode', 'downsample_mode', 'last_activation', 'debug']

    # the reference moves its (non-Module) normalizer together with the model: libs/model.py:1026-1042
    def _move_normalizer(self, fn):
        for owner in (self, getattr(self, "regressor", None)):
            norm = getattr(owner, "normalizer", None) if owner is not None else None
            if norm is not None and hasattr(norm, fn):
                moved = getattr(norm, fn)()
                if moved is not None:
                    owner.normalizer = moved

    def cuda(self, device=None):
        out = super().cuda(device)
        self._move_normalizer("cuda")
        return out

    def cpu(self):
        out = super().cpu()
        self._move_normalizer("cpu")
        return out

    def to(self, *args, **kwargs):
        out = super().to(*args, **kwargs)
        dev = next((a for a in args if isinstance(a, (str, torch.device))), kwargs.get("device"))
        if dev is not None:
            for owner in (self, getattr(self, "regressor", None)):
                norm = getattr(owner, "normalizer", None) if owner is not None else None
                if norm is not None and hasattr(norm, "to"):
                    moved = norm.to(dev)
                    if moved is not None:
                        owner.normalizer = moved
        return out

    def _absorb(self, kwargs):
        self.config = defaultdict(lambda: None, **kwargs)
        for key in list(self.config.keys()) + ADDITIONAL_ATTR + self.KNOWN_KEYS:
            setattr(self, key, self.config[key])
        self.symmetric_init = bool(self.symmetric_init)
        self.batch_norm = bool(self.batch_norm)
        self.layer_norm = bool(self.layer_norm)
        self.last_activation = True if self.last_activation is None else self.last_activation
        self.dim_feedforward = _default(self.dim_feedforward, 2 * self.n_hidden)
        self.dropout = _default(self.dropout, 0.05)
        self.dpo = nn.Dropout(self.dropout)
        if self.decoder_type == 'attention':
            raise NotImplementedError("decoder_type='attention'")
        if self.num_feat_layers and self.num_feat_layers > 0 and self.feat_extract_type in ('gcn', 'gat'):
            raise NotImplementedError("graph feature extractors (GCN/GAT) are out of scope; every shipped "
                                      "config uses num_feat_layers: 0")

    def _encoder_stack(self, **extra):
        layer = SimpleTransformerEncoderLayer(
            d_model=self.n_hidden, n_head=self.n_head, attention_type=self.attention_type,
            dim_feedforward=self.dim_feedforward, layer_norm=self.layer_norm, attn_norm=self.attn_norm,
            pos_dim=self.pos_dim, xavier_init=self.xavier_init, diagonal_weight=self.diagonal_weight,
            dropout=self.encoder_dropout, ffn_dropout=self.ffn_dropout, debug=self.debug, **extra)
        return nn.ModuleList([copy.deepcopy(layer) for _ in range(self.num_encoder_layers)])

    def _drop(self, x):
        return F.dropout(x, self.dpo.p, True) if (self.training and self.dpo.p > 0) else x


class SimpleTransformer(_ConfiguredModel):
    def __init__(self, **kwargs):
        super().__init__()
        self._absorb(kwargs)
        self.spacial_dim = _default(self.spacial_dim, self.pos_dim)
        self.spacial_fc = _default(self.spacial_fc, False)
        if self.n_freq_targets and self.n_freq_targets > 0:
            raise NotImplementedError("frequency regressor heads (n_freq_targets > 0)")
        if self.spacial_residual:
            raise NotImplementedError("spacial_residual")
        self.feat_extract = Identity(in_features=self.node_feats, out_features=self.n_hidden)
        self.encoder_layers = self._encoder_stack(
            norm_type=self.norm_type, batch_norm=self.batch_norm, symmetric_init=self.symmetric_init,
            attn_weight=self.return_attn_weight, residual_type=self.residual_type,
            activation_type=self.attn_activation)
        if self.decoder_type == 'pointwise':
            self.regressor = PointwiseRegressor(
                in_dim=self.n_hidden, n_hidden=self.n_hidden, out_dim=self.n_targets,
                spacial_fc=self.spacial_fc, spacial_dim=self.spacial_dim,
                activation=self.regressor_activation, dropout=self.decoder_dropout, debug=self.debug)
            for prm in self.regressor.parameters():
                nn.init.xavier_uniform_(prm, gain=1e-2) if prm.ndim > 1 else nn.init.constant_(prm, 0)
        elif self.decoder_type == 'ifft':
            self.regressor = SpectralRegressor(
                in_dim=self.n_hidden, n_hidden=self.n_hidden, freq_dim=self.freq_dim,
                out_dim=self.n_targets, num_spectral_layers=self.num_regressor_layers,
                modes=self.fourier_modes, spacial_dim=self.spacial_dim, spacial_fc=self.spacial_fc,
                dim_feedforward=self.freq_dim, activation=self.regressor_activation,
                dropout=self.decoder_dropout)
        else:
            raise NotImplementedError("Decoder type not implemented")
        self.config = dict(self.config)
        self.__name__ = self.attention_type.capitalize() + 'Transformer'

    def forward(self, node, edge, pos, grid=None, weight=None):
        x_latent, attn_weights = [], []
        x = self.feat_extract(node, edge)
        if self.return_latent:
            x_latent.append(x.contiguous())
        for encoder in self.encoder_layers:
            if self.return_attn_weight:
                x, w = encoder(x, pos, weight)
                attn_weights.append(w)
            else:
                x = encoder(x, pos, weight)
            if self.return_latent:
                x_latent.append(x.contiguous())
        x = self.regressor(self._drop(x), grid=grid)
        return dict(preds=x, preds_freq=None, preds_latent=x_latent, attn_weights=attn_weights)


class FourierTransformer2D(_ConfiguredModel):
    def __init__(self, **kwargs):
        super().__init__()
        self._absorb(kwargs)
        self.feat_extract = Identity()
        if self.downscaler_size:
            self.downscaler = DownScaler(in_dim=self.node_feats, out_dim=self.n_hidden,
                                         downsample_mode=self.downsample_mode,
                                         interp_size=self.downscaler_size,
                                         dropout=self.downscaler_dropout,
                                         activation_type=self.downscaler_activation)
        else:
            self.downscaler = Identity(in_features=self.node_feats + self.spacial_dim,
                                       out_features=self.n_hidden)
        if self.upscaler_size:
            self.upscaler = UpScaler(in_dim=self.n_hidden, out_dim=self.n_hidden,
                                     upsample_mode=self.upsample_mode, interp_size=self.upscaler_size,
                                     dropout=self.upscaler_dropout,
                                     activation_type=self.upscaler_activation)
        else:
            self.upscaler = Identity()
        self.encoder_layers = self._encoder_stack(
            batch_norm=self.batch_norm, symmetric_init=self.symmetric_init,
            attn_weight=self.return_attn_weight, norm_eps=self.norm_eps)
        if self.decoder_type == 'pointwise':
            self.regressor = PointwiseRegressor(
                in_dim=self.n_hidden, n_hidden=self.n_hidden, out_dim=self.n_targets,
                num_layers=self.num_regressor_layers, spacial_fc=self.spacial_fc,
                spacial_dim=self.spacial_dim, activation=self.regressor_activation,
                dropout=self.decoder_dropout, return_latent=self.return_latent, debug=self.debug)
        elif self.decoder_type == 'ifft2':
            self.regressor = SpectralRegressor(
                in_dim=self.n_hidden, n_hidden=self.freq_dim, freq_dim=self.freq_dim,
                out_dim=self.n_targets, num_spectral_layers=self.num_regressor_layers,
                modes=self.fourier_modes, spacial_dim=self.spacial_dim, spacial_fc=self.spacial_fc,
                activation=self.regressor_activation, last_activation=self.last_activation,
                dropout=self.decoder_dropout, return_latent=self.return_latent, debug=self.debug)
        else:
            raise NotImplementedError("Decoder type not implemented")
        self.config = dict(self.config)
        self.__name__ = self.attention_type.capitalize() + 'Transformer2D'

    def forward(self, node, edge, pos, grid, weight=None, boundary_value=None):
        """node (B,n,n,node_feats), pos (B,n_s*n_s,pos_dim), grid (B,n,n,2) -> dict(preds=(B,n,n,n_targets))"""
        bsz = node.size(0)
        n_s = int(pos.size(1) ** 0.5)
        x_latent, attn_weights = [], []
        packing = prepack_encoder_layers(self.encoder_layers, node, pos, weight)
        if not self.downscaler_size:
            node = torch.cat([node, pos.contiguous().view(bsz, n_s, n_s, -1)], dim=-1)
        x = self.downscaler(node)
        x = self._drop(x.reshape(bsz, -1, self.n_hidden))
        if packing is not None:
            torch.cuda.current_stream(x.device).wait_stream(packing)
        for encoder in self.encoder_layers:
            if self.return_attn_weight:
                x, w = encoder(x, pos, weight)
                attn_weights.append(w)
            else:
                x = encoder(x, pos, weight)
            if self.return_latent:
                x_latent.append(x.contiguous())
        x = self.upscaler(x.view(bsz, n_s, n_s, self.n_hidden))
        if self.return_latent:
            x_latent.append(x.contiguous())
        x = self._drop(x)
        if self.return_latent:
            x, xr_latent = self.regressor(x, grid=grid)
            x_latent.append(xr_latent)
        else:
            x = self.regressor(x, grid=grid)
        if self.normalizer:
            x = self.normalizer.inverse_transform(x)
        if self.boundary_condition == 'dirichlet':
            x = F.pad(x[:, 1:-1, 1:-1].contiguous(), (0, 0, 1, 1, 1, 1), "constant", 0)
            if boundary_value is not None:
                assert x.size() == boundary_value.size()
                x = x + boundary_value
        return dict(preds=x, preds_latent=x_latent, attn_weights=attn_weights)


class FourierTransformer2DLite(_ConfiguredModel):
    def __init__(self, **kwargs):
        super().__init__()
        self._absorb(kwargs)
        self.spacial_dim = _default(self.spacial_dim, self.pos_dim)
        self.spacial_fc = _default(self.spacial_fc, False)
        self.feat_extract = Identity(in_features=self.node_feats, out_features=self.n_hidden)
        self.encoder_layers = self._encoder_stack(norm_type=self.norm_type)
        self.regressor = SpectralRegressor(
            in_dim=self.n_hidden, n_hidden=self.n_hidden, freq_dim=self.freq_dim, out_dim=self.n_targets,
            num_spectral_layers=self.num_regressor_layers, modes=self.fourier_modes,
            spacial_dim=self.spacial_dim, spacial_fc=self.spacial_fc, dim_feedforward=self.freq_dim,
            activation=self.regressor_activation, dropout=self.decoder_dropout)
        self.config = dict(self.config)

    def forward(self, node, edge, pos, grid=None):
        bsz, n_grid = node.size(0), grid.size(1)
        x = torch.cat([node.reshape(bsz, -1, node.size(-1)), pos], dim=-1)
        x = self.feat_extract(x, edge)
        for encoder in self.encoder_layers:
            x = encoder(x, pos)
        x = self._drop(x).view(bsz, n_grid, n_grid, -1)
        x = self.regressor(x, grid=grid)
        return dict(preds=x, preds_freq=None, preds_latent=None, attn_weights=None)
