"""B200-native operators behind the reference's module interface.

`SimpleAttention`, `FeedForward`, `SpectralConv1d` and `SpectralConv2d` keep the constructor
signatures, forward signatures, attribute names and state_dict keys of
scaomath/galerkin-transformer `libs/layers.py` (:764-951, :954-987, :1040-1106, :1109-1197),
so `load_state_dict` interchanges with the reference and `copy.deepcopy` / pickling work.
Their forward passes run on hand-written sm_100a CUDA through `functional.py`; anything the
CUDA path does not cover raises NotImplementedError -- there is no eager/CPU fallback.
"""
import math

import torch
from torch import nn

from . import functional as GF

_GALERKIN = ("galerkin",)
_FOURIER = ("fourier", "integral", "local")


class Identity(nn.Module):
    """Placeholder that optionally holds a Linear(in, out) under `.id` (libs/layers.py:21-40)."""

    def __init__(self, in_features=None, out_features=None, *args, **kwargs):
        super().__init__()
        self.id = nn.Linear(in_features, out_features) \
            if in_features is not None and out_features is not None else nn.Identity()

    def forward(self, x, edge=None, grid=None):
        if isinstance(self.id, nn.Linear):
            return GF.linear(x, self.id.weight, self.id.bias)
        return x


class SimpleAttention(nn.Module):
    """Softmax-free multi-head attention, Galerkin Q(K^T V)/n or Fourier (Q K^T)V/(sqrt(d) n).

    Same arguments as the reference (libs/layers.py:793-803).  One extra knob:
    `attn_dropout` -- 'reference' (default) reproduces the reference's unconditional
    F.dropout(p=0.5) on the attention matrix (libs/layers.py:730-731), 'off' disables it
    (what parity tests use on both sides); `set_attn_mask(mask)` injects an explicit keep-mask
    for the next call.
    """

    def __init__(self, n_head, d_model, pos_dim: int = 1, attention_type='fourier', dropout=0.1,
                 xavier_init=1e-4, diagonal_weight=1e-2, symmetric_init=False, norm=False,
                 norm_type='layer', eps=1e-5, debug=False):
        super().__init__()
        assert d_model % n_head == 0
        if attention_type not in _GALERKIN + _FOURIER:
            raise NotImplementedError(
                f"attention_type={attention_type!r}: only 'galerkin' and 'fourier'/'integral'/'local' "
                "have a B200 kernel (softmax / linear / cosine / causal are reference baselines)")
        if norm and norm_type != 'layer':
            raise NotImplementedError("norm_type='instance' is not implemented (and is broken upstream)")
        self.attention_type = attention_type
        self.d_k = d_model // n_head
        self.n_head = n_head
        self.pos_dim = pos_dim
        self.linears = nn.ModuleList([nn.Linear(d_model, d_model) for _ in range(3)])
        self.xavier_init = xavier_init
        self.diagonal_weight = diagonal_weight
        self.symmetric_init = symmetric_init
        if xavier_init > 0:
            self._reset_parameters()
        self.add_norm = norm
        self.norm_type = norm_type
        self.eps = eps
        if norm:
            def norms():
                return nn.ModuleList([nn.LayerNorm(self.d_k, eps=eps) for _ in range(n_head)])
            self.norm_K = norms()
            if attention_type in _GALERKIN:
                self.norm_V = norms()
            else:
                self.norm_Q = norms()
        if pos_dim > 0:
            self.fc = nn.Linear(d_model + n_head * pos_dim, d_model)
        self.attn_weight = None
        self.dropout = nn.Dropout(dropout)
        self.debug = debug
        self.attn_dropout = 'reference'
        self.materialize_attn = False
        self._next_mask = None

    def _reset_parameters(self):
        # xavier-uniform with a small gain plus a scaled identity (libs/layers.py:901-913)
        for lin in self.linears:
            nn.init.xavier_uniform_(lin.weight, gain=self.xavier_init)
            if self.diagonal_weight > 0.0:
                with torch.no_grad():
                    lin.weight += self.diagonal_weight * torch.eye(lin.weight.size(-1))
            if self.symmetric_init:
                with torch.no_grad():
                    lin.weight += lin.weight.T.clone()
            nn.init.constant_(lin.bias, 0)

    def set_attn_mask(self, mask):
        """Explicit keep-mask (B,H,d,d), 1 = keep, used instead of a random draw on the next forward."""
        self._next_mask = mask

    def forward(self, query, key, value, pos=None, mask=None, weight=None):
        x, attn_weight = self.forward_heads(query, key, value, pos=pos, mask=mask, weight=weight)
        if pos is not None and self.pos_dim > 0:
            x = GF.linear(x, self.fc.weight, self.fc.bias)
        return x, attn_weight

    def _packed_parts(self):
        """Parameters in the order of the packed vector [W_qkv | b_qkv | gamma_1 | beta_1 | gamma_2 | beta_2] that
        functional._unpack_attention_params slices: block 1 = norm_K, block 2 = norm_V (Galerkin) or norm_Q (Fourier)."""
        parts = [lin.weight for lin in self.linears] + [lin.bias for lin in self.linears]
        if self.add_norm:
            second = self.norm_V if self.attention_type in _GALERKIN else self.norm_Q
            for mods, attr in ((self.norm_K, "weight"), (self.norm_K, "bias"), (second, "weight"), (second, "bias")):
                parts += [getattr(m, attr) for m in mods]
        return parts

    def forward_heads(self, query, key, value, pos=None, mask=None, weight=None):
        """Everything up to (excluding) the output `fc`: returns the head-merged
        (B, n, H*(d_k+pos_dim)) tensor so a caller can fuse fc with its residual add."""
        if mask is not None:
            if self.attention_type in _GALERKIN:
                raise RuntimeError("linear attention does not support casual mask.")
            raise NotImplementedError("attention masks are not supported by the B200 kernels")
        if weight is not None:
            raise NotImplementedError("`weight` (mass-matrix) scaling is not supported by the B200 kernels")
        use_pos = pos is not None and self.pos_dim > 0
        if use_pos:
            assert pos.size(-1) == self.pos_dim
        bsz, n = query.size(0), query.size(1)
        p = self.pos_dim if use_pos else 0
        d = self.d_k + p
        self_attn = (query is key) and (key is value)
        # W_qkv (3 d_model, d_model), b_qkv and the per-head LayerNorm tables, assembled by one pack launch
        flat = GF.pack(self._packed_parts())     # its gradient comes back as one flat buffer too

        keep = self._next_mask
        self._next_mask = None
        mask_p = 0.0
        fourier = self.attention_type in _FOURIER
        if keep is None and self.attn_dropout == 'reference':
            mask_p = 0.5          # drawn inside the kernels (Philox), never materialised
        elif keep is not None:
            want = (bsz, self.n_head, n, n) if fourier else (bsz, self.n_head, d, d)
            assert tuple(keep.shape) == want, f"keep-mask shape {tuple(keep.shape)} != {want}"
            keep = keep.to(device=query.device, dtype=torch.uint8).contiguous()
        # Fourier type: without a dropout between the two products (QK^T)V = Q(K^T V) exactly, so the O(n d^2)
        # linear-form kernels are used; with the n x n dropout the quadratic flash-style kernels run instead.
        quadratic = fourier and (keep is not None or mask_p > 0.0 or self.materialize_attn)
        if quadratic and d > 64:
            raise NotImplementedError(
                f"Fourier-type attention with the n x n dropout (attn_dropout='reference'), a keep-mask or materialize_attn "
                f"runs the quadratic flash-style kernels, which support d_k + pos_dim <= 64 (got {d}); "
                "set_attn_dropout(model, 'off') selects the exact linear-form kernels (d_k + pos_dim <= 128)")

        x, attn = GF.linear_attention(query, key, value, pos if use_pos else None, flat, bool(self.add_norm),
                                      keep, n_head=self.n_head, pos_dim=p,
                                      eps=self.eps, attention_type='fourier' if fourier else 'galerkin',
                                      self_attn=self_attn, mask_p=mask_p, quadratic=quadratic,
                                      want_attn=self.materialize_attn)
        # Galerkin: the (B,H,d,d) matrix K^T V / n (post-dropout), as the reference returns.
        # Fourier: the reference keeps the (B,H,n,n) matrix alive on the module; here it only exists when
        # `materialize_attn` is set (SimpleTransformerEncoderLayer(attn_weight=True) sets it).
        self.attn_weight = attn if (not fourier or self.materialize_attn) else None
        return x, self.attn_weight


class FeedForward(nn.Module):
    """Linear -> activation -> dropout -> Linear (libs/layers.py:954-987); dropout is fused into the
    first GEMM's epilogue (Philox, regenerated in backward)."""

    def __init__(self, in_dim=256, dim_feedforward: int = 1024, out_dim=None, batch_norm=False,
                 activation='relu', dropout=0.1):
        super().__init__()
        if batch_norm:
            raise NotImplementedError("FeedForward(batch_norm=True) is not supported (every shipped config "
                                      "sets batch_norm: False)")
        if activation == 'gelu':
            raise NotImplementedError("FeedForward(activation='gelu')")
        out_dim = in_dim if out_dim is None else out_dim
        self.lr1 = nn.Linear(in_dim, dim_feedforward)
        self.activation = nn.SiLU() if activation == 'silu' else nn.ReLU()
        self.act_name = 'silu' if activation == 'silu' else 'relu'
        self.batch_norm = batch_norm
        self.lr2 = nn.Linear(dim_feedforward, out_dim)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, residual=None, rscale=1.0, out_drop_p=0.0):
        p = self.dropout.p if self.training else 0.0
        if residual is None or residual is x:     # one autograd node: gated backward GEMM, shortcut gradient fused
            return GF.mlp2(x, self.lr1.weight, self.lr1.bias, self.lr2.weight, self.lr2.bias, act=self.act_name,
                           drop_p1=p, drop_p2=out_drop_p, rscale=rscale, shortcut=residual is not None)
        h = GF.linear(x, self.lr1.weight, self.lr1.bias, act=self.act_name, drop_p=p)
        return GF.linear(h, self.lr2.weight, self.lr2.bias, residual=residual, rscale=rscale,
                         drop_p=out_drop_p)


class _SpectralBase(nn.Module):
    def _act_name(self):
        a = self.activation
        if isinstance(a, nn.SiLU):
            return 'silu'
        if isinstance(a, nn.ReLU):
            return 'relu'
        if isinstance(a, (nn.Identity, Identity)):
            return 'none'
        raise NotImplementedError(f"SpectralConv activation {type(a).__name__}")

    def _dropped(self, x):
        # the reference drops the FFT input but takes the residual from the un-dropped x
        # (libs/layers.py:1083-1084, 1172-1173); every shipped config has decoder_dropout 0, so
        # this elementwise pass is normally skipped
        if self.training and self.dropout.p > 0:
            return torch.nn.functional.dropout(x, self.dropout.p, True)
        return None


class SpectralConv1d(_SpectralBase):
    def __init__(self, in_dim, out_dim, modes: int, n_grid=None, dropout=0.1, return_freq=False,
                 activation='silu', debug=False):
        super().__init__()
        self.linear = nn.Linear(in_dim, out_dim)
        self.modes = modes
        activation = 'silu' if activation is None else activation
        self.activation = nn.SiLU() if activation == 'silu' else nn.ReLU()
        self.n_grid = n_grid
        self.fourier_weight = nn.Parameter(torch.empty(in_dim, out_dim, modes, 2))
        nn.init.xavier_normal_(self.fourier_weight, gain=1 / (in_dim * out_dim))
        self.dropout = nn.Dropout(dropout)
        self.return_freq = return_freq
        self.debug = debug

    def forward(self, x):
        """(B, n, in_dim) -> (B, n, out_dim)"""
        n = x.size(1)
        y, of = GF.spectral_conv(x, self.linear.weight, self.linear.bias, self.fourier_weight, None,
                                 modes=self.modes, act=self._act_name(), two_d=False,
                                 x_transform=self._dropped(x))
        if self.return_freq:
            out_ft = torch.zeros(x.size(0), self.linear.out_features, n // 2 + 1, dtype=torch.complex64,
                                 device=x.device)
            out_ft[:, :, :self.modes] = torch.view_as_complex(of).permute(0, 2, 1)
            return y, out_ft
        return y


class SpectralConv2d(_SpectralBase):
    def __init__(self, in_dim, out_dim, modes: int, n_grid=None, dropout=0.1, norm='ortho',
                 activation='silu', return_freq=False, debug=False):
        super().__init__()
        if norm != 'ortho':
            raise NotImplementedError("SpectralConv2d: only norm='ortho'")
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.linear = nn.Linear(in_dim, out_dim)
        self.modes = modes
        activation = 'silu' if activation is None else activation
        self.activation = nn.SiLU() if activation == 'silu' else nn.ReLU()
        self.n_grid = n_grid
        self.fourier_weight = nn.ParameterList(
            [nn.Parameter(torch.empty(in_dim, out_dim, modes, modes, 2)) for _ in range(2)])
        for w in self.fourier_weight:
            nn.init.xavier_normal_(w, gain=1 / (in_dim * out_dim) * math.sqrt(in_dim + out_dim))
        self.dropout = nn.Dropout(dropout)
        self.norm = norm
        self.return_freq = return_freq
        self.debug = debug

    def forward(self, x):
        """(B, n*n, in_dim) or (B, n, n, in_dim) -> same layout with out_dim channels"""
        bsz, n_dim = x.size(0), x.ndim
        if n_dim == 4:
            n = x.size(1)
            assert x.size(1) == x.size(2)
        elif n_dim == 3:
            n = int(x.size(1) ** 0.5)
        else:
            raise ValueError("Dimension not implemented")
        x = x.reshape(-1, n, n, self.in_dim)
        m = self.modes
        y, of = GF.spectral_conv(x, self.linear.weight, self.linear.bias, self.fourier_weight[0],
                                 self.fourier_weight[1], modes=m, act=self._act_name(), two_d=True,
                                 x_transform=self._dropped(x))
        if n_dim == 3:
            y = y.reshape(bsz, n * n, self.out_dim)
        if self.return_freq:
            blk = torch.view_as_complex(of).view(bsz, 2 * m, m, self.out_dim).permute(0, 3, 1, 2)
            out_ft = torch.zeros(bsz, self.out_dim, n, n // 2 + 1, dtype=torch.complex64, device=x.device)
            out_ft[:, :, :m, :m] = blk[:, :, :m]
            out_ft[:, :, -m:, :m] = blk[:, :, m:]
            return y, out_ft
        return y
