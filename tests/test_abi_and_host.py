"""CPU: the C-ABI library loads and exports every symbol include/*.h declares, the host-side
modules mirror the reference's interface (state_dict keys and shapes from the fixtures), and the
product path fails loudly -- never falls back -- without a CUDA device."""
import copy
import ctypes
import glob
import os
import pickle
import re

import pytest
import torch

import galerkin_transformer_b200 as G
from galerkin_transformer_b200 import _lib, build as B
from helpers import golden_names, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(gb200_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_builds_and_exports_every_declared_symbol():
    B.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().gb200_version() >= 100


def test_error_reporting_without_gpu():
    lib = _lib.load()
    rc = lib.gb200_gemm(0, None, 1, 0, None, 1, 0, None, 1, 4, 4, 4, 1, 0, 0, 0, 1.0, None, 0, None, 0, 0.0,
                        0, None, 0, 1.0, 0, 1, None, 0, None)
    assert rc != 0
    assert b"null operand" in lib.gb200_last_error()
    with pytest.raises(RuntimeError, match="null operand"):
        _lib.check(rc, "gb200_gemm")


def build_module(fix):
    name, cfg = fix["name"], fix["config"]
    for prefix, cls in (("attn_", G.SimpleAttention), ("enc_", G.SimpleTransformerEncoderLayer),
                        ("sc1d_", G.SpectralConv1d), ("sc2d_", G.SpectralConv2d),
                        ("model_ft2dlite_", G.FourierTransformer2DLite),
                        ("model_ft2d_", G.FourierTransformer2D), ("model_simple_", G.SimpleTransformer)):
        if name.startswith(prefix):
            return cls(**cfg)
    raise KeyError(name)


@pytest.mark.parametrize("name", golden_names())
def test_state_dict_layout_matches_reference(name):
    fix = load_golden(name)
    mod = build_module(fix)
    sd, ref = mod.state_dict(), fix["state_dict"]
    assert list(sd.keys()) == list(ref.keys())
    for k in sd:
        assert sd[k].shape == ref[k].shape, k
    mod.load_state_dict(ref)                       # reference checkpoints load unchanged
    clone = copy.deepcopy(mod)                     # model.py:896, 1153 deepcopy the layer
    blob = pickle.dumps(mod)                       # utils_ft.py:804 pickles whole models
    assert list(pickle.loads(blob).state_dict().keys()) == list(clone.state_dict().keys())


def test_no_cpu_fallback():
    attn = G.SimpleAttention(n_head=2, d_model=16, pos_dim=1, attention_type="galerkin", norm=True)
    x = torch.randn(2, 8, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        attn(x, x, x, pos=torch.rand(2, 8, 1))
    conv = G.SpectralConv2d(4, 4, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        conv(torch.randn(1, 8, 8, 4))


def test_unsupported_arguments_raise():
    with pytest.raises(NotImplementedError):
        G.SimpleAttention(n_head=2, d_model=16, attention_type="softmax")
    with pytest.raises(AssertionError):
        G.SimpleAttention(n_head=3, d_model=16)
    attn = G.SimpleAttention(n_head=2, d_model=16, attention_type="galerkin")
    x = torch.randn(1, 4, 16)
    with pytest.raises(RuntimeError, match="casual mask"):
        attn(x, x, x, mask=torch.ones(1, 4, 4))
    with pytest.raises(ValueError, match="Dimension not implemented"):
        G.SpectralConv2d(4, 4, 2)(torch.randn(8, 4))


def test_initialisation_statistics_follow_reference():
    torch.manual_seed(0)
    a = G.SimpleAttention(n_head=4, d_model=64, xavier_init=1e-2, diagonal_weight=1e-2, norm=True,
                          attention_type="galerkin", pos_dim=2)
    w = a.linears[0].weight
    off = w - torch.diag(torch.diag(w))
    bound = 1e-2 * (6 / 128) ** 0.5
    assert off.abs().max() <= bound + 1e-7
    assert abs(torch.diag(w).mean().item() - 1e-2) < 5e-3
    assert a.linears[1].bias.abs().max() == 0
    assert a.fc.weight.shape == (64, 64 + 4 * 2)
    s = G.SpectralConv2d(8, 8, 3)
    assert s.fourier_weight[0].shape == (8, 8, 3, 3, 2)


def test_scaler_sizes():
    assert G.scaler_sizes(141, 43) == ((0.555, 0.555), ((77, 77), (141, 141)))
    down, up = G.scaler_sizes(211, 71)
    assert up[1] == (211, 211)
    assert G.scaler_sizes(141, 43, scale_factor=False)[0] == ((77, 77), (43, 43))
