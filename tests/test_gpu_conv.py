"""GPU parity of the scaler convolution block (csrc/conv.cu): y = act(dropout(conv3x3(x))), channel-last, against
torch.nn.functional.conv2d in fp64 -- the channel counts and grids of the BASELINE C3 scalers (1->128 on 141^2,
128->42, 42->42, 42->44, 128->128 on 77^2), odd sizes, ReLU and SiLU, forward / input gradient / weight gradient.
Tolerances: bf16x3 forward and input gradient 5e-5; weight gradient (library TF32 contraction over all pixels) 2e-3."""
import pytest
import torch
import torch.nn.functional as F

import galerkin_transformer_b200 as G
from galerkin_transformer_b200 import functional as GF
from helpers import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _x3():
    G.set_precision("x3")
    yield
    G.set_precision("x3")


def _ref(x, w, act):
    z = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), None, padding=1).permute(0, 2, 3, 1)
    return {"relu": torch.relu, "silu": F.silu}[act](z)


@pytest.mark.parametrize("B,H,W,Cin,Cout,act", [
    (2, 20, 23, 128, 42, "relu"), (1, 77, 77, 42, 42, "relu"), (2, 9, 17, 42, 44, "relu"), (2, 33, 16, 128, 128, "silu"),
    (1, 8, 16, 64, 48, "silu"), (2, 141, 141, 1, 128, "relu"), (3, 7, 5, 1, 128, "relu"), (8, 77, 77, 128, 128, "silu")])
def test_conv_block_forward_backward(B, H, W, Cin, Cout, act):
    g = torch.Generator(device=DEV).manual_seed(B * 100 + Cin)
    x = torch.randn(B, H, W, Cin, device=DEV, generator=g, requires_grad=True)
    w = (torch.randn(Cout, Cin, 3, 3, device=DEV, generator=g) / (3.0 * Cin ** 0.5)).requires_grad_(True)
    y = GF.conv3x3_block(x, w, act=act, drop_p=0.0)
    cot = torch.randn(y.shape, device=DEV, generator=g)
    dx, dw = torch.autograd.grad((y * cot).sum(), [x, w])
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    yr = _ref(xd, wd, act)
    dxr, dwr = torch.autograd.grad((yr * cot.double()).sum(), [xd, wd])
    assert rel_l2(y, yr) < 5e-5
    assert rel_l2(dx, dxr) < (5e-5 if act == "silu" or Cin == 1 else 2e-3)   # ReLU: a few gates flip between evaluation orders
    assert rel_l2(dw, dwr) < 2e-3


def test_conv_block_dropout_mask_is_consistent():
    """dropout inside the block: keep fraction, inverted scale, and backward uses the same mask"""
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn(2, 30, 31, 128, device=DEV, generator=g, requires_grad=True)
    w = (torch.randn(42, 128, 3, 3, device=DEV, generator=g) / 34.0).requires_grad_(True)
    p = 0.25
    GF._seed_counter = 0
    y = GF.conv3x3_block(x, w, act="relu", drop_p=p)
    dense = _ref(x.detach(), w.detach(), "relu")
    pos = dense > 1e-4
    kept = (y != 0) & pos
    assert abs(kept.sum().item() / pos.sum().item() - (1 - p)) < 0.01
    scale = 65536.0 / (65536.0 - round(p * 65536.0))
    assert rel_l2(y[kept], dense[kept] * scale) < 5e-5
    cot = torch.randn(y.shape, device=DEV, generator=g)
    dx, = torch.autograd.grad((y * cot).sum(), [x])
    gz = (cot.double() * (y != 0).double() * scale).permute(0, 3, 1, 2)     # the mask the forward actually drew
    dxr = F.conv_transpose2d(gz, w.detach().double(), padding=1).permute(0, 2, 3, 1)
    assert rel_l2(dx, dxr) < 1e-3
