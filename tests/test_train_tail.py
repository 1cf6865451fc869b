"""Training-step tail: loss + gradient, one-cycle schedule, clip + Adam.

CPU: oracle/train_oracle.py against fixtures recorded from the reference's WeightedL2Loss2d (tests/golden/train/) and
against torch.optim.Adam + OneCycleLR + clip_grad_norm_ run live.  GPU: csrc/train.cu through the C ABI against the
oracle and the same fixtures."""
import glob
import math
import os
import sys

import pytest
import torch

from helpers import rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import train_oracle as TO                                          # noqa: E402

TRAIN_GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "train", "*.pt")))
assert TRAIN_GOLDEN, "tests/golden/train/*.pt missing (tests/golden/make_golden_train.py)"


def _fix(path):
    return torch.load(path, weights_only=False)


def _oracle_on(fix, dtype=torch.float32):
    c = fix["case"]
    cast = lambda t: None if t is None else t.to(dtype)                        # noqa: E731
    K = None if fix["K"] is None else fix["K"][..., 0]
    return TO.weighted_l2_loss2d(cast(fix["preds"]), cast(fix["targets"]), cast(fix["targets_prime"]), cast(K), h=fix["h"],
                                 gamma=c["gamma"], dilation=c["dilation"], regularizer=c["regularizer"],
                                 return_norm=c["return_norm"])


@pytest.mark.parametrize("path", TRAIN_GOLDEN, ids=lambda p: os.path.basename(p)[:-3])
def test_loss_oracle_matches_reference_fixture(path):
    fix = _fix(path)
    o = _oracle_on(fix)
    assert abs(float(o["loss"]) - float(fix["loss"])) <= 2e-6 * abs(float(fix["loss"]))
    assert abs(float(o["reg"]) - float(fix["reg"])) <= 2e-6 * max(abs(float(fix["reg"])), 1e-30)
    assert abs(float(o["metric"]) - fix["metric"]) <= 2e-6 * fix["metric"]
    assert rel_l2(o["dloss"], fix["dloss"]) <= 5e-6
    if float(fix["reg"]) > 0:
        assert rel_l2(o["dreg"], fix["dreg"]) <= 5e-6


def _torch_trajectory(p0, grads, total_steps, max_lr, max_norm, wd=0.0):
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p], lr=max_lr, weight_decay=wd)
    sch = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=max_lr, div_factor=1e4, final_div_factor=1e4, pct_start=0.3,
                                              total_steps=total_steps)
    traj, hyp = [], []
    for g in grads:
        p.grad = g.clone()
        hyp.append((opt.param_groups[0]["lr"], opt.param_groups[0]["betas"][0]))
        torch.nn.utils.clip_grad_norm_([p], max_norm)
        opt.step()
        sch.step()
        traj.append(p.detach().clone())
    return traj, hyp


def test_one_cycle_and_adam_oracle_match_torch():
    gen = torch.Generator().manual_seed(7)
    n, steps = 1000, 9
    p0 = torch.randn(n, generator=gen, dtype=torch.float64)
    grads = [torch.randn(n, generator=gen, dtype=torch.float64) * (0.02 if i % 2 else 3.0) for i in range(steps)]
    traj, hyp = _torch_trajectory(p0, grads, total_steps=12, max_lr=1e-3, max_norm=0.99)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for k, g in enumerate(grads):
        lr, b1 = TO.one_cycle(k, 12, 1e-3, div_factor=1e4, final_div_factor=1e4, pct_start=0.3)
        assert abs(lr - hyp[k][0]) <= 1e-12 * max(1.0, hyp[k][0]) + 1e-18 and abs(b1 - hyp[k][1]) <= 1e-12
        p, m, v, _ = TO.clip_adam_step(p, g, m, v, k, lr, b1, max_norm=0.99)
        assert rel_l2(p, traj[k]) <= 1e-12


def test_host_one_cycle_matches_oracle():
    from galerkin_transformer_b200.train import one_cycle
    for total in (10, 37, 1000):
        for k in range(total):
            a = one_cycle(k, total, 1e-3, div_factor=1e4, final_div_factor=1e4, pct_start=0.3)
            b = TO.one_cycle(k, total, 1e-3, div_factor=1e4, final_div_factor=1e4, pct_start=0.3)
            assert abs(a[0] - b[0]) <= 1e-15 + 1e-12 * b[0] and abs(a[1] - b[1]) <= 1e-12


# ------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("path", TRAIN_GOLDEN, ids=lambda p: os.path.basename(p)[:-3])
def test_gpu_loss_matches_reference_fixture(path):
    from galerkin_transformer_b200.train import WeightedL2Loss2d
    fix = _fix(path)
    c = fix["case"]
    dev = torch.device("cuda", 0)
    lf = WeightedL2Loss2d(regularizer=c["regularizer"], h=fix["h"], gamma=c["gamma"], dilation=c["dilation"],
                          return_norm=c["return_norm"])
    preds = fix["preds"].to(dev).requires_grad_(True)
    tp = None if fix["targets_prime"] is None else fix["targets_prime"].to(dev)
    K = None if fix["K"] is None else fix["K"].to(dev)
    loss, reg, metric, _ = lf(preds, fix["targets"].to(dev), targets_prime=tp, K=K)
    assert abs(float(loss) - float(fix["loss"])) <= 5e-6 * abs(float(fix["loss"]))
    assert abs(metric - fix["metric"]) <= 5e-6 * fix["metric"]
    dloss, = torch.autograd.grad(loss, preds, retain_graph=True)
    assert rel_l2(dloss.cpu(), fix["dloss"]) <= 1e-5
    if float(fix["reg"]) > 0:
        assert abs(float(reg) - float(fix["reg"])) <= 5e-6 * float(fix["reg"])
        dreg, = torch.autograd.grad(reg, preds, retain_graph=True)
        assert rel_l2(dreg.cpu(), fix["dreg"]) <= 1e-5
        dtot, = torch.autograd.grad(loss + reg, preds)
        assert rel_l2(dtot.cpu(), fix["dloss"] + fix["dreg"]) <= 1e-5
    out4, dp = lf.loss_and_grad(preds.detach(), fix["targets"].to(dev), targets_prime=tp, K=K)
    assert rel_l2(dp.cpu(), fix["dloss"] + fix["dreg"]) <= 1e-5
    assert abs(float(out4[3]) - float(fix["loss"] + fix["reg"])) <= 5e-6 * float(fix["loss"] + fix["reg"])


@pytest.mark.gpu
def test_gpu_loss_at_darcy_size_matches_oracle():
    """C3's loss call: B=8, 141x141, K and target gradients given, regulariser on (examples/ex2_darcy.py:118)."""
    from galerkin_transformer_b200.train import WeightedL2Loss2d
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(11)
    B, n = 8, 141
    t = torch.randn(B, n, n, generator=g)
    p = t + 0.1 * torch.randn(B, n, n, generator=g)
    tp = 5.0 * torch.randn(B, n, n, 2, generator=g)
    K = 0.5 + torch.rand(B, n, n, generator=g)
    lf = WeightedL2Loss2d(regularizer=True, h=1.0 / n, gamma=0.5)
    out4, dp = lf.loss_and_grad(p.to(dev), t.to(dev), targets_prime=tp.to(dev), K=K.to(dev))
    o = TO.weighted_l2_loss2d(p.double(), t.double(), tp.double(), K.double(), h=1.0 / n, gamma=0.5, regularizer=True)
    assert abs(float(out4[0]) - float(o["loss"])) <= 2e-6 * float(o["loss"])
    assert abs(float(out4[1]) - float(o["reg"])) <= 2e-6 * float(o["reg"])
    assert rel_l2(dp.cpu(), o["dloss"] + o["dreg"]) <= 2e-6
    out4b, dpb = lf.loss_and_grad(p.to(dev), t.to(dev), targets_prime=tp.to(dev), K=K.to(dev))
    assert torch.equal(dp, dpb) and torch.equal(out4, out4b)          # fixed-order reductions


@pytest.mark.gpu
@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_gpu_fused_adam_matches_torch(wd):
    """FusedAdam (flat buffers, clip + Adam in two launches, one-cycle lr / beta1 from the device array) against
    torch.optim.Adam + clip_grad_norm_ + OneCycleLR on the same gradients."""
    from galerkin_transformer_b200.train import FusedAdam
    dev = torch.device("cuda", 0)
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Linear(37, 50), torch.nn.SiLU(), torch.nn.Linear(50, 3)).to(dev)
    ref = torch.nn.Sequential(torch.nn.Linear(37, 50), torch.nn.SiLU(), torch.nn.Linear(50, 3)).to(dev)
    ref.load_state_dict(net.state_dict())
    total = 12
    opt = FusedAdam(net, lr=1e-3, weight_decay=wd, max_grad_norm=0.99,
                    one_cycle=dict(total_steps=total, div_factor=1e4, final_div_factor=1e4, pct_start=0.3))
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=wd)
    rsch = torch.optim.lr_scheduler.OneCycleLR(ropt, max_lr=1e-3, div_factor=1e4, final_div_factor=1e4, pct_start=0.3,
                                               total_steps=total)
    x = torch.randn(64, 37, device=dev)
    for k in range(9):
        scale = 30.0 if k % 2 == 0 else 0.01                              # clip active / inactive
        for m, o in ((net, opt), (ref, ropt)):
            o.zero_grad()
            (m(x).pow(2).mean() * scale).backward()
        rnorm = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.99)
        ropt.step()
        rsch.step()
        opt.step()
        assert abs(float(opt.grad_norm) - float(rnorm)) <= 1e-5 * float(rnorm)
        for a, b in zip(net.parameters(), ref.parameters()):
            assert rel_l2(a.detach(), b.detach()) <= 2e-6, (k, rel_l2(a.detach(), b.detach()))
    for a in net.parameters():                                            # parameters are views of the flat buffer
        assert a.data_ptr() >= opt.flat_param.data_ptr()


@pytest.mark.gpu
def test_gpu_train_batch_darcy_fused_step_runs_and_learns():
    """The reference's train_batch_darcy signature over the fused tail: a small FourierTransformer2D overfits one batch."""
    import galerkin_transformer_b200 as G
    from galerkin_transformer_b200.train import FusedAdam, WeightedL2Loss2d, train_batch_darcy
    from helpers import load_golden
    dev = torch.device("cuda", 0)
    G.set_precision("x3")
    fix = load_golden("model_ft2d_darcy_small")
    torch.manual_seed(5)
    cfg = dict(fix["config"])
    for k in ("dropout", "downscaler_dropout", "upscaler_dropout", "ffn_dropout", "encoder_dropout", "decoder_dropout"):
        cfg[k] = 0.0
    model = G.FourierTransformer2D(**cfg).to(dev)
    G.set_attn_dropout(model, "off")         # a deterministic descent: the assertion below is about the update, not the noise
    model.train()
    node, pos, grid = (fix["inputs"][k].to(dev) for k in ("node", "pos", "grid"))
    B, n = node.shape[0], node.shape[1]
    with torch.no_grad():
        shape = model(node, None, pos=pos, grid=grid)["preds"].shape
    xs = torch.linspace(0, 1, n)
    u = torch.sin(math.pi * xs)[:, None] * torch.sin(2 * math.pi * xs)[None, :]
    ux = math.pi * torch.cos(math.pi * xs)[:, None] * torch.sin(2 * math.pi * xs)[None, :]
    uy = 2 * math.pi * torch.sin(math.pi * xs)[:, None] * torch.cos(2 * math.pi * xs)[None, :]
    data = dict(coeff=torch.ones(B, n, n, 1), node=node, edge=torch.zeros(1), pos=pos, grid=grid,
                target=u.expand(B, n, n)[..., None].contiguous().reshape(shape),
                target_grad=torch.stack([ux, uy], -1).expand(B, n, n, 2).contiguous())
    lf = WeightedL2Loss2d(regularizer=True, h=1.0 / n, gamma=0.1)
    opt = FusedAdam(model, lr=5e-3, one_cycle=dict(total_steps=100, div_factor=10.0, final_div_factor=10.0, pct_start=0.3))
    p0 = opt.flat_param.clone()
    first = last = None
    for _ in range(60):
        (loss, reg), u_pred, _ = train_batch_darcy(model, lf, data, opt, None, dev, grad_clip=0.99)
        first = float(loss) if first is None else first
        last = float(loss)
    assert math.isfinite(last) and last < first, (first, last)
    assert opt.step_count == 60 and math.isfinite(float(opt.grad_norm)) and not torch.equal(p0, opt.flat_param)
