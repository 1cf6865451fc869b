"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the training-step tail.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this; the product path never does.

  weighted_l2_loss2d  libs/ft.py:1035-1101 (WeightedL2Loss2d.forward with preds_prime=None, weights=None) and
                      libs/ft.py:1020-1034 (central_diff), plus the CLOSED-FORM gradients of the loss and of the
                      regulariser with respect to preds (what autograd through those lines produces)
  one_cycle           torch.optim.lr_scheduler.OneCycleLR as configured in examples/ex2_darcy.py:111-116
                      (cos annealing, two phases, cycle_momentum on Adam's beta1)
  clip_adam_step      nn.utils.clip_grad_norm_ + torch.optim.Adam.step, libs/utils_ft.py:676-681

Pinned by tests/test_train_tail.py against fixtures recorded from the reference loss class (tests/golden/train/, made by
tests/golden/make_golden_train.py) and against torch.optim.Adam / OneCycleLR run live."""
import math

import torch


def weighted_l2_loss2d(preds, targets, targets_prime=None, K=None, *, h, beta=1.0, gamma=0.1, eps=1e-10, dilation=2,
                       regularizer=False, return_norm=True):
    """preds, targets (B, n, n); targets_prime (B, n, n, 2) | None; K (B, n, n) | None.
    Returns dict(loss, reg, metric, dloss, dreg): scalars and the (B, n, n) gradients of loss / reg w.r.t. preds."""
    B, n, _ = preds.shape
    N = n * n
    s = dilation // 2
    Kf = torch.ones_like(preds) if K is None else K
    tnorm = targets.pow(2).mean(dim=(1, 2)) + eps                                      # ft.py:1054
    if targets_prime is not None:
        tpnorm = 2 * (Kf[..., None] * targets_prime.pow(2)).mean(dim=(1, 2, 3)) + eps   # ft.py:1056-1058 (d = 2)
    else:
        tpnorm = torch.ones(B, dtype=preds.dtype)
    lb = beta * (preds - targets).pow(2).mean(dim=(1, 2)) / tnorm                      # ft.py:1062-1063
    metric = lb.sqrt().mean()                                                          # ft.py:1075-1076 ('L1')
    loss = lb.sqrt().mean() if return_norm else lb.mean()                              # ft.py:1080
    fl = 0.5 / lb.sqrt() if return_norm else torch.ones_like(lb)
    dloss = (fl * beta * 2.0 / (N * tnorm) / B)[:, None, None] * (preds - targets)
    reg = torch.zeros((), dtype=preds.dtype)
    dreg = torch.zeros_like(preds)
    if regularizer and targets_prime is not None:
        inv = 1.0 / (dilation * h)
        dx = (preds[:, dilation:, s:-s] - preds[:, :-dilation, s:-s]) * inv            # ft.py:1029-1032
        dy = (preds[:, s:-s, dilation:] - preds[:, s:-s, :-dilation]) * inv
        tp = targets_prime[:, s:-s, s:-s, :]
        Ki = Kf[:, s:-s, s:-s]
        rx, ry = Ki * (tp[..., 0] - dx), Ki * (tp[..., 1] - dy)
        ni2 = 2.0 * (n - 2 * s) ** 2
        rb = gamma * h * (rx.pow(2) + ry.pow(2)).sum(dim=(1, 2)) / ni2 / tpnorm       # ft.py:1090-1091
        reg = rb.sqrt().mean() if return_norm else rb.mean()                           # ft.py:1093
        fr = 0.5 / rb.sqrt() if return_norm else torch.ones_like(rb)
        c = (fr * gamma * h * 2.0 / (ni2 * tpnorm) * inv / B)[:, None, None]
        gx, gy = Ki * Ki * (dx - tp[..., 0]) * c, Ki * Ki * (dy - tp[..., 1]) * c      # d rb / d (dx, dy), scaled
        dreg[:, dilation:, s:-s] += gx
        dreg[:, :-dilation, s:-s] -= gx
        dreg[:, s:-s, dilation:] += gy
        dreg[:, s:-s, :-dilation] -= gy
    return dict(loss=loss, reg=reg, metric=metric, dloss=dloss, dreg=dreg)


def one_cycle(step, total_steps, max_lr, div_factor=25.0, final_div_factor=1e4, pct_start=0.3, base_momentum=0.85,
              max_momentum=0.95):
    """(lr, beta1) in force when optimizer.step() number `step` (0-based) runs."""
    initial, low = max_lr / div_factor, max_lr / div_factor / final_div_factor
    e1, e2 = float(pct_start * total_steps) - 1.0, float(total_steps - 1)
    if step <= e1:
        a, b, ma, mb, pct = initial, max_lr, max_momentum, base_momentum, step / e1
    else:
        a, b, ma, mb, pct = max_lr, low, base_momentum, max_momentum, (step - e1) / (e2 - e1)
    w = (math.cos(math.pi * pct) + 1.0) / 2.0
    return b + (a - b) * w, mb + (ma - mb) * w


def clip_adam_step(p, g, m, v, k, lr, beta1, beta2=0.999, eps=1e-8, weight_decay=0.0, max_norm=0.0):
    """One clip_grad_norm_ + Adam step (the (k+1)-th) on flat tensors; returns (p, m, v, total_norm)."""
    total = g.pow(2).sum().sqrt()
    if max_norm > 0:
        g = g * torch.clamp(max_norm / (total + 1e-6), max=1.0)
    if weight_decay != 0.0:
        g = g + weight_decay * p
    m = m + (g - m) * (1.0 - beta1)
    v = v * beta2 + (1.0 - beta2) * g * g
    t = k + 1
    bc1, bc2 = 1.0 - beta1 ** t, 1.0 - beta2 ** t
    p = p - (lr / bc1) * m / (v.sqrt() / math.sqrt(bc2) + eps)
    return p, m, v, total
