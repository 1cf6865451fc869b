"""Tail of the reference's training step on the device, without host round trips.

Mirrors, for the path `train_batch_darcy` (libs/utils_ft.py:656-690) takes:

    loss, reg, _, _ = loss_func(u_pred, target, targets_prime=gradu, K=a)     # WeightedL2Loss2d, libs/ft.py:983-1105
    (loss + reg).backward()
    nn.utils.clip_grad_norm_(model.parameters(), 0.99)
    optimizer.step()                                                          # torch.optim.Adam
    lr_scheduler.step()                                                       # OneCycleLR (also cycles Adam's beta1)

* `WeightedL2Loss2d` -- same constructor and `forward` signature as the reference class; the loss, the regulariser and
  both gradients come from two launches (`gb200_weighted_l2_loss2d`).  `forward` returns the reference's 4-tuple
  (`metric` is a Python float, i.e. one host sync, exactly like the reference's `.item()`); `loss_and_grad` is the
  sync-free form the fused step uses.
* `FusedAdam` -- all parameters, gradients and Adam moments live in flat fp32 buffers (the gradient buffer is the
  data-parallel bucket of `parallel.FlatGradBucket`); `step()` = global-norm clip + Adam in two launches
  (`gb200_adam_clip_step`), the one-cycle schedule (lr and beta1) evaluated on the host and handed over through a 4-float
  device array, so the whole step can be captured in a CUDA graph.
* `one_cycle` -- torch.optim.lr_scheduler.OneCycleLR (two-phase cosine, cycle_momentum) as a pure function of the step.
* `train_batch_darcy` -- the reference function's signature over the fused pieces.

Unsupported arguments raise NotImplementedError (no fallback to eager PyTorch)."""
import math

import torch
from torch import nn

from . import _lib
from ._lib import ptr, stream_of, workspace
from .functional import _dev, _launch, require_cuda_f32


# --------------------------------------------------------------------------------------------------------------------
# loss
# --------------------------------------------------------------------------------------------------------------------
def _loss_launch(preds, targets, targets_prime, K, cfg, want_grad):
    lib = _lib.load()
    B, n, n2 = preds.shape
    assert n == n2, "square grids only (libs/ft.py central_diff)"
    h, beta, gamma, eps, dilation, regularizer, return_norm = cfg
    out4 = torch.empty(4, dtype=torch.float32, device=preds.device)
    dl = torch.empty_like(preds) if want_grad else None
    dr = torch.empty_like(preds) if (want_grad and regularizer and targets_prime is not None) else None
    wsb = lib.gb200_weighted_l2_loss2d_workspace_bytes(B, n)
    ws = workspace(wsb, preds)
    nbytes = 4.0 * preds.numel() * (2 + (2 if targets_prime is not None else 0) + (K is not None) + 2 * want_grad)
    _launch("loss_l2_h1", 30.0 * preds.numel(), nbytes, lib.gb200_weighted_l2_loss2d, _dev(preds), ptr(preds), ptr(targets),
            ptr(targets_prime), ptr(K), B, n, h, beta, gamma, eps, dilation, int(regularizer), int(return_norm), ptr(out4),
            ptr(dl), ptr(dr), ptr(ws), wsb, stream_of(preds))
    return out4, dl, dr


class _WeightedL2LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, preds, targets, targets_prime, K, cfg):
        out4, dl, dr = _loss_launch(preds, targets, targets_prime, K, cfg, preds.requires_grad)
        ctx.save_for_backward(dl, dr)
        loss, reg, metric = out4[0], out4[1], out4[2]
        ctx.mark_non_differentiable(metric)
        return loss, reg, metric

    @staticmethod
    def backward(ctx, gl, gr, _gm):
        dl, dr = ctx.saved_tensors
        g = dl * gl
        if dr is not None:
            g = torch.addcmul(g, dr, gr)
        return g, None, None, None, None


def _grid3(t, name):
    """(B, n, n) or (B, n, n, 1) -> contiguous (B, n, n)"""
    if t.ndim == 4 and t.shape[-1] == 1:
        t = t[..., 0]
    if t.ndim != 3:
        raise NotImplementedError(f"WeightedL2Loss2d: {name} must be (B, n, n) or (B, n, n, 1), got {tuple(t.shape)}")
    return t.contiguous()


class WeightedL2Loss2d(nn.Module):
    """libs/ft.py:983-1105, same constructor.  Supported: dim=2, noise=0, metric_reduction='L1', alpha=0 /
    preds_prime=None (the call `train_batch_darcy` makes for 3-D outputs), weights=None."""

    def __init__(self, dim=2, dilation=2, regularizer=False, h=1 / 421, beta=1.0, gamma=1e-1, alpha=0.0, delta=0.0,
                 metric_reduction='L1', return_norm=True, noise=0.0, eps=1e-10, debug=False):
        super().__init__()
        if dim != 2:
            raise NotImplementedError("WeightedL2Loss2d: dim != 2")
        if noise != 0.0:
            raise NotImplementedError("WeightedL2Loss2d: target noise")
        if metric_reduction != 'L1':
            raise NotImplementedError("WeightedL2Loss2d: metric_reduction other than 'L1'")
        assert dilation % 2 == 0
        self.noise, self.regularizer, self.dilation, self.dim, self.h = noise, regularizer, dilation, dim, h
        self.beta, self.gamma, self.alpha, self.delta = beta, gamma, alpha, delta * h ** dim
        self.eps, self.metric_reduction, self.return_norm, self.debug = eps, metric_reduction, return_norm, debug

    def _cfg(self):
        return (float(self.h), float(self.beta), float(self.gamma), float(self.eps), int(self.dilation),
                bool(self.regularizer), bool(self.return_norm))

    def _prepare(self, preds, targets, preds_prime, targets_prime, weights, K):
        if weights is not None:
            raise NotImplementedError("WeightedL2Loss2d: nonuniform mesh weights")
        if preds_prime is not None and self.alpha > 0:
            raise NotImplementedError("WeightedL2Loss2d: the alpha (predicted-gradient) term")
        preds, targets = _grid3(preds, "preds"), _grid3(targets, "targets")
        if targets_prime is not None:
            if targets_prime.ndim != 4 or targets_prime.shape[-1] != 2:
                raise NotImplementedError("WeightedL2Loss2d: targets_prime must be (B, n, n, 2)")
            targets_prime = targets_prime.contiguous()
        if K is not None:
            K = _grid3(K, "K")
        require_cuda_f32(preds, targets, targets_prime, K)
        return preds, targets, targets_prime, K

    def forward(self, preds, targets, preds_prime=None, targets_prime=None, weights=None, K=None):
        preds, targets, targets_prime, K = self._prepare(preds, targets, preds_prime, targets_prime, weights, K)
        loss, reg, metric = _WeightedL2LossFn.apply(preds, targets, targets_prime, K, self._cfg())
        if not (self.regularizer and targets_prime is not None):
            reg = torch.zeros(1, device=preds.device, requires_grad=True)         # libs/ft.py:1096-1098
        return loss, reg, metric.item(), dict(L2=None, H1=None)

    def loss_and_grad(self, preds, targets, targets_prime=None, K=None):
        """(out4, d(loss + reg)/d preds) with no autograd node and no host sync; out4 = device floats
        [loss, regularizer, metric, loss + regularizer]."""
        preds, targets, targets_prime, K = self._prepare(preds, targets, None, targets_prime, None, K)
        out4, dl, dr = _loss_launch(preds.detach(), targets, targets_prime, K, self._cfg(), True)
        if dr is not None:
            dl.add_(dr)
        return out4, dl


# --------------------------------------------------------------------------------------------------------------------
# one-cycle schedule + fused clip / Adam
# --------------------------------------------------------------------------------------------------------------------
def one_cycle(step, total_steps, max_lr, div_factor=25.0, final_div_factor=1e4, pct_start=0.3, base_momentum=0.85,
              max_momentum=0.95):
    """(lr, beta1) that torch.optim.lr_scheduler.OneCycleLR (anneal_strategy='cos', three_phase=False,
    cycle_momentum=True) has set when optimizer.step() number `step` (0-based) runs."""
    initial_lr = max_lr / div_factor
    min_lr = initial_lr / final_div_factor
    end1 = float(pct_start * total_steps) - 1.0
    end2 = float(total_steps - 1)

    def cos(start, end, pct):
        return end + (start - end) / 2.0 * (math.cos(math.pi * pct) + 1.0)

    if step <= end1:
        pct = step / end1
        return cos(initial_lr, max_lr, pct), cos(max_momentum, base_momentum, pct)
    pct = (step - end1) / (end2 - end1)
    return cos(max_lr, min_lr, pct), cos(base_momentum, max_momentum, pct)


class FusedAdam:
    """Adam + clip_grad_norm_ + OneCycleLR over flat buffers.

        opt = FusedAdam(model, lr=1e-3, max_grad_norm=0.99, one_cycle=dict(total_steps=..., div_factor=1e4,
                                                                            final_div_factor=1e4, pct_start=0.3))
        loss.backward(); opt.step()            # no zero_grad needed: gradients are re-pointed, not accumulated

    Every `p.data` becomes a view of `self.flat_param`; gradients are packed into `self.bucket.flat`
    (`parallel.FlatGradBucket`, so `opt.bucket.all_reduce()` is the data-parallel hook)."""

    def __init__(self, module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=0.0, one_cycle=None,
                 process_group=None):
        from .parallel import FlatGradBucket
        self.bucket = FlatGradBucket(module, process_group)
        params = self.bucket.params
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdam: parameters must live on a CUDA device (there is no CPU path)")
        self.n = self.bucket.flat.numel()
        self.flat_param = torch.empty(self.n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in params:
                v = self.flat_param[off:off + p.numel()].view_as(p)
                v.copy_(p.data)
                p.data = v
                off += p.numel()
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), tuple(betas), float(eps), float(weight_decay)
        self.max_grad_norm = float(max_grad_norm)
        self.one_cycle = dict(one_cycle) if one_cycle else None
        self.step_count = 0
        self.hyper = torch.zeros(4, dtype=torch.float32, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self._ws_bytes = _lib.load().gb200_adam_clip_step_workspace_bytes(self.n)
        self._ws = torch.empty(max(1, (self._ws_bytes + 3) // 4), dtype=torch.float32, device=dev)

    def hyper_for_step(self, k):
        """lr, beta1 and the two bias corrections torch.optim.Adam uses at its (k+1)-th step under the schedule"""
        if self.one_cycle:
            oc = dict(self.one_cycle)
            lr, beta1 = one_cycle(k, oc.pop("total_steps"), oc.pop("max_lr", self.lr), **oc)
        else:
            lr, beta1 = self.lr, self.betas[0]
        t = k + 1
        return lr, beta1, 1.0 - beta1 ** t, 1.0 - self.betas[1] ** t

    def set_hyper(self, k=None):
        """Stage step k's hyper-parameters in the device array (stream-ordered copy on the current stream)."""
        k = self.step_count if k is None else k
        # a fresh pageable source: the runtime stages it at call time, so the host may run steps ahead of the device
        self.hyper.copy_(torch.tensor(self.hyper_for_step(k), dtype=torch.float32))

    def launch(self, grads_packed=False):
        """clip + Adam on the flat buffers (graph-capturable: reads `self.hyper` from the device)."""
        if not grads_packed:
            self.bucket.pack()
        _launch("adam_clip", 12.0 * self.n, 28.0 * self.n, _lib.load().gb200_adam_clip_step, _dev(self.flat_param),
                ptr(self.flat_param), ptr(self.bucket.flat), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.n,
                ptr(self.hyper), self.betas[1], self.eps, self.weight_decay, self.max_grad_norm, ptr(self.grad_norm),
                ptr(self._ws), self._ws_bytes, stream_of(self.flat_param))

    def step(self, grads_packed=False):
        self.set_hyper()
        self.launch(grads_packed)
        self.step_count += 1

    def zero_grad(self, set_to_none=True):
        self.bucket.zero()

    def state_dict(self):
        return dict(step=self.step_count, exp_avg=self.exp_avg.clone(), exp_avg_sq=self.exp_avg_sq.clone())

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])


def train_batch_darcy(model, loss_func, data, optimizer, lr_scheduler=None, device=None, grad_clip=0.99):
    """libs/utils_ft.py:656-690 over the fused tail.  `optimizer` is a `FusedAdam` (its one-cycle schedule replaces
    `lr_scheduler`, which must be None); returns ((loss, reg) device tensors instead of Python floats -- no host sync --,
    u_pred, up_pred)."""
    if not isinstance(optimizer, FusedAdam) or not isinstance(loss_func, WeightedL2Loss2d):
        raise NotImplementedError("train_batch_darcy: needs galerkin_transformer_b200.train.FusedAdam / WeightedL2Loss2d")
    if lr_scheduler is not None:
        raise NotImplementedError("train_batch_darcy: pass the schedule to FusedAdam(one_cycle=...) instead")
    device = device if device is not None else optimizer.flat_param.device
    optimizer.zero_grad()
    optimizer.max_grad_norm = float(grad_clip)
    a, x, edge = data["coeff"].to(device), data["node"].to(device), data["edge"].to(device)
    pos, grid = data["pos"].to(device), data["grid"].to(device)
    u, gradu = data["target"].to(device), data["target_grad"].to(device)
    out_ = model(x, edge, pos=pos, grid=grid)
    out = out_["preds"] if isinstance(out_, dict) else out_[0]
    if out.ndim != 4 or out.shape[-1] != 1:
        raise NotImplementedError("train_batch_darcy: predicted-gradient outputs (the alpha term)")
    u_pred = out[..., 0]
    out4, dpred = loss_func.loss_and_grad(u_pred, u[..., 0], targets_prime=gradu, K=a)
    u_pred.backward(dpred)
    optimizer.bucket.all_reduce()
    optimizer.step(grads_packed=optimizer.bucket.world_size() > 1)
    return (out4[3], out4[1]), u_pred, u_pred
