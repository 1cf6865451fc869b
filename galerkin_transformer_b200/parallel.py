"""Data parallelism for batches of independent PDE grids: one process per GPU, weights
replicated, the batch axis sharded, and ONE all-reduce of a single flat fp32 gradient bucket per
step (NCCL over NVLink/NVSwitch; gloo in the CPU tests).  The reference has no distributed code
at all; this is the only collective the path needs because every operator is
per-sample (no BatchNorm, per-token LayerNorm, per-(b,h) K^T V, per-(b,c) transforms).

Gradients are packed into the bucket by one multi-tensor copy; averaging (1/G) happens inside
the collective (ncclAvg), before any clip_grad_norm_, so the clip sees the same global norm as a
single-GPU run on the full batch (reference order: utils_ft.py:676-681).
"""
import torch
import torch.distributed as dist


class FlatGradBucket:
    """Flat fp32 gradient bucket.

    During backward autograd hands each parameter its gradient tensor as-is (`p.grad is None`
    beforehand, so AccumulateGrad steals the tensor: no per-parameter add kernels).  `all_reduce`
    packs the ~170 gradients into the bucket with one multi-tensor copy, runs ONE collective, and
    re-points every `p.grad` at its slice of the averaged bucket (views, no unpack copy)."""

    def __init__(self, module: torch.nn.Module, process_group=None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = process_group
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.nbytes = total * 4

    def zero(self):
        """Drop last step's gradients (set_to_none semantics, no memset kernels)."""
        for p in self.params:
            p.grad = None

    def world_size(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def pack(self, grads=None):
        """Copy gradients (default: each `p.grad`) into the bucket and re-point `p.grad` at its slices.
        Pass `grads` explicitly when they live in static CUDA-graph buffers."""
        if grads is None:
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        torch._foreach_copy_(self.views, grads)
        for p, v in zip(self.params, self.views):
            p.grad = v

    def all_reduce(self, grads=None):
        """Average gradients over ranks (one collective); a no-op for a single process."""
        if not dist.is_initialized() or self.world_size() == 1:
            return
        self.pack(grads)
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(self.world_size())

    def all_reduce_packed(self):
        """The collective alone, for a bucket that something else already filled -- e.g. the last node of a captured CUDA
        graph (graphs.GraphedStep(post_backward=bucket.pack)): enqueue this right after `graph.replay()` and only the
        NVLink transfer itself stays on the critical path (no pack launch, no host gap)."""
        if not dist.is_initialized() or self.world_size() == 1:
            return
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(self.world_size())


def shard_batch(tensors, rank, world_size):
    """Contiguous equal shards of the leading (batch) axis."""
    out = []
    for t in tensors:
        b = t.shape[0]
        assert b % world_size == 0, "global batch must divide evenly over ranks"
        s = b // world_size
        out.append(t[rank * s:(rank + 1) * s])
    return out
