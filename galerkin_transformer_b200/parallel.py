"""Data parallelism for batches of independent PDE grids: one process per GPU, weights
replicated, the batch axis sharded, and ONE all-reduce of a single flat fp32 gradient bucket per
step (NCCL over NVLink/NVSwitch; gloo in the CPU tests).  The reference has no distributed code
at all; this is the only collective the path needs because every operator is
per-sample (no BatchNorm, per-token LayerNorm, per-(b,h) K^T V, per-(b,c) transforms).

Gradients are accumulated by autograd directly into views of the flat bucket, so there is no
pack/unpack copy around the collective; averaging (1/G) happens inside the collective (ncclAvg)
so a following clip_grad_norm_ sees the same global norm as a single-GPU run on the full batch
(reference order: utils_ft.py:676-681).
"""
import torch
import torch.distributed as dist


class FlatGradBucket:
    def __init__(self, module: torch.nn.Module, process_group=None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = process_group
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)     # autograd accumulates in place
            off += p.numel()
        self.nbytes = total * 4

    def zero(self):
        self.flat.zero_()

    def world_size(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def all_reduce(self, async_op=False):
        """Average the bucket over ranks; a no-op for a single process."""
        if not dist.is_initialized() or self.world_size() == 1:
            return None
        if dist.get_backend(self.group) == "nccl":
            return dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        if async_op:
            work.wait()
        self.flat.div_(self.world_size())
        return None


def shard_batch(tensors, rank, world_size):
    """Contiguous equal shards of the leading (batch) axis."""
    out = []
    for t in tensors:
        b = t.shape[0]
        assert b % world_size == 0, "global batch must divide evenly over ranks"
        s = b // world_size
        out.append(t[rank * s:(rank + 1) * s])
    return out
