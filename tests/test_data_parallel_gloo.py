"""CPU, world_size 2 over gloo: the flat-bucket gradient all-reduce reproduces the single-process
full-batch gradients (batch sharding + one averaged collective), on a stand-in torch module
(the CUDA operators themselves need a GPU; the bucket logic does not)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from galerkin_transformer_b200.parallel import FlatGradBucket, shard_batch


def _model():
    torch.manual_seed(5)
    return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.SiLU(), torch.nn.Linear(16, 3))


def _data():
    g = torch.Generator().manual_seed(11)
    return torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = _model()
    bucket = FlatGradBucket(model)
    x, y = shard_batch(_data(), rank, world)
    for _ in range(2):                       # second step checks zero() + in-place accumulation
        bucket.zero()
        ((model(x) - y) ** 2).mean().backward()
        bucket.all_reduce()
    q.put((rank, bucket.flat.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_matches_full_batch():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    model = _model()
    x, y = _data()
    ((model(x) - y) ** 2).mean().backward()
    full = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(got[0], got[1])
    assert torch.allclose(got[0], full, rtol=1e-5, atol=1e-7)


def test_bucket_pack_points_grads_at_flat_views():
    model = _model()
    bucket = FlatGradBucket(model)
    x, y = _data()
    bucket.zero()
    assert all(p.grad is None for p in model.parameters())
    ((model(x) - y) ** 2).mean().backward()
    ref = [p.grad.clone() for p in model.parameters()]
    bucket.all_reduce()                       # single process: no-op, grads untouched
    assert all(torch.equal(p.grad, r) for p, r in zip(model.parameters(), ref))
    bucket.pack()
    off = 0
    for p, r in zip(model.parameters(), ref):
        assert p.grad.data_ptr() == bucket.flat.data_ptr() + 4 * off
        assert torch.equal(p.grad, r)
        off += p.numel()
    assert bucket.nbytes == 4 * off
