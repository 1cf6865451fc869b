"""Kernel-time breakdown of one eager C3 training step (torch.profiler / CUPTI); diagnostics only, not a bench value.

    python tools/prof_step.py [--precision x3] [--steps 3] [--top 45]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                                      # noqa: E402
import galerkin_transformer_b200 as G                                             # noqa: E402
from galerkin_transformer_b200 import functional as GF                            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="x3")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--timeline", default="", help="write every kernel of the LAST profiled step, in start order, to this file")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    G.set_precision(args.precision)
    torch.manual_seed(1127802)
    model = G.FourierTransformer2D(**bench.c3_config()).to(dev)
    model.train()
    node, pos, grid, target = bench.c3_inputs(bench.BATCH, dev)

    def step():
        for p in model.parameters():
            p.grad = None
        GF.advance_rng()
        loss = ((model(node, None, pos, grid)["preds"] - target) ** 2).mean()
        loss.backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
    if args.timeline:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time > 0]
        evs.sort(key=lambda e: e.time_range.start)
        per = len(evs) // args.steps
        evs = evs[-per:]
        t0 = evs[0].time_range.start
        with open(args.timeline, "w") as f:
            for e in evs:
                f.write(f"{e.time_range.start - t0:10.1f} {e.device_time:8.1f}  {e.name[:100]}\n")
    rows = []
    total = 0.0
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = e.cuda_time_total
        if t > 0:
            rows.append((t / args.steps, e.count / args.steps, e.key))
            total += t / args.steps
    rows.sort(reverse=True)
    print(f"total kernel time per step: {total / 1e3:.3f} ms ({args.precision})")
    for t, c, k in rows[:args.top]:
        print(f"{t:9.1f} us {100 * t / total:5.1f}%  x{c:5.1f}  {k[:110]}")


if __name__ == "__main__":
    main()
