"""CUDA-graph capture of a whole training step (forward + loss + backward).

The hot path launches several hundred small kernels per step; on a B200 the GPU finishes them
faster than one Python thread can enqueue them.  `GraphedStep` records the step once and replays it
with a single `cudaGraphLaunch`, the B200-native answer to what the reference leaves to eager
dispatch (there is no tracing compiler in the loop: the graph is the literal kernel sequence).

Everything in the step is capture-safe: the library never allocates or synchronises, workspaces
come from PyTorch's graph-private pool, TMA descriptors are plain kernel arguments, and
fused-dropout seeds are offset by a device-side counter that a captured kernel bumps on every
replay (fresh masks per step; torch's own dropout / rand use its graph-safe Philox state).
"""
import torch

from . import functional as GF


class GraphedStep:
    """step_fn(*static_inputs) -> scalar loss; gradients land in `p.grad` of `params`.

        graphed = GraphedStep(lambda node, pos, grid, y: ((model(node, None, pos, grid)['preds'] - y) ** 2).mean(),
                              example_inputs, model.parameters())
        loss = graphed(node, pos, grid, y)      # copies inputs into the static buffers, replays
    """

    def __init__(self, step_fn, example_inputs, params, warmup=3):
        self.params = [p for p in params if p.requires_grad]
        self.static_inputs = [t.clone() for t in example_inputs]
        self.step_fn = step_fn
        GF.rng_step_counter(self.static_inputs[0].device)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        for p in self.params:
            p.grad = None
        with torch.cuda.graph(self.graph):
            GF.advance_rng()
            self.static_loss = self.step_fn(*self.static_inputs)
            self.static_loss.backward()
        self.static_grads = [p.grad for p in self.params]

    def _eager_step(self):
        for p in self.params:
            p.grad = None
        GF.advance_rng()
        loss = self.step_fn(*self.static_inputs)
        loss.backward()
        return loss

    def load_inputs(self, *inputs, non_blocking=True):
        for dst, src in zip(self.static_inputs, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=non_blocking)

    def replay(self):
        self.graph.replay()
        return self.static_loss

    def __call__(self, *inputs):
        self.load_inputs(*inputs)
        return self.replay()
