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

    def __init__(self, step_fn, example_inputs, params, warmup=3, batch_streams=1, post_backward=None):
        """batch_streams = k > 1: the static inputs are split into k equal chunks along dim 0 and step_fn runs on each
        chunk on its own stream (forward AND, because autograd replays a node on its forward stream, backward), the
        k losses are averaged.  For per-sample-independent models with a batch-mean loss (every model of this
        package: no BatchNorm, attention / spectral conv / resizes are per sample, libs/ft.py:1084) this is the same
        loss and the same gradients; inside the captured graph the k chains are parallel branches.  At C3 sizes every
        kernel is a single latency-bound wave, so two half-batch chains overlap almost freely."""
        self.params = [p for p in params if p.requires_grad]
        self.static_inputs = [t.clone() for t in example_inputs]
        self.step_fn = step_fn
        self.batch_streams = int(batch_streams)
        if self.batch_streams > 1:       # per-chain streams: weight-gradient side streams must be joined inside each chain
            GF._DEFER_WGRAD_JOIN = False
        if self.batch_streams > 1:
            for t in self.static_inputs:
                assert t.shape[0] % self.batch_streams == 0, "batch must divide evenly over batch_streams"
        self._chunk_streams = [torch.cuda.Stream() for _ in range(self.batch_streams - 1)]
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
            self.static_loss = self._loss()
            self.static_loss.backward()
            self.static_grads = [p.grad for p in self.params]
            if post_backward is not None:        # e.g. FlatGradBucket.pack: the gradient bucket is filled inside the graph
                post_backward(self.static_grads)
        self.static_grads = [p.grad for p in self.params]

    def _eager_step(self):
        for p in self.params:
            p.grad = None
        GF.advance_rng()
        loss = self._loss()
        loss.backward()
        return loss

    def _loss(self):
        k = self.batch_streams
        if k == 1:
            return self.step_fn(*self.static_inputs)
        cur = torch.cuda.current_stream()
        for s in self._chunk_streams:              # fork BEFORE anything of chunk 0 is enqueued: the chains are independent
            s.wait_stream(cur)
        chunks = [t.chunk(k, dim=0) for t in self.static_inputs]
        losses = []
        for i in range(k):
            s = cur if i == 0 else self._chunk_streams[i - 1]
            with torch.cuda.stream(s):
                losses.append(self.step_fn(*[c[i] for c in chunks]) / k)
        for s in self._chunk_streams:
            cur.wait_stream(s)
        total = losses[0]
        for l in losses[1:]:
            total = total + l
        return total

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
