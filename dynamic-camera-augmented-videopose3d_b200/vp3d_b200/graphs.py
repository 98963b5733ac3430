"""CUDA-graph capture of a whole training step (forward -> loss -> backward -> optimiser).

One step of the 243-frame 1f model is ~140 kernel launches of 5-150 us each; issued one by one from Python the GPU
idles between them (~0.7 ms of a 3.5 ms step on a B200). Shapes are static in training (fixed batch, fixed window), so
the step is captured once with torch.cuda.graphs and replayed: the only per-step host work is one graph launch.
Everything the step launches is capturable: the C ABI launches on torch's current (capturing) stream, allocates
nothing, bakes only device pointers that live in the graph's private memory pool, and the dropout masks are keyed by a
device-side step counter that a captured kernel increments (vp3d_counter_add), so every replay draws new masks.
"""
import torch


class GraphedTrainStep:
    """step = GraphedTrainStep(model, optimizer, loss_fn, example_inputs, example_target [, preprocess])
       loss = step(inputs, target)     # tensors with the example's shapes; returns the (device) loss of this step

    `preprocess(*inputs) -> model input` runs inside the graph (e.g. the dynamic-camera projection world_to_image).
    The optimiser must be capturable (torch.optim.Adam(..., capturable=True))."""

    def __init__(self, model, optimizer, loss_fn, example_inputs, example_target, preprocess=None, warmup=3):
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        self.preprocess = preprocess if preprocess is not None else (lambda *a: a[0])
        self.static_inputs = [t.clone() for t in example_inputs]
        # (a tuple of tensors is passed to loss_fn as a tuple: composite losses with several targets)
        self.static_target = (tuple(t.clone() for t in example_target) if isinstance(example_target, (tuple, list))
                              else example_target.clone())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.static_loss = self._eager_step()

    def _eager_step(self):
        self.optimizer.zero_grad(set_to_none=True)
        pred = self.model(self.preprocess(*self.static_inputs))
        loss = self.loss_fn(pred, self.static_target)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def __call__(self, inputs, target=None):
        for dst, src in zip(self.static_inputs, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        if target is not None:
            pairs = (zip(self.static_target, target) if isinstance(self.static_target, tuple)
                     else [(self.static_target, target)])
            for dst, src in pairs:
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_loss
