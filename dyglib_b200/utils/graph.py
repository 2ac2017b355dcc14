"""CUDA-graph replay of a fixed-shape step.

The memory models advance one 200-event batch at a time through ~30 small kernels: at that size the
step is launch-bound, not bandwidth-bound.  Every op of the package is a plain stream-ordered launch through the C ABI
with no host synchronisation, so a whole step (sampling, embedding, memory update, link scores) can be captured once
into a CUDA graph and replayed with new inputs copied into the captured buffers.
"""
from __future__ import annotations

import torch

from .. import ops


class GraphedStep:
    """``fn(*tensors) -> tensor`` captured for one input signature; ``__call__`` copies the inputs into the captured
    buffers (device-to-device, or host-to-device from pinned memory) and replays the graph.

    ``fn`` must be free of host synchronisation and of data-dependent shapes; state it mutates in place (the TGN memory
    bank) is mutated by every replay exactly as by a direct call.  ``warmup`` direct calls are made before capture so
    that lazily built caches exist; ``after_warmup`` (e.g. a memory reset) runs after them, before capture."""

    def __init__(self, fn, example_inputs, warmup: int = 2, after_warmup=None, grad: bool = False):
        """``grad=True`` captures a whole training step (forward, ``backward()``, gradient all-reduce, a ``capturable``
        optimizer step): gradients must then accumulate into preallocated buffers (``utils.dist.GradBucket``)."""
        self.fn = fn
        self.grad = grad
        mode = torch.enable_grad if grad else torch.no_grad
        self.static_in = [x.clone() for x in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), mode():
            for _ in range(warmup):
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if after_warmup is not None:
            after_warmup()
            torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.launch_count
        with mode(), torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)
        self.launches = ops.launch_count - n0     # kernels of the package inside one replay
        # capture records, it does not run: the state is still what after_warmup left

    def __call__(self, *inputs):
        if all(x.is_cuda for x in inputs):
            # device-resident inputs: one multi-tensor copy kernel per dtype instead of one launch per tensor (a 200-event memory
            # step is ~80 us of kernel time; five 3-us copy launches in front of it were visible in the step time)
            by_dtype = {}
            for dst, src in zip(self.static_in, inputs):
                by_dtype.setdefault((dst.dtype, src.dtype), ([], []))
                by_dtype[(dst.dtype, src.dtype)][0].append(dst)
                by_dtype[(dst.dtype, src.dtype)][1].append(src)
            for dsts, srcs in by_dtype.values():
                if len(dsts) > 1:
                    torch._foreach_copy_(dsts, srcs, non_blocking=True)
                else:
                    dsts[0].copy_(srcs[0], non_blocking=True)
        else:
            for dst, src in zip(self.static_in, inputs):
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        if self.grad:
            ops.bump_weights_epoch()   # the replay moved the parameters without bumping their version counters
        ops.launch_count += self.launches
        return self.static_out
