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
        # the captured inputs are views of ONE device slab, so that host inputs arrive with a single H2D copy (see __call__)
        offs, total = [], 0
        for x in example_inputs:
            offs.append(total)
            total += (x.numel() * x.element_size() + 15) // 16 * 16
        dev = example_inputs[0].device
        self._slab = torch.empty(max(total, 16), dtype=torch.uint8, device=dev)
        self._offs, self._nbytes = offs, [x.numel() * x.element_size() for x in example_inputs]
        self.static_in = [self._slab[o:o + n].view(x.dtype).view(x.shape) for o, n, x in zip(offs, self._nbytes, example_inputs)]
        for dst, src in zip(self.static_in, example_inputs):
            dst.copy_(src)
        self._host_ring, self._host_slot = [], 0
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
        elif all((not x.is_cuda) and x.is_pinned() and x.is_contiguous() and x.dtype == d.dtype and x.shape == d.shape
                 for x, d in zip(inputs, self.static_in)):
            # pinned host inputs: packed into a pinned slab (host memcpy of a few KB) and moved with ONE H2D copy; a ring of slabs,
            # each guarded by an event, keeps a slab untouched until its copy has run
            if not self._host_ring:
                self._host_ring = [(torch.empty(self._slab.numel(), dtype=torch.uint8).pin_memory(), torch.cuda.Event()) for _ in range(8)]
            slab, ev = self._host_ring[self._host_slot]
            self._host_slot = (self._host_slot + 1) % len(self._host_ring)
            ev.synchronize()
            for o, n, x in zip(self._offs, self._nbytes, inputs):
                slab[o:o + n].copy_(x.view(-1).view(torch.uint8))
            self._slab.copy_(slab, non_blocking=True)
            ev.record()
        else:
            for dst, src in zip(self.static_in, inputs):
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        if self.grad:
            ops.bump_weights_epoch()   # the replay moved the parameters without bumping their version counters
        ops.launch_count += self.launches
        return self.static_out
