"""Data-parallel training plumbing (the only collective on the path is the gradient all-reduce).

The reference has no distributed code; its training loop (``train_link_prediction.py:159-263``) is one process.  Here
every rank runs the same loop on its own reference batches and the gradients are averaged once per optimizer step.
All gradients live in ONE flat fp32 buffer (0.96-1.46 M parameters = 4-6 MB for the models of the path): ``.grad`` of every
parameter is a view into it, autograd accumulates in place, and the step's collective is a single ``all_reduce`` whose
cost is launch latency, not bandwidth, on NVLink / NVSwitch.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradBucket:

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        """Instead of ``optimizer.zero_grad()`` (which would drop the views when ``set_to_none`` is on)."""
        self.flat.zero_()

    def check_views(self):
        base = self.flat.untyped_storage().data_ptr()
        for p in self.params:
            if p.grad is None or p.grad.untyped_storage().data_ptr() != base:
                raise RuntimeError('a .grad was replaced: use GradBucket.zero(), not optimizer.zero_grad(set_to_none=True)')

    def allreduce(self):
        """Average the gradients over the ranks (no-op in a single process)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())
