"""Training path: ``torch.autograd.Function`` wrappers around the sm_100a kernels.

The reference trains by differentiating its eager PyTorch modules (``train_link_prediction.py:255-257``).  Here the forward
passes stay on the fused kernels; the backward of the neighbour gather / temporal attention is its own kernel
(``dyg_temporal_attend_bwd``), the backward of the dense layers are plain fp32 library GEMMs (``dX = dY W``,
``dW = dY^T X``), and LayerNorm / dropout / ReLU masks are elementwise torch ops on (n, 272)-sized tensors.
There is still no CPU path: every function below needs CUDA tensors.
"""
from __future__ import annotations

import torch

from . import ops


class _Linear(torch.autograd.Function):
    """act(cat(xs, dim=1) @ weight.T + bias): forward on ``ops.linear`` (the concatenation is never materialised)."""

    @staticmethod
    def forward(ctx, weight, bias, act, *xs):
        xs = [x.detach().float().contiguous() for x in xs]
        M = xs[0].shape[0]
        w = weight.detach()
        y = ops.linear([ops.seg_rows(x) for x in xs], M, w, bias.detach() if bias is not None else None, act=act,
                       ldw=w.stride(0))
        ctx.act = act
        ctx.has_bias = bias is not None
        ctx.widths = [x.shape[1] for x in xs]
        ctx.save_for_backward(weight, y if act == ops.ACT_RELU else None, *xs)
        return y

    @staticmethod
    def backward(ctx, gy):
        weight, y, *xs = ctx.saved_tensors
        gy = gy.contiguous()
        if ctx.act == ops.ACT_RELU:
            gy = gy * (y > 0)
        elif ctx.act != ops.ACT_NONE:
            raise NotImplementedError('backward of this activation')
        w = weight.detach()
        gx = gy @ w                                               # (M, K)
        x = xs[0] if len(xs) == 1 else torch.cat(xs, dim=1)
        gw = gy.t() @ x if ctx.needs_input_grad[0] else None
        gb = gy.sum(dim=0) if (ctx.has_bias and ctx.needs_input_grad[1]) else None
        gxs = torch.split(gx, ctx.widths, dim=1) if len(xs) > 1 else (gx,)
        gxs = tuple(g if ctx.needs_input_grad[3 + i] else None for i, g in enumerate(gxs))
        return (gw, gb, None) + gxs


def linear(xs, weight, bias=None, act=ops.ACT_NONE):
    """Differentiable ``act(cat(xs) @ weight.T + bias)``; ``xs``: one (M, K_i) tensor or a list of them."""
    if isinstance(xs, torch.Tensor):
        xs = [xs]
    return _Linear.apply(weight, bias, act, *xs)


class _TemporalAttend(torch.autograd.Function):
    """``ops.temporal_attend`` with the neighbour time encoding computed in the kernel.  Differentiable inputs: the folded
    queries ``qk``, dense neighbour rows ``nbr_dense`` (deeper layers; ``None`` when rows come from a constant table) and the
    time encoder's ``w`` / ``b``."""

    @staticmethod
    def forward(ctx, qk, nbr_dense, w, b, cfg):
        n, k, H = cfg['n'], cfg['k'], cfg['H']
        qk_c = qk.detach().contiguous()
        wv, bv = w.detach().reshape(-1).contiguous(), b.detach().contiguous()
        node_tab = nbr_dense.detach().contiguous() if nbr_dense is not None else cfg['node_tab']
        node_idx = None if nbr_dense is not None else cfg['node_idx']
        p = cfg['dropout']
        prob_scale = None
        if p > 0.0:
            keep = torch.rand((n, H, k), device=qk.device) >= p
            prob_scale = keep.float() / (1.0 - p)
        s, probs = ops.temporal_attend(qk_c, n, k, H, node_tab, node_idx, cfg['F'], cfg['edge_tab'], cfg['edge_idx'], cfg['E'],
                                       cfg['T'], cfg['mask_ids'], t_query=cfg['t_query'], t_nbr=cfg['t_nbr'], w=wv, b=bv,
                                       want_scores=True, zero_row0=cfg['zero_row0'] if nbr_dense is None else cfg['zero_row0'] & 2,
                                       prob_scale=prob_scale)
        ctx.cfg = cfg
        ctx.dense = nbr_dense is not None
        ctx.save_for_backward(qk_c, node_tab if ctx.dense else None, wv, bv, probs, prob_scale, s)
        return s

    @staticmethod
    def backward(ctx, gs):
        qk, nbr_dense, wv, bv, probs, prob_scale, s = ctx.saved_tensors
        cfg = ctx.cfg
        node_tab = nbr_dense if ctx.dense else cfg['node_tab']
        node_idx = None if ctx.dense else cfg['node_idx']
        want_nbr = ctx.dense and ctx.needs_input_grad[1]
        gqk, gnbr, gw, gb = ops.temporal_attend_bwd(
            qk, cfg['n'], cfg['k'], cfg['H'], node_tab, node_idx, cfg['F'], cfg['edge_tab'], cfg['edge_idx'], cfg['E'], cfg['T'],
            cfg['mask_ids'], cfg['t_query'], cfg['t_nbr'], wv, bv, probs, prob_scale, s, gs.contiguous(), want_nbr=want_nbr)
        return (gqk, gnbr if want_nbr else None, gw.reshape(-1, 1) if ctx.needs_input_grad[2] else None,
                gb if ctx.needs_input_grad[3] else None, None)


def temporal_attend(qk, nbr_dense, w, b, **cfg):
    """``w``: the time encoder's (T, 1) weight parameter, ``b``: its (T,) bias."""
    return _TemporalAttend.apply(qk, nbr_dense, w, b, cfg)


class _TimeEncode(torch.autograd.Function):
    """``TimeEncoder.forward`` (``models/modules.py:27-39``) with gradients for ``w`` / ``b``: forward on ``dyg_time_encode`` (the fp32
    FMA argument of the reference), backward on ``dyg_time_encode_bwd`` (sin of the same argument)."""

    @staticmethod
    def forward(ctx, dt, w, b):
        d = dt.detach().float().reshape(-1).contiguous()
        wv, bv = w.detach().reshape(-1).contiguous(), b.detach().contiguous()
        ctx.save_for_backward(d, wv, bv)
        ctx.shape = tuple(dt.shape)
        return ops.time_encode(d, wv, bv).reshape(*dt.shape, wv.numel())

    @staticmethod
    def backward(ctx, g):
        d, wv, bv = ctx.saved_tensors
        g2 = g.reshape(-1, wv.numel()).contiguous()
        gw, gb = ops.time_encode_bwd(d, wv, bv, g2)
        return None, gw.reshape(-1, 1), gb


def time_encode(dt, w, b):
    """Differentiable time encoding of ``dt`` (any shape) -> (*dt.shape, T); ``w``: (T, 1) weight, ``b``: (T,) bias."""
    return _TimeEncode.apply(dt, w, b)
