"""Training path: ``torch.autograd.Function`` wrappers around the sm_100a kernels.

The reference trains by differentiating its eager PyTorch modules (``train_link_prediction.py:255-257``).  Here the forward
passes stay on the fused kernels; the backward of the neighbour gather / temporal attention is its own kernel
(``dyg_temporal_attend_bwd``), the backward of a dense layer is ``dX = dY W`` on the tcgen05 GEMM plus ``dW = dY^T X`` / ``db`` on
``dyg_gemm_dw``; DyGFormer's transformer block differentiates through ``dyg_layernorm_bwd``, ``dyg_gelu_bwd`` and
``dyg_seq_attention_train_bwd`` (csrc/train.cu).  Dropout masks, ReLU masks and residual adds are elementwise torch ops.
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
        gy = gy.contiguous().float()
        if ctx.act not in (ops.ACT_RELU, ops.ACT_NONE):
            raise NotImplementedError('backward of this activation')
        w = weight.detach().float()
        need_x = any(ctx.needs_input_grad[3 + i] for i in range(len(xs)))
        need_w = ctx.needs_input_grad[0]
        want_b = ctx.has_bias and ctx.needs_input_grad[1]
        Kt = sum(ctx.widths)
        if 2.0 * gy.shape[0] * gy.shape[1] * Kt < ops.LINEAR_BWD_FUSED_FLOPS and w.stride(1) == 1:
            # small layer: dX, dW, db and the ReLU mask in one launch per input segment
            gx, gw, gb = ops.linear_bwd(gy, y if ctx.act == ops.ACT_RELU else None, xs, ctx.widths, w, need_x, need_w, want_b)
        else:
            gx = gw = gb = None
            relu = ctx.act == ops.ACT_RELU
            if need_x:
                # dX = (dY * mask) W: one BF16x3 mma.sync launch; the tcgen05 GEMM on planes of dY and W^T for very large layers
                gx = ops.gemm_dx(gy, w, y if relu else None)
            if relu and (need_w or want_b):
                gy = gy * (y > 0)
            if need_w or want_b:
                # dW = dY^T X (and db = column sums of dY) split over the rows (dyg_gemm_dw); one call per concatenated segment
                buf = torch.zeros(w.shape[0] * (Kt + 1), dtype=torch.float32, device=gy.device)      # dW | db: one memset
                gw = buf[:w.shape[0] * Kt].view(w.shape[0], Kt)
                gb = buf[w.shape[0] * Kt:] if want_b else None
                off = 0
                for i, x in enumerate(xs):
                    ops.gemm_dw(gy, x, dw=gw[:, off:off + ctx.widths[i]], db=gb if i == 0 else None)
                    off += ctx.widths[i]
                if not need_w:
                    gw = None
        if gx is None:
            return (gw, gb, None) + (None,) * len(xs)
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


class _LayerNorm(torch.autograd.Function):
    """``nn.LayerNorm`` over the last dimension of a (M, D) matrix: forward ``dyg_layernorm``, backward ``dyg_layernorm_bwd``."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        xc = x.detach().float().contiguous()
        g, b = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        ctx.save_for_backward(xc, g)
        ctx.eps = eps
        return ops.layernorm(xc, g, b, eps=eps)

    @staticmethod
    def backward(ctx, gy):
        xc, g = ctx.saved_tensors
        dx, dg, db = ops.layernorm_bwd(xc, g, ctx.eps, gy.float().contiguous())
        return dx, dg, db, None


def layer_norm(x, gamma, beta, eps=1e-5):
    """Differentiable LayerNorm of a (M, D) matrix on the library's own kernels."""
    return _LayerNorm.apply(x, gamma, beta, eps)


class _Gelu(torch.autograd.Function):
    """``dropout(F.gelu(v))`` (exact erf, ``models/DyGFormer.py:458``): the dropout multipliers ride inside the kernels."""

    @staticmethod
    def forward(ctx, v, p):
        vc = v.detach().float().contiguous()
        mask = None
        if p > 0.0:
            mask = (torch.rand(vc.shape, device=vc.device) >= p).float() / (1.0 - p)
        ctx.save_for_backward(vc, mask)
        return ops.gelu_fwd(vc, mask, want='f32')

    @staticmethod
    def backward(ctx, gh):
        vc, mask = ctx.saved_tensors
        return ops.gelu_bwd(vc, mask, gh.float().contiguous()), None


def gelu(v, dropout=0.0):
    """Differentiable ``dropout(gelu(v))`` of a (M, N) matrix."""
    return _Gelu.apply(v, dropout)


class _SeqAttention(torch.autograd.Function):
    """The attention core of ``nn.MultiheadAttention`` (``models/DyGFormer.py:454``) on packed (B*S, 3*H*hd) projections, with
    attention dropout: forward ``dyg_seq_attention_train_fwd`` (keeps the softmax), backward ``dyg_seq_attention_train_bwd``."""

    @staticmethod
    def forward(ctx, qkv, B, S, H, hd, p):
        q = qkv.detach().float().contiguous()
        mask = None
        if p > 0.0:
            mask = (torch.rand((B, H, S, S), device=q.device) >= p).float() / (1.0 - p)
        out, probs = ops.seq_attention_train_fwd(q, B, S, H, hd, mask)
        ctx.save_for_backward(q, mask, probs)
        ctx.dims = (B, S, H, hd)
        return out

    @staticmethod
    def backward(ctx, go):
        q, mask, probs = ctx.saved_tensors
        B, S, H, hd = ctx.dims
        return ops.seq_attention_train_bwd(q, B, S, H, hd, mask, probs, go.float().contiguous()), None, None, None, None, None


def seq_attention(qkv, B, S, H, hd, dropout=0.0):
    """Differentiable softmax(q k^T / sqrt(hd)) v per (sequence, head); ``qkv``: (B*S, 3*H*hd) packed [q | k | v]."""
    return _SeqAttention.apply(qkv, B, S, H, hd, dropout)
