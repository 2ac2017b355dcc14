"""Fused temporal graph-convolution layer shared by TGAT and the TGN/DyRep embedding module.

One layer = reference steps ``models/TGAT.py:116-134`` / ``models/MemoryModel.py:644-662``:
time-encode deltas, gather edge rows, MultiHeadAttention, MergeLayer.  Here:
  GEMM  q~   = conv W_qk[:, :F]^T + (W_qk[:, F:] cos(b))          (n, H*Dk)
  K5    s    = dyg_temporal_attend (gather + time enc + softmax)   (n, H*Dk)   <- HBM-bound part
  GEMM  o    = s W_vr^T + residual_fc.bias                          (n, Dq)
  LN    y    = LayerNorm(o + [conv | cos(b)])
  GEMM  out  = fc2(relu(fc1([y | root_feat])))                      (n, F)
"""
from __future__ import annotations

import torch

from .. import ops


_const_cache = {}


def _cached(tag, params, build):
    """Per-weight-version constants of the eval path (they cost one tiny launch each, in steps that are launch-bound)."""
    key = (tag, ops.WEIGHTS_EPOCH) + tuple((p.data_ptr(), p._version) for p in params)
    ent = _const_cache.get(key)
    if ent is None:
        if len(_const_cache) > 64:
            _const_cache.clear()
        ent = _const_cache[key] = (build(), params)     # params kept alive so that data_ptr stays unique
    return ent[0]


def zero_time_features(time_encoder, device):
    """cos(b): the time encoding of a zero interval, a constant of the weights (``models/TGAT.py:82``)."""
    def build():
        w, b = time_encoder.wb()
        return ops.time_encode(torch.zeros(1, dtype=torch.float32, device=device), w, b).reshape(-1)
    return _cached('t0', (time_encoder.w.weight, time_encoder.w.bias), build)


def query_constant(attn, time_encoder, t0):
    """W_qk[:, F:] @ cos(b): the part of the folded query that comes from the root's (zero-interval) time encoding."""
    wqk, _ = attn.folded()
    F_, T_ = attn.node_feat_dim, attn.time_feat_dim
    return _cached('cq', (attn.query_projection.weight, attn.key_projection.weight, time_encoder.w.weight, time_encoder.w.bias),
                   lambda: ops.linear([ops.seg_rows(t0.reshape(1, T_))], 1, wqk[:, F_:], ldw=wqk.stride(0)).reshape(-1))


def temporal_conv(attn, merge, time_encoder, t0, conv, root_feat, node_tab, node_tab2, nbr_ids, nbr_dense,
                  edge_tab, nbr_eids, tq, nbr_t, k, zero_row0=0):
    """conv, root_feat: (n, F) dense; neighbours either lazily gathered from node_tab (+node_tab2) by nbr_ids
    (layer 1) or dense (n*k, F) rows ``nbr_dense`` (deeper layers).  tq float64 (n,), nbr_t float32 (n,k)."""
    n = conv.shape[0]
    F_, E_, T_ = attn.node_feat_dim, attn.edge_feat_dim, attn.time_feat_dim
    H = attn.num_heads
    wqk, wvr = attn.folded()
    w, b = time_encoder.wb()
    cq = query_constant(attn, time_encoder, t0)
    qk = ops.linear([ops.seg_rows(conv)], n, wqk[:, :F_], bias=cq, ldw=wqk.stride(0))
    flat_ids = nbr_ids.reshape(-1)
    if nbr_dense is None:
        s, _ = ops.temporal_attend(qk, n, k, H, node_tab, flat_ids, F_, edge_tab, nbr_eids.reshape(-1), E_, T_, flat_ids,
                                   node_tab2=node_tab2, t_query=tq, t_nbr=nbr_t.reshape(-1), w=w, b=b, zero_row0=zero_row0)
    else:
        s, _ = ops.temporal_attend(qk, n, k, H, nbr_dense, None, F_, edge_tab, nbr_eids.reshape(-1), E_, T_, flat_ids,
                                   t_query=tq, t_nbr=nbr_t.reshape(-1), w=w, b=b, zero_row0=zero_row0 & 2)
    o = ops.linear([ops.seg_rows(s)], n, wvr, attn.residual_fc.bias.detach())
    y = ops.layernorm(o, attn.layer_norm.weight.detach(), attn.layer_norm.bias.detach(), r1=conv, F1=F_, rconst=t0,
                      eps=attn.layer_norm.eps)
    return ops.mlp2([y, root_feat], merge.fc1.weight.detach(), merge.fc1.bias.detach(), merge.fc2.weight.detach(), merge.fc2.bias.detach())


def temporal_conv_train(attn, merge, time_encoder, conv, root_feat, node_tab, nbr_ids, nbr_dense, edge_tab, nbr_eids, tq,
                        nbr_t, k, zero_row0=0):
    """Training-mode ``temporal_conv``: the same folded formulation with autograd (``dyglib_b200/autograd.py``).  The folded
    weights are rebuilt from the parameters every step (four (136 x 272/444) products per head), so their gradients reach
    ``query_projection`` / ``key_projection`` / ``value_projection`` / ``residual_fc``; dropout sits where the reference has
    it (attention probabilities and the ``residual_fc`` output, ``models/modules.py:187,196``)."""
    from .. import autograd as ag
    n = conv.shape[0]
    F_, E_, T_ = attn.node_feat_dim, attn.edge_feat_dim, attn.time_feat_dim
    H, hd = attn.num_heads, attn.head_dim
    wq, wk, wv, r = attn.query_projection.weight, attn.key_projection.weight, attn.value_projection.weight, attn.residual_fc.weight
    wqk = torch.cat([attn.scaling_factor * (wk[h * hd:(h + 1) * hd].t() @ wq[h * hd:(h + 1) * hd]) for h in range(H)], dim=0)
    wvr = torch.cat([r[:, h * hd:(h + 1) * hd] @ wv[h * hd:(h + 1) * hd] for h in range(H)], dim=1)
    t0 = torch.cos(time_encoder.w.bias)                           # time encoding of a zero interval (models/TGAT.py:82)
    query_in = torch.cat([conv, t0.unsqueeze(0).expand(n, T_)], dim=1)
    qk = ag.linear(query_in, wqk)
    flat_ids = nbr_ids.reshape(-1)
    p = attn.dropout.p if attn.training else 0.0
    s = ag.temporal_attend(qk, nbr_dense, time_encoder.w.weight, time_encoder.w.bias, n=n, k=k, H=H, node_tab=node_tab,
                           node_idx=flat_ids, F=F_, edge_tab=edge_tab, edge_idx=nbr_eids.reshape(-1), E=E_, T=T_, mask_ids=flat_ids,
                           t_query=tq, t_nbr=nbr_t.reshape(-1).contiguous(), zero_row0=zero_row0, dropout=p)
    o = ag.linear(s, wvr, attn.residual_fc.bias)
    o = torch.nn.functional.dropout(o, p, attn.training)
    y = ag.layer_norm(o + query_in, attn.layer_norm.weight, attn.layer_norm.bias, attn.layer_norm.eps)
    h = ag.linear([y, root_feat], merge.fc1.weight, merge.fc1.bias, act=ops.ACT_RELU)
    return ag.linear(h, merge.fc2.weight, merge.fc2.bias)
