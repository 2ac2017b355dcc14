"""Drop-ins for the reference's shared modules (``models/modules.py``): ``TimeEncoder``, ``MergeLayer``,
``MultiHeadAttention``.  Same constructor arguments, parameter names and shapes (``state_dict``-compatible);
``forward`` runs the sm_100a kernels of libdygb200.so.  ``MergeLayer`` differentiates through them in training mode
(``dyglib_b200/autograd.py``); the models train through ``_temporal.temporal_conv_train``, so the stand-alone
``MultiHeadAttention.forward`` is eval only and raises in training mode instead of silently skipping dropout.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import ops


def _eval_only(module):
    if module.training:
        raise NotImplementedError(f'{type(module).__name__}: the stand-alone forward of this module is eval only (the models train '
                                  f'through temporal_conv_train); call .eval()')


def _f32(x):
    return x.detach().to(torch.float32).contiguous()


class TimeEncoder(nn.Module):

    def __init__(self, time_dim: int, parameter_requires_grad: bool = True):
        """``TimeEncoder.__init__`` (``models/modules.py:9-25``): w = 1 / 10^linspace(0, 9, time_dim), b = 0."""
        super().__init__()
        self.time_dim = time_dim
        self.w = nn.Linear(1, time_dim)
        self.w.weight = nn.Parameter((torch.from_numpy(1 / 10 ** np.linspace(0, 9, time_dim, dtype=np.float32))).reshape(time_dim, -1))
        self.w.bias = nn.Parameter(torch.zeros(time_dim))
        if not parameter_requires_grad:
            self.w.weight.requires_grad = False
            self.w.bias.requires_grad = False

    def wb(self):
        """(w, b) as flat float32 device vectors for the fused kernels."""
        return self.w.weight.detach().reshape(-1), self.w.bias.detach()

    def forward(self, timestamps: torch.Tensor):
        """``TimeEncoder.forward`` (``models/modules.py:27-39``): (batch, seq) -> (batch, seq, time_dim),
        out = cos(fma(t, w, b)) in fp32."""
        shape = tuple(timestamps.shape)
        dt = _f32(timestamps).reshape(-1)
        w, b = self.wb()
        return ops.time_encode(dt, w, b).reshape(*shape, self.time_dim)


class MergeLayer(nn.Module):

    def __init__(self, input_dim1: int, input_dim2: int, hidden_dim: int, output_dim: int):
        """``MergeLayer`` (``models/modules.py:42-68``)."""
        super().__init__()
        self.fc1 = nn.Linear(input_dim1 + input_dim2, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, output_dim)
        self.act = nn.ReLU()

    def forward(self, input_1: torch.Tensor, input_2: torch.Tensor):
        """fc2(relu(fc1([input_1 | input_2]))) -- the concatenation is never materialised."""
        if self.training and torch.is_grad_enabled():
            from .. import autograd as ag
            h = ag.linear([input_1, input_2], self.fc1.weight, self.fc1.bias, act=ops.ACT_RELU)
            return ag.linear(h, self.fc2.weight, self.fc2.bias)
        return ops.mlp2([_f32(input_1), _f32(input_2)], self.fc1.weight.detach(), self.fc1.bias.detach(), self.fc2.weight.detach(),
                        self.fc2.bias.detach())


class MultiHeadAttention(nn.Module):

    def __init__(self, node_feat_dim: int, edge_feat_dim: int, time_feat_dim: int, num_heads: int = 2, dropout: float = 0.1):
        """``MultiHeadAttention.__init__`` (``models/modules.py:101-135``)."""
        super().__init__()
        self.node_feat_dim = node_feat_dim
        self.edge_feat_dim = edge_feat_dim
        self.time_feat_dim = time_feat_dim
        self.num_heads = num_heads
        self.query_dim = node_feat_dim + time_feat_dim
        self.key_dim = node_feat_dim + edge_feat_dim + time_feat_dim
        assert self.query_dim % num_heads == 0, "The sum of node_feat_dim and time_feat_dim should be divided by num_heads!"
        self.head_dim = self.query_dim // num_heads
        self.query_projection = nn.Linear(self.query_dim, num_heads * self.head_dim, bias=False)
        self.key_projection = nn.Linear(self.key_dim, num_heads * self.head_dim, bias=False)
        self.value_projection = nn.Linear(self.key_dim, num_heads * self.head_dim, bias=False)
        self.scaling_factor = self.head_dim ** -0.5
        self.layer_norm = nn.LayerNorm(self.query_dim)
        self.residual_fc = nn.Linear(num_heads * self.head_dim, self.query_dim)
        self.dropout = nn.Dropout(dropout)
        self._fold_key = None
        self._fold = None

    def folded(self):
        """Eval-mode algebra (exact, rounding differs): score_h = (scale * W_k,h^T W_q,h q_in) . x and
        residual_fc(concat_h(W_v,h sum_j a_j x_j)) = sum_h (R[:, h] W_v,h) (sum_j a_hj x_j) + bias.
        Returns W_qk (H*Dk, Dq) and W_vr (Dq, H*Dk), folded in float64 once per weight version."""
        ps = (self.query_projection.weight, self.key_projection.weight, self.value_projection.weight, self.residual_fc.weight)
        key = (ops.WEIGHTS_EPOCH,) + tuple((p.data_ptr(), p._version) for p in ps)
        if key != self._fold_key:
            H, hd = self.num_heads, self.head_dim
            wq, wk, wv, r = (p.detach().double() for p in ps)
            qk, vr = [], []
            for h in range(H):
                sl = slice(h * hd, (h + 1) * hd)
                qk.append(self.scaling_factor * (wk[sl].t() @ wq[sl]))   # (Dk, Dq)
                vr.append(r[:, sl] @ wv[sl])                             # (Dq, Dk)
            self._fold = (torch.cat(qk, dim=0).float().contiguous(), torch.cat(vr, dim=1).float().contiguous())
            self._fold_key = key
        return self._fold

    def forward(self, node_features: torch.Tensor, node_time_features: torch.Tensor, neighbor_node_features: torch.Tensor,
                neighbor_node_time_features: torch.Tensor, neighbor_node_edge_features: torch.Tensor, neighbor_masks):
        """``MultiHeadAttention.forward`` (``models/modules.py:137-206``); returns (output (n, Dq), scores (n, H, k))."""
        _eval_only(self)
        n, k = neighbor_node_features.shape[0], neighbor_node_features.shape[1]
        F_, E_, T_ = self.node_feat_dim, self.edge_feat_dim, self.time_feat_dim
        dev = node_features.device
        x = _f32(node_features)
        tq = _f32(node_time_features).reshape(n, T_)
        nf = _f32(neighbor_node_features).reshape(n * k, F_)
        ef = _f32(neighbor_node_edge_features).reshape(n * k, E_)
        tf = _f32(neighbor_node_time_features).reshape(n * k, T_)
        if isinstance(neighbor_masks, np.ndarray):
            mask = torch.from_numpy(np.ascontiguousarray(neighbor_masks).astype(np.int64)).to(dev)
        else:
            mask = neighbor_masks.to(device=dev, dtype=torch.int64).contiguous()
        wqk, wvr = self.folded()
        qk = ops.linear([ops.seg_rows(x), ops.seg_rows(tq)], n, wqk)
        s, scores = ops.temporal_attend(qk, n, k, self.num_heads, nf, None, F_, ef, None, E_, T_, mask, time_feat=tf,
                                        want_scores=True)
        o = ops.linear([ops.seg_rows(s)], n, wvr, self.residual_fc.bias.detach())
        res = torch.cat([x, tq], dim=1)
        out = ops.layernorm(o, self.layer_norm.weight.detach(), self.layer_norm.bias.detach(), r1=res, F1=self.query_dim,
                            eps=self.layer_norm.eps)
        return out, scores


class TransformerEncoder(nn.Module):

    def __init__(self, attention_dim: int, num_heads: int, dropout: float = 0.1):
        """``TransformerEncoder`` (``models/modules.py:209-266``, used by TCL): post-norm block with an ``nn.MultiheadAttention``
        parameter container (``state_dict``-compatible), separate query / key inputs and a key padding mask."""
        super().__init__()
        self.multi_head_attention = nn.MultiheadAttention(embed_dim=attention_dim, num_heads=num_heads, dropout=dropout)
        self.dropout = nn.Dropout(dropout)
        self.linear_layers = nn.ModuleList([nn.Linear(attention_dim, 4 * attention_dim), nn.Linear(4 * attention_dim, attention_dim)])
        self.norm_layers = nn.ModuleList([nn.LayerNorm(attention_dim), nn.LayerNorm(attention_dim)])
        self.num_heads = num_heads

    def forward(self, inputs_query: torch.Tensor, inputs_key: torch.Tensor = None, inputs_value: torch.Tensor = None, neighbor_masks=None):
        """``TransformerEncoder.forward`` (``models/modules.py:233-266``).  The five dense layers run forward on the sm_100a GEMMs
        (``autograd.linear``, which also gives the training path); the (B, H, Lq, Lk) masked softmax over ~21 positions,
        LayerNorm, ReLU and dropout are library elementwise ops.  ``inputs_value`` must be ``inputs_key`` (as in every caller)."""
        from .. import autograd as ag
        import torch.nn.functional as F
        if inputs_key is None or inputs_value is None:
            assert inputs_key is None and inputs_value is None
            inputs_key = inputs_value = inputs_query
        assert inputs_value is inputs_key
        B, Lq, D = inputs_query.shape
        Lk = inputs_key.shape[1]
        H = self.num_heads
        hd = D // H
        mha = self.multi_head_attention
        p = self.dropout.p
        W, bias = mha.in_proj_weight, mha.in_proj_bias
        q = ag.linear(inputs_query.reshape(B * Lq, D), W[:D], bias[:D]).reshape(B, Lq, H, hd).transpose(1, 2)
        kv = ag.linear(inputs_key.reshape(B * Lk, D), W[D:], bias[D:]).reshape(B, Lk, 2, H, hd)
        k, v = kv[:, :, 0].transpose(1, 2), kv[:, :, 1].transpose(1, 2)
        att = (q @ k.transpose(-1, -2)) * (hd ** -0.5)
        if neighbor_masks is not None:
            if isinstance(neighbor_masks, np.ndarray):
                neighbor_masks = torch.from_numpy(neighbor_masks)
            pad = (neighbor_masks.to(inputs_query.device) == 0).reshape(B, 1, 1, Lk)
            att = att.masked_fill(pad, float('-inf'))
        att = F.dropout(torch.softmax(att, dim=-1), mha.dropout, self.training)
        a = (att @ v).transpose(1, 2).reshape(B * Lq, D)
        hidden = ag.linear(a, mha.out_proj.weight, mha.out_proj.bias).reshape(B, Lq, D)
        n0, n1 = self.norm_layers
        out = F.layer_norm(inputs_query + F.dropout(hidden, p, self.training), (D,), n0.weight, n0.bias, n0.eps)
        h = F.dropout(torch.relu(ag.linear(out.reshape(B * Lq, D), self.linear_layers[0].weight, self.linear_layers[0].bias)), p, self.training)
        h = ag.linear(h, self.linear_layers[1].weight, self.linear_layers[1].bias).reshape(B, Lq, D)
        return F.layer_norm(out + F.dropout(h, p, self.training), (D,), n1.weight, n1.bias, n1.eps)
