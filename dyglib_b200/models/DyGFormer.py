"""Drop-in for the reference's ``DyGFormer`` and ``NeighborCooccurrenceEncoder`` (``models/DyGFormer.py``).

Fused device path per batch:
  K3  first-hop search + pad       -> (B, Lp) ids / edge ids / times directly in HBM  (no ragged lists)
  K4  co-occurrence counts         -> exact integer counts, shared-memory O(L^2) per pair
  K6  patch projections            -> 4 gather-GEMMs (node rows, edge rows, on-the-fly time encoding,
                                      count-LUT rows) written straight into the (B, S, 4C) token buffer
  K7  transformer blocks           -> LayerNorm, QKV GEMM, small-sequence attention, out-proj + residual, FFN
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
from torch.nn import MultiheadAttention

from .. import ops
from ..utils.utils import NeighborSampler, _as_dev
from .modules import TimeEncoder


import os

# LayerNorm + feed-forward as one kernel (dyg_ln_ffn_bf16x3); DYG_FUSED_FFN=0 selects the three-kernel path
FUSED_FFN = os.environ.get('DYG_FUSED_FFN', '1') != '0'
# attention sub-block: 1 (default) = LayerNorm + [q | k | v'] projection GEMM, then dyg_seq_attention_fold (tcgen05),
# 2 = one kernel (dyg_attn_block; parity-green but measured slower: L2-bound scratch round trip),
# 0 = QKV GEMM + mma.sync attention + out-projection GEMM
FUSED_ATTN = int(os.environ.get('DYG_FUSED_ATTN', '1'))
# LayerNorm inside the projection GEMM (dyg_ln_gemm_bf16x3); DYG_FUSED_LN=0 runs layernorm_split + dyg_gemm_bf16x3
FUSED_LN = os.environ.get('DYG_FUSED_LN', '1') != '0'


class NeighborCooccurrenceEncoder(nn.Module):

    def __init__(self, neighbor_co_occurrence_feat_dim: int, device: str = 'cuda'):
        """``NeighborCooccurrenceEncoder.__init__`` (``models/DyGFormer.py:322-335``)."""
        super().__init__()
        self.neighbor_co_occurrence_feat_dim = neighbor_co_occurrence_feat_dim
        self.device = device
        self.neighbor_co_occurrence_encode_layer = nn.Sequential(
            nn.Linear(in_features=1, out_features=self.neighbor_co_occurrence_feat_dim),
            nn.ReLU(),
            nn.Linear(in_features=self.neighbor_co_occurrence_feat_dim, out_features=self.neighbor_co_occurrence_feat_dim))
        self._lut_key = None
        self._lut = None

    def _ids(self, x):
        dev = self.neighbor_co_occurrence_encode_layer[0].weight.device
        return _as_dev(x, torch.int64, dev)

    def count_nodes_appearances(self, src_padded_nodes_neighbor_ids, dst_padded_nodes_neighbor_ids):
        """``count_nodes_appearances`` (``models/DyGFormer.py:337-393``): float tensors (B, Ls, 2), (B, Ld, 2)."""
        s, d = self._ids(src_padded_nodes_neighbor_ids), self._ids(dst_padded_nodes_neighbor_ids)
        fs, fd, _, _ = ops.cooc_count(s, d, want_float=True)
        return fs, fd

    def lut(self, max_count: int):
        """Rows c = 0..max_count of MLP(c) = W2 relu(W1 c + b1) + b2, padded to a multiple of 4 columns.
        Counts are small integers, so the per-position MLP of the reference (``:409-411``) is a table lookup."""
        l0, l2 = self.neighbor_co_occurrence_encode_layer[0], self.neighbor_co_occurrence_encode_layer[2]
        key = (max_count, ops.WEIGHTS_EPOCH) + tuple((p.data_ptr(), p._version) for p in (l0.weight, l0.bias, l2.weight, l2.bias))
        if key != self._lut_key:
            dev = l0.weight.device
            C = self.neighbor_co_occurrence_feat_dim
            Cp = (C + 3) // 4 * 4
            c = torch.arange(max_count + 1, dtype=torch.float32, device=dev).reshape(-1, 1)
            h = ops.linear([ops.seg_rows(c)], max_count + 1, l0.weight.detach(), l0.bias.detach(), act=ops.ACT_RELU)
            lut = torch.zeros((max_count + 1, Cp), dtype=torch.float32, device=dev)
            ops.linear([ops.seg_rows(h)], max_count + 1, l2.weight.detach(), l2.bias.detach(), out=lut[:, :C])
            self._lut, self._lut_key = lut, key
        return self._lut

    def forward(self, src_padded_nodes_neighbor_ids, dst_padded_nodes_neighbor_ids):
        """``NeighborCooccurrenceEncoder.forward`` (``models/DyGFormer.py:395-415``): two (B, L, C) tensors."""
        s, d = self._ids(src_padded_nodes_neighbor_ids), self._ids(dst_padded_nodes_neighbor_ids)
        _, _, cs, cd = ops.cooc_count(s, d, want_float=False, want_int=True)
        C = self.neighbor_co_occurrence_feat_dim
        lut = self.lut(s.shape[1] + d.shape[1])
        outs = []
        for cnt, ids in ((cs, s), (cd, d)):
            B, L = ids.shape
            a = ops.gather_rows(lut, cnt[0].reshape(-1))
            b = ops.gather_rows(lut, cnt[1].reshape(-1))
            outs.append((a + b)[:, :C].reshape(B, L, C))
        return outs[0], outs[1]


class TransformerEncoder(nn.Module):

    def __init__(self, attention_dim: int, num_heads: int, dropout: float = 0.1):
        """``TransformerEncoder.__init__`` (``models/DyGFormer.py:420-440``); parameter container + fused forward."""
        super().__init__()
        self.multi_head_attention = MultiheadAttention(embed_dim=attention_dim, num_heads=num_heads, dropout=dropout)
        self.dropout = nn.Dropout(dropout)
        self.linear_layers = nn.ModuleList([nn.Linear(attention_dim, 4 * attention_dim), nn.Linear(4 * attention_dim, attention_dim)])
        self.norm_layers = nn.ModuleList([nn.LayerNorm(attention_dim), nn.LayerNorm(attention_dim)])
        self.attention_dim = attention_dim
        self.num_heads = num_heads
        self._fold_key = None
        self._fold = None

    def forward(self, inputs: torch.Tensor):
        """``TransformerEncoder.forward`` (``models/DyGFormer.py:442-461``): pre-norm block, no padding mask."""
        if self.training and torch.is_grad_enabled():
            return self._forward_train(inputs)
        B, S, D = inputs.shape
        x = inputs.detach().to(torch.float32).contiguous().reshape(B * S, D)
        mha = self.multi_head_attention
        n0, n1 = self.norm_layers
        l0, l1 = self.linear_layers
        # every dense contraction runs on tcgen05 from BF16x3 operand planes (ops.gemm); the residual stream stays fp32
        hd = D // self.num_heads
        if FUSED_ATTN == 2 and ops.attn_fold_fusable(S, D, self.num_heads):
            # LayerNorm + [q | k | v'] projection (out-projection folded into v') + attention + residual in ONE kernel: the
            # projected rows never reach HBM
            wcat, bcat, bout = self._folded()
            x1 = ops.attn_block(x, n0.weight.detach(), n0.bias.detach(), n0.eps, wcat, bcat, bout, B, S, self.num_heads, D)
        elif FUSED_ATTN and ops.attn_fold_fusable(S, D, self.num_heads):
            # two launches: LayerNorm + [q | k | v'] projection (out-projection folded into v'), then the tcgen05 attention
            # kernel, which also adds the residual
            wcat, bcat, bout = self._folded()
            if FUSED_LN and ops.ln_gemm_fusable(D):
                pl = ops.ln_gemm(x, n0.weight.detach(), n0.bias.detach(), n0.eps, wcat, bcat, want='split')
            else:
                pl = ops.gemm(ops.layernorm_split(x, n0.weight.detach(), n0.bias.detach(), eps=n0.eps), wcat, bcat, want='split')
            x1 = ops.seq_attention_fold(pl, B, S, self.num_heads, D, x, bout)
        else:
            y = ops.layernorm_split(x, n0.weight.detach(), n0.bias.detach(), eps=n0.eps)
            qkv = ops.gemm(y, mha.in_proj_weight, mha.in_proj_bias.detach())
            if S <= 128 and hd <= 128 and hd % 2 == 0:
                a = ops.seq_attention_tc(qkv, B, S, self.num_heads, hd, want='split')
            else:
                a = ops.split_bf16(ops.seq_attention(qkv, B, S, self.num_heads, hd))
            x1 = ops.gemm(a, mha.out_proj.weight, mha.out_proj.bias.detach(), residual=x)
        if FUSED_FFN and ops.ffn_fusable(D, l0.weight.shape[0]):
            # LayerNorm + FFN in one kernel: the normalised rows and the 4D hidden activation stay on the SM
            out = ops.ln_ffn(x1, n1.weight.detach(), n1.bias.detach(), n1.eps, l0.weight, l0.bias.detach(), l1.weight, l1.bias.detach())
        else:
            y = ops.layernorm_split(x1, n1.weight.detach(), n1.bias.detach(), eps=n1.eps)
            h = ops.gemm(y, l0.weight, l0.bias.detach(), act=ops.ACT_GELU, want='split')
            out = ops.gemm(h, l1.weight, l1.bias.detach(), residual=x1)
        return out.reshape(B, S, D)


    def _folded(self):
        """(Split of W_cat, b_cat, b_out) of ops.attn_fold_weights, rebuilt when a parameter of the attention changes."""
        mha = self.multi_head_attention
        ps = (mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias)
        key = (ops.WEIGHTS_EPOCH,) + tuple((q.data_ptr(), q._version) for q in ps)
        if key != self._fold_key:
            W, b, bo = ops.attn_fold_weights(ps[0], ps[1], ps[2], ps[3], self.num_heads)
            self._fold, self._fold_key = (ops.split_bf16(W), b, bo), key
        return self._fold

    def _forward_train(self, inputs):
        """Training mode: the same block under autograd, on the library's kernels both ways (``dyglib_b200/autograd.py``): dense
        layers forward on the GEMM kernels, backward ``dX`` on the tcgen05 GEMM and ``dW`` / ``db`` on ``dyg_gemm_dw``; LayerNorm,
        GELU and the softmax(QK^T)V core with their own backward kernels (``csrc/train.cu``).  Dropout sits where the reference has it:
        attention probabilities (``nn.MultiheadAttention(dropout=...)``), the attention output, the hidden activation and the FFN
        output; the dropout masks of the two residual branches and the residual adds are elementwise library ops."""
        from .. import autograd as ag
        import torch.nn.functional as F
        B, S, D = inputs.shape
        H = self.num_heads
        hd = D // H
        mha = self.multi_head_attention
        n0, n1 = self.norm_layers
        l0, l1 = self.linear_layers
        p = self.dropout.p
        x = inputs.reshape(B * S, D)
        if S > 64 or hd > 128:
            return self._forward_train_library(inputs)
        y = ag.layer_norm(x, n0.weight, n0.bias, n0.eps)
        qkv = ag.linear(y, mha.in_proj_weight, mha.in_proj_bias)                        # (B*S, 3D) packed [q | k | v]
        a = ag.seq_attention(qkv, B, S, H, hd, mha.dropout)
        x1 = x + F.dropout(ag.linear(a, mha.out_proj.weight, mha.out_proj.bias), p, True)
        h = ag.gelu(ag.linear(ag.layer_norm(x1, n1.weight, n1.bias, n1.eps), l0.weight, l0.bias), p)
        out = x1 + F.dropout(ag.linear(h, l1.weight, l1.bias), p, True)
        return out.reshape(B, S, D)

    def _forward_train_library(self, inputs):
        """Sequences longer than 64 tokens (outside the training kernels' range): LayerNorm / GELU / softmax as library ops."""
        from .. import autograd as ag
        import torch.nn.functional as F
        B, S, D = inputs.shape
        H = self.num_heads
        hd = D // H
        mha = self.multi_head_attention
        n0, n1 = self.norm_layers
        l0, l1 = self.linear_layers
        p = self.dropout.p
        x = inputs.reshape(B * S, D)
        y = F.layer_norm(x, (D,), n0.weight, n0.bias, n0.eps)
        qkv = ag.linear(y, mha.in_proj_weight, mha.in_proj_bias).reshape(B, S, 3, H, hd)
        q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))                       # (B, H, S, hd)
        att = torch.softmax((q @ k.transpose(-1, -2)) * (hd ** -0.5), dim=-1)
        att = F.dropout(att, mha.dropout, True)
        a = (att @ v).transpose(1, 2).reshape(B * S, D)
        x1 = x + F.dropout(ag.linear(a, mha.out_proj.weight, mha.out_proj.bias), p, True)
        h = F.dropout(F.gelu(ag.linear(F.layer_norm(x1, (D,), n1.weight, n1.bias, n1.eps), l0.weight, l0.bias)), p, True)
        out = x1 + F.dropout(ag.linear(h, l1.weight, l1.bias), p, True)
        return out.reshape(B, S, D)


class DyGFormer(nn.Module):

    def __init__(self, node_raw_features: np.ndarray, edge_raw_features: np.ndarray, neighbor_sampler: NeighborSampler,
                 time_feat_dim: int, channel_embedding_dim: int, patch_size: int = 1, num_layers: int = 2, num_heads: int = 2,
                 dropout: float = 0.1, max_input_sequence_length: int = 512, device: str = 'cuda'):
        """Same arguments as ``DyGFormer.__init__`` (``models/DyGFormer.py:13-66``)."""
        super().__init__()
        self.node_raw_features = torch.from_numpy(node_raw_features.astype(np.float32)).to(device).contiguous()
        self.edge_raw_features = torch.from_numpy(edge_raw_features.astype(np.float32)).to(device).contiguous()
        self.neighbor_sampler = neighbor_sampler
        self.node_feat_dim = self.node_raw_features.shape[1]
        self.edge_feat_dim = self.edge_raw_features.shape[1]
        self.time_feat_dim = time_feat_dim
        self.channel_embedding_dim = channel_embedding_dim
        self.patch_size = patch_size
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.dropout = dropout
        self.max_input_sequence_length = max_input_sequence_length
        self.device = device
        self.time_encoder = TimeEncoder(time_dim=time_feat_dim)
        self.neighbor_co_occurrence_feat_dim = self.channel_embedding_dim
        self.neighbor_co_occurrence_encoder = NeighborCooccurrenceEncoder(self.neighbor_co_occurrence_feat_dim, device=self.device)
        self.projection_layer = nn.ModuleDict({
            'node': nn.Linear(self.patch_size * self.node_feat_dim, self.channel_embedding_dim, bias=True),
            'edge': nn.Linear(self.patch_size * self.edge_feat_dim, self.channel_embedding_dim, bias=True),
            'time': nn.Linear(self.patch_size * self.time_feat_dim, self.channel_embedding_dim, bias=True),
            'neighbor_co_occurrence': nn.Linear(self.patch_size * self.neighbor_co_occurrence_feat_dim, self.channel_embedding_dim, bias=True)})
        self.num_channels = 4
        self.transformers = nn.ModuleList([
            TransformerEncoder(self.num_channels * self.channel_embedding_dim, self.num_heads, self.dropout)
            for _ in range(self.num_layers)])
        self.output_layer = nn.Linear(self.num_channels * self.channel_embedding_dim, self.node_feat_dim, bias=True)
        self._cooc_w_key = None
        self._cooc_w = None
        self._planes = None
        self._proj_key = None
        self._proj_ops = None
        self.to(device)

    # ------------------------------------------------------------------ reference-compatible pieces
    def pad_sequences(self, node_ids, node_interact_times, nodes_neighbor_ids_list=None, nodes_edge_ids_list=None,
                      nodes_neighbor_times_list=None, patch_size: int = 1, max_input_sequence_length: int = 256):
        """``pad_sequences`` (``models/DyGFormer.py:196-245``).  The ragged lists of the reference are not needed:
        the device kernel searches and pads in one pass.  Returns numpy arrays like the reference."""
        pn, pe, pt = self._padded_device(node_ids, node_interact_times, patch_size, max_input_sequence_length)
        return pn.cpu().numpy(), pe.cpu().numpy(), pt.cpu().numpy()

    def _padded_device(self, node_ids, node_interact_times, patch_size, L):
        pn, pe, pt, ln, _ = self.neighbor_sampler.get_all_first_hop_neighbors_device(node_ids, node_interact_times, L, patch_size)
        # batch-dependent padded length: (batch max + self) rounded up to the patch size (models/DyGFormer.py:219-226)
        mx = int(ln.max().item()) if ln.numel() else 1
        Lp = (mx + patch_size - 1) // patch_size * patch_size
        if Lp != pn.shape[1]:
            pn, pe, pt = pn[:, :Lp].contiguous(), pe[:, :Lp].contiguous(), pt[:, :Lp].contiguous()
        return pn, pe, pt

    def get_features(self, node_interact_times, padded_nodes_neighbor_ids, padded_nodes_edge_ids, padded_nodes_neighbor_times,
                     time_encoder: TimeEncoder = None):
        """``get_features`` (``models/DyGFormer.py:247-268``): gathered node / edge rows and time encodings
        (zeroed at padded positions), as dense tensors (compat API; the fused path never materialises them)."""
        dev = self.node_raw_features.device
        pn = _as_dev(padded_nodes_neighbor_ids, torch.int64, dev)
        pe = _as_dev(padded_nodes_edge_ids, torch.int64, dev)
        pt = _as_dev(padded_nodes_neighbor_times, torch.float32, dev)
        tq = _as_dev(node_interact_times, torch.float64, dev)
        B, Lp = pn.shape
        nf = ops.gather_rows(self.node_raw_features, pn.reshape(-1)).reshape(B, Lp, -1)
        ef = ops.gather_rows(self.edge_raw_features, pe.reshape(-1)).reshape(B, Lp, -1)
        w, b = (time_encoder or self.time_encoder).wb()
        T = self.time_feat_dim
        eye = torch.eye(T, dtype=torch.float32, device=dev)
        tf = ops.linear([ops.seg_time(pt.reshape(-1), w, b, mask_ids=pn.reshape(-1), t_query=tq, tq_div=Lp)], B * Lp, eye, tc=False)
        return nf, ef, tf.reshape(B, Lp, T)

    def get_patches(self, padded_nodes_neighbor_node_raw_features, padded_nodes_edge_raw_features,
                    padded_nodes_neighbor_time_features, padded_nodes_neighbor_co_occurrence_features=None, patch_size: int = 1):
        """``get_patches`` (``models/DyGFormer.py:270-306``): a reshape, (B, L, F) -> (B, L/P, P*F)."""
        outs = []
        for x in (padded_nodes_neighbor_node_raw_features, padded_nodes_edge_raw_features,
                  padded_nodes_neighbor_time_features, padded_nodes_neighbor_co_occurrence_features):
            assert x.shape[1] % patch_size == 0
            outs.append(x.reshape(x.shape[0], x.shape[1] // patch_size, patch_size * x.shape[2]))
        return tuple(outs)

    def _cooc_weight(self):
        """projection_layer['neighbor_co_occurrence'].weight with every patch slot padded to the LUT row width."""
        wt = self.projection_layer['neighbor_co_occurrence'].weight
        key = (wt.data_ptr(), wt._version, ops.WEIGHTS_EPOCH)
        if key != self._cooc_w_key:
            C, P = self.neighbor_co_occurrence_feat_dim, self.patch_size
            Cp = (C + 3) // 4 * 4
            wp = torch.zeros((wt.shape[0], P, Cp), dtype=torch.float32, device=wt.device)
            wp[:, :, :C] = wt.detach().reshape(wt.shape[0], P, C)
            self._cooc_w, self._cooc_w_key = wp.reshape(wt.shape[0], P * Cp).contiguous(), key
        return self._cooc_w

    # ------------------------------------------------------------------ fused forward
    def compute_src_dst_node_temporal_embeddings(self, src_node_ids: np.ndarray, dst_node_ids: np.ndarray, node_interact_times: np.ndarray,
                                                 batch_size: int = None):
        """``compute_src_dst_node_temporal_embeddings`` (``models/DyGFormer.py:68-194``).

        ``batch_size`` (extension): the rows are treated as consecutive reference batches of that many events, each
        keeping its own padded length (the reference's padding unit); batches that share padded
        lengths run through the kernels together, so results equal calling the model once per batch."""
        dev = self.node_raw_features.device
        P, L = self.patch_size, self.max_input_sequence_length
        tq = _as_dev(node_interact_times, torch.float64, dev)
        src = _as_dev(src_node_ids, torch.int64, dev)
        dst = _as_dev(dst_node_ids, torch.int64, dev)
        B = src.numel()
        if batch_size is None or batch_size >= B:
            s_pn, s_pe, s_pt = self._padded_device(src, tq, P, L)
            d_pn, d_pe, d_pt = self._padded_device(dst, tq, P, L)
            return self._forward_padded(tq, s_pn, s_pe, s_pt, d_pn, d_pe, d_pt)
        samp = self.neighbor_sampler
        s_pn, s_pe, s_pt, _, s_gm = samp.get_all_first_hop_neighbors_device(src, tq, L, P, group_size=batch_size)
        d_pn, d_pe, d_pt, _, d_gm = samp.get_all_first_hop_neighbors_device(dst, tq, L, P, group_size=batch_size)
        gm = torch.stack([s_gm, d_gm]).cpu().numpy()                 # the one host sync of the step
        lp = (gm + P - 1) // P * P
        buckets = {}
        for gi in range(lp.shape[1]):
            buckets.setdefault((int(lp[0, gi]), int(lp[1, gi])), []).append(gi)
        out_s = torch.empty((B, self.node_feat_dim), dtype=torch.float32, device=dev)
        out_d = torch.empty((B, self.node_feat_dim), dtype=torch.float32, device=dev)
        for (ls, ld), groups in buckets.items():
            if len(groups) == lp.shape[1]:
                rows = None
            else:
                rows = torch.cat([torch.arange(gi * batch_size, min((gi + 1) * batch_size, B), device=dev) for gi in groups])

            def pick(x, width):
                x = x if rows is None else x[rows]
                return x if width == x.shape[1] else x[:, :width].contiguous()
            es, ed = self._forward_padded(tq if rows is None else tq[rows], pick(s_pn, ls), pick(s_pe, ls), pick(s_pt, ls),
                                          pick(d_pn, ld), pick(d_pe, ld), pick(d_pt, ld))
            if rows is None:
                return es, ed
            out_s[rows] = es
            out_d[rows] = ed
        return out_s, out_d

    def _forward_padded(self, tq, s_pn, s_pe, s_pt, d_pn, d_pe, d_pt):
        """Everything after padding (``models/DyGFormer.py:102-194``) for rows that share padded lengths."""
        dev = self.node_raw_features.device
        P, C = self.patch_size, self.channel_embedding_dim
        D = self.num_channels * C
        B = tq.numel()
        Ls, Ld = s_pn.shape[1], d_pn.shape[1]
        ns, nd = Ls // P, Ld // P
        S = ns + nd
        _, _, cs, cd = ops.cooc_count(s_pn, d_pn, want_float=False, want_int=True)
        Wmax = (self.max_input_sequence_length + P - 1) // P * P
        if self.training and torch.is_grad_enabled():
            return self._forward_padded_train(tq, s_pn, s_pe, s_pt, cs, d_pn, d_pe, d_pt, cd)
        lut = self.neighbor_co_occurrence_encoder.lut(2 * Wmax)   # one table for every padded length
        w, b = self.time_encoder.wb()
        X = torch.empty((B * S, D), dtype=torch.float32, device=dev)
        pl = self.projection_layer
        if C <= 64 and C % 2 == 0:
            # one fused gather + time-encode + patch-projection kernel over both sides (dyg_patch_project)
            node_pl, edge_pl, zero_rows = self._table_planes()
            lut_pl, packed, bias = self._projection_operands(lut)
            sides = [(s_pn, s_pe, s_pt, cs[0], cs[1], ns, 0), (d_pn, d_pe, d_pt, cd[0], cd[1], nd, ns)]
            ops.patch_project(sides, node_pl, self.node_feat_dim, edge_pl, self.edge_feat_dim, lut_pl,
                              self.neighbor_co_occurrence_feat_dim, tq, w, b, packed, bias, P, C, S, X, zero_rows=zero_rows)
        else:
            self._project_unfused(X, tq, s_pn, s_pe, s_pt, cs, d_pn, d_pe, d_pt, cd, lut, w, b)
        x = X.reshape(B, S, D)
        for tr in self.transformers:
            x = tr(x)
        means = torch.empty((2 * B, D), dtype=torch.float32, device=dev)
        ops.mean_tokens(x, B, S, D, 0, ns, out=means[:B])
        ops.mean_tokens(x, B, S, D, ns, nd, out=means[B:])
        out = ops.gemm(ops.split_bf16(means), self.output_layer.weight, self.output_layer.bias.detach())
        return out[:B], out[B:]

    def _forward_padded_train(self, tq, s_pn, s_pe, s_pt, cs, d_pn, d_pe, d_pt, cd):
        """Training mode of ``_forward_padded`` (``models/DyGFormer.py:102-194`` under autograd).  Sampling, padding and the
        co-occurrence counts come from the same kernels (integers, no gradient); node / edge rows are constant gathers; the
        time encoding is ``autograd.time_encode`` (gradients for the encoder's w / b from ``dyg_time_encode_bwd``); the
        co-occurrence MLP is evaluated once per distinct count (a differentiable table) and indexed; the four channel
        projections, the transformer's dense layers and the output layer run on ``autograd.linear``."""
        from .. import autograd as ag
        import torch.nn.functional as F
        P, C = self.patch_size, self.channel_embedding_dim
        B = tq.numel()
        pl = self.projection_layer
        enc = self.neighbor_co_occurrence_encoder.neighbor_co_occurrence_encode_layer
        max_count = int(max(cs[0].max().item(), cs[1].max().item(), cd[0].max().item(), cd[1].max().item()))
        counts = torch.arange(max_count + 1, dtype=torch.float32, device=tq.device).reshape(-1, 1)
        lut = ag.linear(ag.linear(counts, enc[0].weight, enc[0].bias, act=ops.ACT_RELU), enc[2].weight, enc[2].bias)   # (max_count + 1, C), models/DyGFormer.py:409-411
        toks = []
        for pn, pe, pt, cnt in ((s_pn, s_pe, s_pt, cs), (d_pn, d_pe, d_pt, cd)):
            Lp = pn.shape[1]
            nodef = ops.gather_rows(self.node_raw_features, pn.reshape(-1)).reshape(B, Lp, -1)      # constant tables: no gradient
            edgef = ops.gather_rows(self.edge_raw_features, pe.reshape(-1)).reshape(B, Lp, -1)
            dt = (tq.reshape(B, 1) - pt.double()).float()
            te = ag.time_encode(dt, self.time_encoder.w.weight, self.time_encoder.w.bias) * (pn != 0).unsqueeze(-1)
            co = lut[cnt[0]] + lut[cnt[1]]
            chans = []
            for name, x in (('node', nodef), ('edge', edgef), ('time', te), ('neighbor_co_occurrence', co)):
                patches = x.reshape(B * (Lp // P), P * x.shape[2])
                chans.append(ag.linear(patches, pl[name].weight, pl[name].bias).reshape(B, Lp // P, C))
            toks.append(torch.stack(chans, dim=2).reshape(B, Lp // P, self.num_channels * C))
        ns = toks[0].shape[1]
        x = torch.cat(toks, dim=1)
        for tr in self.transformers:
            x = tr(x)
        means = torch.cat([x[:, :ns].mean(dim=1), x[:, ns:].mean(dim=1)], dim=0)
        out = ag.linear(means, self.output_layer.weight, self.output_layer.bias)
        return out[:B], out[B:]

    def _project_unfused(self, X, tq, s_pn, s_pe, s_pt, cs, d_pn, d_pe, d_pt, cd, lut, w, b):
        """Channel-by-channel gather-GEMMs (dyg_linear_tc) for channel widths the fused kernel does not take."""
        P, C = self.patch_size, self.channel_embedding_dim
        B = tq.numel()
        Ls, Ld = s_pn.shape[1], d_pn.shape[1]
        ns, nd = Ls // P, Ld // P
        S = ns + nd
        pl = self.projection_layer
        for pn, pe, pt, cnt, Lp, ntok, off in ((s_pn, s_pe, s_pt, cs, Ls, ns, 0), (d_pn, d_pe, d_pt, cd, Ld, nd, ns)):
            M = B * ntok
            ids_flat = pn.reshape(-1)
            chans = (
                (ops.seg_rows(self.node_raw_features, self.node_feat_dim, ids_flat, group=P), pl['node'].weight.detach(), pl['node'].bias),
                (ops.seg_rows(self.edge_raw_features, self.edge_feat_dim, pe.reshape(-1), group=P), pl['edge'].weight.detach(), pl['edge'].bias),
                (ops.seg_time(pt.reshape(-1), w, b, mask_ids=ids_flat, group=P, t_query=tq, tq_div=Lp), pl['time'].weight.detach(), pl['time'].bias),
                (ops.seg_rows(lut, lut.shape[1], cnt[0].reshape(-1), group=P, table2=lut, idx2=cnt[1].reshape(-1)),
                 self._cooc_weight(), pl['neighbor_co_occurrence'].bias),
            )
            for ch, (seg, wt, bias) in enumerate(chans):
                ops.linear([seg], M, wt, bias.detach(), out=X[:, ch * C:(ch + 1) * C], c_group=ntok, c_group_stride=S, c_offset=off,
                           tc=True)   # same kernel for every batch size: a row's result never depends on its batch

    def _table_planes(self):
        """BF16x3 operand planes of the (constant) node / edge feature tables, built on first use."""
        if self._planes is None:
            zero = int(bool((self.node_raw_features[0] == 0).all())) | (int(bool((self.edge_raw_features[0] == 0).all())) << 1)
            self._planes = (ops.table_planes(self.node_raw_features), ops.table_planes(self.edge_raw_features), zero)
        return self._planes

    def _projection_operands(self, lut):
        """(LUT planes, packed projection weights, concatenated bias), rebuilt when a parameter changes."""
        pl = self.projection_layer
        ps = [pl[k].weight for k in ('node', 'edge', 'time', 'neighbor_co_occurrence')] + \
             [pl[k].bias for k in ('node', 'edge', 'time', 'neighbor_co_occurrence')]
        key = (lut.data_ptr(), tuple(lut.shape), ops.WEIGHTS_EPOCH) + tuple((q.data_ptr(), q._version) for q in ps)
        if key != self._proj_key:
            C = self.neighbor_co_occurrence_feat_dim
            lut_pl = ops.table_planes(lut[:, :C])
            packed = ops.pack_patch_weights(ps[0], ps[1], ps[2], ps[3], self.patch_size)
            bias = torch.cat([q.detach().float() for q in ps[4:]]).contiguous()
            self._proj_ops, self._proj_key = (lut_pl, packed, bias, lut), key
        return self._proj_ops[:3]

    def set_neighbor_sampler(self, neighbor_sampler: NeighborSampler):
        """``set_neighbor_sampler`` (``models/DyGFormer.py:308-317``)."""
        self.neighbor_sampler = neighbor_sampler
        if self.neighbor_sampler.sample_neighbor_strategy in ['uniform', 'time_interval_aware']:
            assert self.neighbor_sampler.seed is not None
            self.neighbor_sampler.reset_random_state()
