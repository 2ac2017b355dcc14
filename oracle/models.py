"""Oracle restatement (torch CPU, fp32) of the float part of the path:
TimeEncoder, temporal MultiHeadAttention, MergeLayer, TGAT, DyGFormer, MemoryModel(TGN / DyRep / JODIE), GraphMixer, TCL.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Eval-mode semantics
(dropout = identity).  Weights come in as a ``state_dict`` with the reference's
parameter names, so the same dict drives the reference, the oracle and the CUDA path.
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np
import torch
import torch.nn.functional as F

from .sampler import OracleSampler, pad_sequences, count_nodes_appearances


def _ids(x):
    return torch.from_numpy(np.ascontiguousarray(x).astype(np.int64))


def time_encode(sd, prefix, dt):
    """``TimeEncoder.forward`` (``models/modules.py:27-39``): cos(Linear(1->T)(dt)); dt (n, L) f32."""
    # nn.Linear(1, T) on the reference's CPU path evaluates each output as ONE fp32 fused multiply-add (SURVEY.md A.3:
    # 400 000 / 400 000 matches against fma, 20 % mismatches against mul-then-add).  Whether torch's CPU kernels fuse
    # depends on the host, so the oracle states the FMA explicitly: the fp32 x fp32 product is exact in float64, one
    # float64 add, one rounding to fp32 (double rounding can differ from a true FMA only in ~1e-8 of the cases).
    w = sd[prefix + 'w.weight'].reshape(-1).double()
    b = sd[prefix + 'w.bias'].double()
    arg = (dt.double().unsqueeze(-1) * w + b).float()
    return torch.cos(arg)


def temporal_attention(sd, prefix, node_features, node_time_features, nbr_features, nbr_time_features,
                       nbr_edge_features, nbr_ids, num_heads=2):
    """``MultiHeadAttention.forward`` (``models/modules.py:137-206``)."""
    n, k = nbr_ids.shape
    res = torch.cat([node_features.unsqueeze(1), node_time_features], dim=2)          # (n,1,Dq)
    dq = res.shape[2]
    hd = dq // num_heads
    kv_in = torch.cat([nbr_features, nbr_edge_features, nbr_time_features], dim=2)      # (n,k,Dk)
    q = F.linear(res, sd[prefix + 'query_projection.weight']).reshape(n, 1, num_heads, hd).permute(0, 2, 1, 3)
    kk = F.linear(kv_in, sd[prefix + 'key_projection.weight']).reshape(n, k, num_heads, hd).permute(0, 2, 1, 3)
    vv = F.linear(kv_in, sd[prefix + 'value_projection.weight']).reshape(n, k, num_heads, hd).permute(0, 2, 1, 3)
    att = torch.einsum('bhld,bhnd->bhln', q, kk) * (hd ** -0.5)
    pad = (_ids(nbr_ids) == 0).reshape(n, 1, 1, k).expand(n, num_heads, 1, k)
    att = att.masked_fill(pad, -1e10)                                                 # -1e10, not -inf (:184)
    scores = torch.softmax(att, dim=-1)
    o = torch.einsum('bhln,bhnd->bhld', scores, vv).permute(0, 2, 1, 3).flatten(start_dim=2)
    o = F.linear(o, sd[prefix + 'residual_fc.weight'], sd[prefix + 'residual_fc.bias'])
    o = F.layer_norm(o + res, (dq,), sd[prefix + 'layer_norm.weight'], sd[prefix + 'layer_norm.bias'])
    return o.squeeze(1), scores.squeeze(2)


def merge_layer(sd, prefix, a, b):
    """``MergeLayer.forward`` (``models/modules.py:57-68``)."""
    h = F.relu(F.linear(torch.cat([a, b], dim=1), sd[prefix + 'fc1.weight'], sd[prefix + 'fc1.bias']))
    return F.linear(h, sd[prefix + 'fc2.weight'], sd[prefix + 'fc2.bias'])


class OracleTGAT:
    """``TGAT`` (``models/TGAT.py:48-136``)."""

    def __init__(self, sd, node_raw_features, edge_raw_features, sampler: OracleSampler, num_layers=2, num_heads=2):
        self.sd = sd
        self.nf = torch.from_numpy(node_raw_features.astype(np.float32))
        self.ef = torch.from_numpy(edge_raw_features.astype(np.float32))
        self.sampler = sampler
        self.num_layers = num_layers
        self.num_heads = num_heads

    def node_embeddings(self, node_ids, times, layer, k):
        sd = self.sd
        raw = self.nf[_ids(node_ids)]
        if layer == 0:
            return raw
        t0 = time_encode(sd, 'time_encoder.', torch.zeros(len(node_ids), 1))
        conv = self.node_embeddings(node_ids, times, layer - 1, k)
        nn_, ne_, nt_ = self.sampler.get_historical_neighbors(node_ids, times, k)
        # hop >= 2 queries use the float32-rounded neighbour times (models/TGAT.py:107-110)
        nconv = self.node_embeddings(nn_.flatten(), nt_.flatten(), layer - 1, k).reshape(len(node_ids), k, -1)
        dt = torch.from_numpy(np.asarray(times)[:, None] - nt_).float()               # (:116-119)
        te = time_encode(sd, 'time_encoder.', dt)
        out, _ = temporal_attention(sd, f'temporal_conv_layers.{layer - 1}.', conv, t0, nconv, te,
                                    self.ef[_ids(ne_)], nn_, self.num_heads)
        return merge_layer(sd, f'merge_layers.{layer - 1}.', out, raw)

    def compute_src_dst_node_temporal_embeddings(self, src, dst, times, num_neighbors=20):
        return (self.node_embeddings(src, times, self.num_layers, num_neighbors),
                self.node_embeddings(dst, times, self.num_layers, num_neighbors))


class OracleDyGFormer:
    """``DyGFormer`` (``models/DyGFormer.py:68-194``, ``:247-306``, ``:395-461``)."""

    def __init__(self, sd, node_raw_features, edge_raw_features, sampler: OracleSampler, channel_dim=50,
                 patch_size=1, num_layers=2, num_heads=2, max_input_sequence_length=512):
        self.sd = sd
        self.nf = torch.from_numpy(node_raw_features.astype(np.float32))
        self.ef = torch.from_numpy(edge_raw_features.astype(np.float32))
        self.sampler = sampler
        self.C = channel_dim
        self.P = patch_size
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.L = max_input_sequence_length

    def padded(self, node_ids, times):
        ln, le, lt = self.sampler.get_all_first_hop_neighbors(node_ids, times)
        return pad_sequences(node_ids, times, ln, le, lt, self.P, self.L)

    def cooc_features(self, src_ids, dst_ids):
        sd = self.sd
        p = 'neighbor_co_occurrence_encoder.neighbor_co_occurrence_encode_layer.'
        outs = []
        for c in count_nodes_appearances(src_ids, dst_ids):
            c = torch.from_numpy(c).unsqueeze(-1)                                      # (B,L,2,1)
            h = F.relu(F.linear(c, sd[p + '0.weight'], sd[p + '0.bias']))
            outs.append(F.linear(h, sd[p + '2.weight'], sd[p + '2.bias']).sum(dim=2))  # (:409-411)
        return outs

    def features(self, times, pn, pe, pt):
        nodef = self.nf[_ids(pn)]
        edgef = self.ef[_ids(pe)]
        te = time_encode(self.sd, 'time_encoder.', torch.from_numpy(np.asarray(times)[:, None] - pt).float())
        te[_ids(pn) == 0] = 0.0                                                         # (:266)
        return nodef, edgef, te

    def patches(self, x):
        B, L, Fd = x.shape
        return x.reshape(B, L // self.P, self.P * Fd)                                  # (:298-304)

    def transformer(self, i, x):
        """``TransformerEncoder.forward`` (``models/DyGFormer.py:442-461``); x (B,S,D); no padding mask."""
        sd, p = self.sd, f'transformers.{i}.'
        B, S, D = x.shape
        H = self.num_heads
        hd = D // H
        y = F.layer_norm(x, (D,), sd[p + 'norm_layers.0.weight'], sd[p + 'norm_layers.0.bias'])
        qkv = F.linear(y, sd[p + 'multi_head_attention.in_proj_weight'], sd[p + 'multi_head_attention.in_proj_bias'])
        q, k, v = qkv.split(D, dim=2)
        q = q.reshape(B, S, H, hd).transpose(1, 2) * (hd ** -0.5)
        k = k.reshape(B, S, H, hd).transpose(1, 2)
        v = v.reshape(B, S, H, hd).transpose(1, 2)
        a = torch.softmax(q @ k.transpose(2, 3), dim=-1)
        o = (a @ v).transpose(1, 2).reshape(B, S, D)
        o = F.linear(o, sd[p + 'multi_head_attention.out_proj.weight'], sd[p + 'multi_head_attention.out_proj.bias'])
        x = x + o
        y = F.layer_norm(x, (D,), sd[p + 'norm_layers.1.weight'], sd[p + 'norm_layers.1.bias'])
        h = F.gelu(F.linear(y, sd[p + 'linear_layers.0.weight'], sd[p + 'linear_layers.0.bias']))
        return x + F.linear(h, sd[p + 'linear_layers.1.weight'], sd[p + 'linear_layers.1.bias'])

    def compute_src_dst_node_temporal_embeddings(self, src, dst, times):
        sd = self.sd
        s_pn, s_pe, s_pt = self.padded(src, times)
        d_pn, d_pe, d_pt = self.padded(dst, times)
        s_co, d_co = self.cooc_features(s_pn, d_pn)
        toks = []
        for pn, pe, pt, co in ((s_pn, s_pe, s_pt, s_co), (d_pn, d_pe, d_pt, d_co)):
            nodef, edgef, te = self.features(times, pn, pe, pt)
            chans = []
            for name, x in (('node', nodef), ('edge', edgef), ('time', te), ('neighbor_co_occurrence', co)):
                chans.append(F.linear(self.patches(x), sd[f'projection_layer.{name}.weight'],
                                      sd[f'projection_layer.{name}.bias']))
            toks.append(torch.stack(chans, dim=2).reshape(len(src), -1, 4 * self.C))
        ns = toks[0].shape[1]
        x = torch.cat(toks, dim=1)
        for i in range(self.num_layers):
            x = self.transformer(i, x)
        s = x[:, :ns].mean(dim=1)
        d = x[:, ns:].mean(dim=1)
        return (F.linear(s, sd['output_layer.weight'], sd['output_layer.bias']),
                F.linear(d, sd['output_layer.weight'], sd['output_layer.bias']))


class OracleMemoryModel:
    """``MemoryModel`` for TGN / DyRep / JODIE (``models/MemoryModel.py:87-251``, ``:275-300``, ``:435-487``,
    ``:588-664``).  The pending-message store keeps the reference's dict-of-lists form."""

    def __init__(self, sd, node_raw_features, edge_raw_features, sampler: OracleSampler, model_name='TGN',
                 num_layers=1, num_heads=2, src_mean=0.0, src_std=1.0, dst_mean=0.0, dst_std=1.0):
        self.sd = sd
        self.nf = torch.from_numpy(node_raw_features.astype(np.float32))
        self.ef = torch.from_numpy(edge_raw_features.astype(np.float32))
        self.sampler = sampler
        self.model_name = model_name
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.num_nodes = self.nf.shape[0]
        self.shift = (src_mean, src_std, dst_mean, dst_std)
        self.reset()

    def reset(self):
        """``MemoryBank.__init_memory_bank__`` (``:325-332``)."""
        self.memory = torch.zeros(self.num_nodes, self.nf.shape[1])
        self.last_update = torch.zeros(self.num_nodes)
        self.raw_messages = defaultdict(list)

    def _cell(self, msg, mem):
        sd, p = self.sd, 'memory_updater.memory_updater.'
        if self.model_name == 'TGN':
            return torch.gru_cell(msg, mem, sd[p + 'weight_ih'], sd[p + 'weight_hh'], sd[p + 'bias_ih'], sd[p + 'bias_hh'])
        return torch.rnn_tanh_cell(msg, mem, sd[p + 'weight_ih'], sd[p + 'weight_hh'], sd[p + 'bias_ih'], sd[p + 'bias_hh'])

    def _aggregate(self, node_ids):
        """``MessageAggregator.aggregate_messages`` (``:275-300``): last appended message per node."""
        ids, msgs, ts = [], [], []
        for v in np.unique(node_ids):
            lst = self.raw_messages[v]
            if len(lst) > 0:
                ids.append(v)
                msgs.append(lst[-1][0])
                ts.append(lst[-1][1])
        return np.array(ids, dtype=np.int64), (torch.stack(msgs) if msgs else torch.zeros(0)), np.array(ts)

    def _updated_view(self):
        """``get_updated_memories`` over all nodes (``:170-191``, ``:461-487``): nothing persisted."""
        ids, msgs, ts = self._aggregate(np.arange(self.num_nodes))
        mem, lu = self.memory.clone(), self.last_update.clone()
        if len(ids) > 0:
            assert bool((self.last_update[_ids(ids)] <= torch.from_numpy(ts).float()).all())
            mem[_ids(ids)] = self._cell(msgs, mem[_ids(ids)])
            lu[_ids(ids)] = torch.from_numpy(ts).float()
        return mem, lu

    def _persist(self, node_ids):
        """``update_memories`` (``:193-210``, ``:435-459``)."""
        ids, msgs, ts = self._aggregate(node_ids)
        if len(ids) == 0:
            return
        assert bool((self.last_update[_ids(ids)] <= torch.from_numpy(ts).float()).all())
        self.memory[_ids(ids)] = self._cell(msgs, self.memory[_ids(ids)])
        self.last_update[_ids(ids)] = torch.from_numpy(ts).float()

    def _embed(self, mem, node_ids, times, layer, k):
        """``GraphAttentionEmbedding.compute_node_temporal_embeddings`` (``:588-664``)."""
        sd = self.sd
        feat = mem[_ids(node_ids)] + self.nf[_ids(node_ids)]
        if layer == 0:
            return feat
        t0 = time_encode(sd, 'time_encoder.', torch.zeros(len(node_ids), 1))
        conv = self._embed(mem, node_ids, times, layer - 1, k)
        nn_, ne_, nt_ = self.sampler.get_historical_neighbors(node_ids, times, k)
        nconv = self._embed(mem, nn_.flatten(), nt_.flatten(), layer - 1, k).reshape(len(node_ids), k, -1)
        te = time_encode(sd, 'time_encoder.', torch.from_numpy(np.asarray(times)[:, None] - nt_).float())
        out, _ = temporal_attention(sd, f'embedding_module.temporal_conv_layers.{layer - 1}.', conv, t0, nconv, te,
                                    self.ef[_ids(ne_)], nn_, self.num_heads)
        return merge_layer(sd, f'embedding_module.merge_layers.{layer - 1}.', out, feat)

    def _new_messages(self, a_ids, b_ids, b_emb, times, edge_ids):
        """``compute_new_node_raw_messages`` (``:212-251``) for role a (message owner) / b (other end)."""
        ma = self.memory[_ids(a_ids)]
        mb = b_emb if self.model_name == 'DyRep' else self.memory[_ids(b_ids)]
        dt = torch.from_numpy(np.asarray(times)).float() - self.last_update[_ids(a_ids)]
        te = time_encode(self.sd, 'time_encoder.', dt.unsqueeze(1)).reshape(len(a_ids), -1)
        msg = torch.cat([ma, mb, te, self.ef[_ids(edge_ids)]], dim=1)
        new = defaultdict(list)
        for i in range(len(a_ids)):
            new[a_ids[i]].append((msg[i], times[i]))
        return np.unique(a_ids), new

    def compute_src_dst_node_temporal_embeddings(self, src, dst, times, edge_ids, edges_are_positive=True,
                                                 num_neighbors=20):
        node_ids = np.concatenate([src, dst])
        mem, lu = self._updated_view()
        if self.model_name == 'JODIE':
            sd = self.sd
            tt = torch.from_numpy(np.asarray(times)).float()
            s_iv = (tt - lu[_ids(src)] - self.shift[0]) / self.shift[1]
            d_iv = (tt - lu[_ids(dst)] - self.shift[2]) / self.shift[3]
            iv = torch.cat([s_iv, d_iv]).unsqueeze(1)
            emb = mem[_ids(node_ids)] * (1 + F.linear(iv, sd['embedding_module.linear_layer.weight'],
                                                      sd['embedding_module.linear_layer.bias']))   # (:543)
        else:
            emb = self._embed(mem, node_ids, np.concatenate([times, times]), self.num_layers, num_neighbors)
        s_emb, d_emb = emb[:len(src)], emb[len(src):]
        if edges_are_positive:
            self._persist(node_ids)
            for v in node_ids:                                                        # clear (:400-407)
                self.raw_messages[v] = []
            u1, m1 = self._new_messages(src, dst, d_emb, times, edge_ids)
            u2, m2 = self._new_messages(dst, src, s_emb, times, edge_ids)
            for v in u1:                                                              # src role first (:160-161)
                self.raw_messages[v].extend(m1[v])
            for v in u2:
                self.raw_messages[v].extend(m2[v])
        if self.model_name == 'DyRep':
            s_emb, d_emb = mem[_ids(src)], mem[_ids(dst)]
        return s_emb, d_emb


class OracleGraphMixer:
    """``GraphMixer`` (``models/GraphMixer.py:58-151``, ``:164-238``), eval-mode semantics."""

    def __init__(self, sd, node_raw_features, edge_raw_features, sampler: OracleSampler, num_layers=2):
        self.sd = sd
        self.nf = torch.from_numpy(node_raw_features.astype(np.float32))
        self.ef = torch.from_numpy(edge_raw_features.astype(np.float32))
        self.sampler = sampler
        self.num_layers = num_layers

    def _ffn(self, prefix, x):
        sd = self.sd
        h = F.gelu(F.linear(x, sd[prefix + 'ffn.0.weight'], sd[prefix + 'ffn.0.bias']))
        return F.linear(h, sd[prefix + 'ffn.3.weight'], sd[prefix + 'ffn.3.bias'])

    def _mixer(self, i, x):
        sd, p = self.sd, f'mlp_mixers.{i}.'
        k, c = x.shape[1], x.shape[2]
        h = F.layer_norm(x.permute(0, 2, 1), (k,), sd[p + 'token_norm.weight'], sd[p + 'token_norm.bias'])
        out = self._ffn(p + 'token_feedforward.', h).permute(0, 2, 1) + x
        h = F.layer_norm(out, (c,), sd[p + 'channel_norm.weight'], sd[p + 'channel_norm.bias'])
        return self._ffn(p + 'channel_feedforward.', h) + out

    def node_embeddings(self, node_ids, times, k=20, time_gap=2000):
        sd = self.sd
        nn_, ne_, nt_ = self.sampler.get_historical_neighbors(node_ids, times, k)
        te = time_encode(sd, 'time_encoder.', torch.from_numpy(np.asarray(times)[:, None] - nt_).float())
        te = te * (_ids(nn_) != 0).unsqueeze(-1)                                                     # (:103-104)
        x = F.linear(torch.cat([self.ef[_ids(ne_)], te], dim=-1), sd['projection_layer.weight'], sd['projection_layer.bias'])
        for i in range(self.num_layers):
            x = self._mixer(i, x)
        link = x.mean(dim=1)
        gap, _, _ = self.sampler.get_historical_neighbors(node_ids, times, time_gap)
        mask = torch.from_numpy((gap > 0).astype(np.float32))
        mask[mask == 0] = -1e10
        scores = torch.softmax(mask, dim=1)
        agg = torch.mean(self.nf[_ids(gap)] * scores.unsqueeze(-1), dim=1)                          # (:139)
        node = agg + self.nf[_ids(node_ids)]
        return F.linear(torch.cat([link, node], dim=1), sd['output_layer.weight'], sd['output_layer.bias'])

    def compute_src_dst_node_temporal_embeddings(self, src, dst, times, num_neighbors=20, time_gap=2000):
        return (self.node_embeddings(src, times, num_neighbors, time_gap), self.node_embeddings(dst, times, num_neighbors, time_gap))


class OracleTCL:
    """``TCL`` (``models/TCL.py:56-183``) with the post-norm ``TransformerEncoder`` of ``models/modules.py:209-266``; the attention
    itself is delegated to ``F.multi_head_attention_forward`` (what ``nn.MultiheadAttention`` calls in the reference)."""

    def __init__(self, sd, node_raw_features, edge_raw_features, sampler: OracleSampler, num_layers=2, num_heads=2):
        self.sd = sd
        self.nf = torch.from_numpy(node_raw_features.astype(np.float32))
        self.ef = torch.from_numpy(edge_raw_features.astype(np.float32))
        self.sampler = sampler
        self.num_layers = num_layers
        self.num_heads = num_heads

    def _block(self, i, q_in, k_in, key_ids):
        sd, p = self.sd, f'transformers.{i}.'
        D = q_in.shape[2]
        pad = _ids(key_ids) == 0
        h, _ = F.multi_head_attention_forward(
            q_in.transpose(0, 1), k_in.transpose(0, 1), k_in.transpose(0, 1), D, self.num_heads,
            sd[p + 'multi_head_attention.in_proj_weight'], sd[p + 'multi_head_attention.in_proj_bias'], None, None, False, 0.0,
            sd[p + 'multi_head_attention.out_proj.weight'], sd[p + 'multi_head_attention.out_proj.bias'], training=False,
            key_padding_mask=pad, need_weights=False)
        out = F.layer_norm(q_in + h.transpose(0, 1), (D,), sd[p + 'norm_layers.0.weight'], sd[p + 'norm_layers.0.bias'])
        h = F.linear(F.relu(F.linear(out, sd[p + 'linear_layers.0.weight'], sd[p + 'linear_layers.0.bias'])),
                     sd[p + 'linear_layers.1.weight'], sd[p + 'linear_layers.1.bias'])
        return F.layer_norm(out + h, (D,), sd[p + 'norm_layers.1.weight'], sd[p + 'norm_layers.1.bias'])

    def _sequence(self, node_ids, times, k):
        sd = self.sd
        nn_, ne_, nt_ = self.sampler.get_historical_neighbors(node_ids, times, k)
        ids = np.concatenate((np.asarray(node_ids)[:, None], nn_), axis=1)
        eids = np.concatenate((np.zeros((len(node_ids), 1), dtype=np.int64), ne_), axis=1)
        ts = np.concatenate((np.asarray(times)[:, None], nt_), axis=1)
        te = time_encode(sd, 'time_encoder.', torch.from_numpy(np.asarray(times)[:, None] - ts).float())
        x = (F.linear(self.nf[_ids(ids)], sd['projection_layer.node.weight'], sd['projection_layer.node.bias']) +
             F.linear(self.ef[_ids(eids)], sd['projection_layer.edge.weight'], sd['projection_layer.edge.bias']) +
             F.linear(te, sd['projection_layer.time.weight'], sd['projection_layer.time.bias']) + sd['depth_embedding.weight'])
        return ids, x

    def compute_src_dst_node_temporal_embeddings(self, src, dst, times, num_neighbors=20):
        sd = self.sd
        s_ids, s = self._sequence(src, times, num_neighbors)
        d_ids, d = self._sequence(dst, times, num_neighbors)
        for i in range(self.num_layers):
            s = self._block(i, s, s, s_ids)
            d = self._block(i, d, d, d_ids)
            s, d = self._block(i, s, d, d_ids), self._block(i, d, s, s_ids)
        return (F.linear(s[:, 0], sd['output_layer.weight'], sd['output_layer.bias']),
                F.linear(d[:, 0], sd['output_layer.weight'], sd['output_layer.bias']))
