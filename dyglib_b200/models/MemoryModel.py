"""Drop-in for the reference's ``MemoryModel`` (TGN / DyRep / JODIE, ``models/MemoryModel.py``).

The python dict-of-lists message store, the per-call loop over all nodes and the full-table clone of the
reference become flat device tables (see csrc/memory.cu): persisted memory, an incrementally maintained
look-ahead view (what ``get_updated_memories`` would return for every node), and the last raw message of
every node.  Per positive batch only the batch's <= 2B new messages go through the GRU; results are
identical to recomputing all pending nodes each call because a pending node's inputs cannot change before
its message is consumed.  This holds for fixed weights (eval).  In training mode (``.train()`` with grad enabled) the
weights move between batches, so every call recomputes ``get_updated_memories`` for all pending nodes from the stored
last messages with autograd (``models/MemoryModel.py:170-191``), embeds on those rows through ``dyglib_b200/autograd.py``
and persists the recomputed rows; the gradient reaches the memory updater, the time encoder and the embedding module
exactly where the reference's does.
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np
import torch
import torch.nn as nn

from .. import _native, ops
from ..ops import _p, _stream
from ..utils.utils import NeighborSampler, _as_dev
from .modules import TimeEncoder, MergeLayer, MultiHeadAttention
from ._temporal import temporal_conv, temporal_conv_train, zero_time_features


class MessageAggregator(nn.Module):
    """``MessageAggregator`` (``models/MemoryModel.py:267-300``): keep the last message of every node."""

    def __init__(self):
        super().__init__()

    def aggregate_messages(self, node_ids: np.ndarray, node_raw_messages: dict):
        """Compat API over the reference's dict-of-lists format (host bookkeeping only; the fused model path
        selects last messages on the device with dyg_tgn_select_last)."""
        unique_node_ids = np.unique(node_ids)
        msgs, ts, ids = [], [], []
        for v in unique_node_ids:
            lst = node_raw_messages.get(v, []) if isinstance(node_raw_messages, dict) else node_raw_messages[v]
            if len(lst) > 0:
                ids.append(v)
                msgs.append(lst[-1][0])
                ts.append(lst[-1][1])
        return (np.array(ids), torch.stack(msgs, dim=0) if msgs else torch.Tensor([]), np.array(ts))


class MemoryBank(nn.Module):

    def __init__(self, num_nodes: int, memory_dim: int, message_dim: int = 0):
        """``MemoryBank`` (``models/MemoryModel.py:304-422``); ``node_memories`` and ``node_last_updated_times`` are
        non-grad Parameters so they land in the state_dict like the reference's."""
        super().__init__()
        self.num_nodes = num_nodes
        self.memory_dim = memory_dim
        self.message_dim = message_dim
        self.node_memories = nn.Parameter(torch.zeros((num_nodes, memory_dim)), requires_grad=False)
        self.node_last_updated_times = nn.Parameter(torch.zeros(num_nodes), requires_grad=False)
        self._state = None

    def _ensure(self):
        dev = self.node_memories.device
        if self._state is None or self._state['mem_view'].device != dev:
            N, D = self.num_nodes, self.memory_dim
            self._state = dict(
                mem_view=self.node_memories.data.clone(),
                lu_view=self.node_last_updated_times.data.clone(),
                pending=torch.zeros(N, dtype=torch.uint8, device=dev),
                winner=torch.full((N,), -1, dtype=torch.int32, device=dev),
                msg_store=torch.zeros((N, max(self.message_dim, 1)), dtype=torch.float32, device=dev),
                msg_time=torch.zeros(N, dtype=torch.float64, device=dev),
                flag=torch.zeros(1, dtype=torch.int32, device=dev))
        return self._state

    def __init_memory_bank__(self):
        """``__init_memory_bank__`` (``models/MemoryModel.py:325-332``)."""
        self.node_memories.data.zero_()
        self.node_last_updated_times.data.zero_()
        if self._state is None or self._state['mem_view'].device != self.node_memories.device:
            self._state = None
            self._ensure()
        else:   # in place: a captured CUDA graph keeps pointing at these buffers
            st = self._state
            for k in ('mem_view', 'lu_view', 'pending', 'msg_store', 'msg_time', 'flag'):
                st[k].zero_()
            st['winner'].fill_(-1)

    def get_memories(self, node_ids):
        ids = _as_dev(node_ids, torch.int64, self.node_memories.device)
        return ops.gather_rows(self.node_memories.data, ids)

    def set_memories(self, node_ids, updated_node_memories: torch.Tensor):
        ids = _as_dev(node_ids, torch.int64, self.node_memories.device)
        self.node_memories.data[ids] = updated_node_memories
        st = self._ensure()
        st['mem_view'][ids] = updated_node_memories

    def get_node_last_updated_times(self, unique_node_ids):
        ids = _as_dev(unique_node_ids, torch.int64, self.node_memories.device)
        return self.node_last_updated_times.data[ids]

    @property
    def node_raw_messages(self):
        """The pending messages in the reference's format {node_id: [(message, time)]} (export for checkpoints,
        ``utils/EarlyStopping.py:73-86``)."""
        st = self._ensure()
        out = defaultdict(list)
        for v in torch.nonzero(st['pending']).reshape(-1).tolist():
            out[v].append((st['msg_store'][v].clone(), np.float64(st['msg_time'][v].item())))
        return out

    def backup_memory_bank(self):
        """``backup_memory_bank`` (``models/MemoryModel.py:351-360``): flat device copies."""
        st = self._ensure()
        return (self.node_memories.data.clone(), self.node_last_updated_times.data.clone(),
                {k: v.clone() for k, v in st.items()})

    def reload_memory_bank(self, backup_memory_bank: tuple):
        """``reload_memory_bank`` (``models/MemoryModel.py:362-372``)."""
        self.node_memories.data.copy_(backup_memory_bank[0])
        self.node_last_updated_times.data.copy_(backup_memory_bank[1])
        st = self._ensure()
        for k, v in backup_memory_bank[2].items():   # in place (graph-safe)
            st[k].copy_(v)

    def detach_memory_bank(self):
        """``detach_memory_bank`` (``models/MemoryModel.py:374-387``): nothing carries gradients here."""
        self.node_memories.detach_()

    def extra_repr(self):
        return 'num_nodes={}, memory_dim={}'.format(self.node_memories.shape[0], self.node_memories.shape[1])


class MemoryUpdater(nn.Module):
    """``MemoryUpdater`` (``models/MemoryModel.py:425-487``); subclasses own the recurrent cell's parameters."""

    def __init__(self, memory_bank: MemoryBank):
        super().__init__()
        self.memory_bank = memory_bank


class GRUMemoryUpdater(MemoryUpdater):

    def __init__(self, memory_bank: MemoryBank, message_dim: int, memory_dim: int):
        super().__init__(memory_bank)
        self.memory_updater = nn.GRUCell(input_size=message_dim, hidden_size=memory_dim)
        self.gates = 3


class RNNMemoryUpdater(MemoryUpdater):

    def __init__(self, memory_bank: MemoryBank, message_dim: int, memory_dim: int):
        super().__init__(memory_bank)
        self.memory_updater = nn.RNNCell(input_size=message_dim, hidden_size=memory_dim)
        self.gates = 1


class TimeProjectionEmbedding(nn.Module):

    def __init__(self, memory_dim: int, dropout: float):
        """``TimeProjectionEmbedding`` (``models/MemoryModel.py:519-545``)."""
        super().__init__()
        self.memory_dim = memory_dim
        self.dropout = nn.Dropout(dropout)
        self.linear_layer = nn.Linear(1, self.memory_dim)


class GraphAttentionEmbedding(nn.Module):

    def __init__(self, node_raw_features: torch.Tensor, edge_raw_features: torch.Tensor, neighbor_sampler: NeighborSampler,
                 time_encoder: TimeEncoder, node_feat_dim: int, edge_feat_dim: int, time_feat_dim: int,
                 num_layers: int = 2, num_heads: int = 2, dropout: float = 0.1):
        """``GraphAttentionEmbedding`` (``models/MemoryModel.py:548-586``)."""
        super().__init__()
        self.node_raw_features = node_raw_features
        self.edge_raw_features = edge_raw_features
        # the memory table changes every batch: only the (static) edge table's padding row is declared zero
        self._zero_row0 = ops.zero_row0_flags(None, edge_raw_features) & 2
        self.neighbor_sampler = neighbor_sampler
        self.time_encoder = time_encoder
        self.node_feat_dim = node_feat_dim
        self.edge_feat_dim = edge_feat_dim
        self.time_feat_dim = time_feat_dim
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.dropout = dropout
        self.temporal_conv_layers = nn.ModuleList([
            MultiHeadAttention(node_feat_dim, edge_feat_dim, time_feat_dim, num_heads, dropout) for _ in range(num_layers)])
        self.merge_layers = nn.ModuleList([
            MergeLayer(node_feat_dim + time_feat_dim, node_feat_dim, node_feat_dim, node_feat_dim) for _ in range(num_layers)])

    def compute_node_temporal_embeddings(self, node_memories: torch.Tensor, node_ids, node_interact_times,
                                         current_layer_num: int, num_neighbors: int = 20):
        """``compute_node_temporal_embeddings`` (``models/MemoryModel.py:588-664``): layer-0 features are
        memory + raw features; neighbours' layer-0 rows are gathered inside the attention kernel."""
        assert current_layer_num >= 0
        dev = self.node_raw_features.device
        ids = _as_dev(node_ids, torch.int64, dev)
        tq = _as_dev(node_interact_times, torch.float64, dev)
        t0 = zero_time_features(self.time_encoder, dev)
        return self._embed(node_memories, ids, tq, current_layer_num, num_neighbors, t0)

    def embed_train(self, mem, ids, tq, layer, k):
        """The recursion of ``_embed`` with autograd: ``mem`` is the differentiable (num_nodes, memory_dim) table of updated
        memories; layer-0 rows ``mem[ids] + raw[ids]`` are gathered by index (their backward is a scatter-add)."""
        feat = mem[ids] + self.node_raw_features[ids]
        if layer == 0:
            return feat
        conv = feat if layer == 1 else self.embed_train(mem, ids, tq, layer - 1, k)
        nbr, eid, nt = self.neighbor_sampler.get_historical_neighbors_device(ids, tq, k)
        nbr_dense = self.embed_train(mem, nbr.reshape(-1), nt.reshape(-1).double(), layer - 1, k)
        return temporal_conv_train(self.temporal_conv_layers[layer - 1], self.merge_layers[layer - 1], self.time_encoder, conv, feat,
                                   self.node_raw_features, nbr, nbr_dense, self.edge_raw_features, eid, tq, nt, k,
                                   zero_row0=self._zero_row0)

    def _embed(self, mem, ids, tq, layer, k, t0):
        feat = ops.gather_rows(self.node_raw_features, ids, table2=mem)
        if layer == 0:
            return feat
        conv = feat if layer == 1 else self._embed(mem, ids, tq, layer - 1, k, t0)
        nbr, eid, nt = self.neighbor_sampler.get_historical_neighbors_device(ids, tq, k)
        nbr_dense = None
        if layer > 1:
            nbr_dense = self._embed(mem, nbr.reshape(-1), nt.reshape(-1).double(), layer - 1, k, t0)
        return temporal_conv(self.temporal_conv_layers[layer - 1], self.merge_layers[layer - 1], self.time_encoder, t0,
                             conv, feat, self.node_raw_features, mem, nbr, nbr_dense, self.edge_raw_features, eid, tq, nt, k,
                             zero_row0=self._zero_row0)


class MemoryModel(torch.nn.Module):

    def __init__(self, node_raw_features: np.ndarray, edge_raw_features: np.ndarray, neighbor_sampler: NeighborSampler,
                 time_feat_dim: int, model_name: str = 'TGN', num_layers: int = 2, num_heads: int = 2, dropout: float = 0.1,
                 src_node_mean_time_shift: float = 0.0, src_node_std_time_shift: float = 1.0, dst_node_mean_time_shift_dst: float = 0.0,
                 dst_node_std_time_shift: float = 1.0, device: str = 'cuda'):
        """Same arguments as ``MemoryModel.__init__`` (``models/MemoryModel.py:12-85``)."""
        super().__init__()
        self.node_raw_features = torch.from_numpy(node_raw_features.astype(np.float32)).to(device).contiguous()
        self.edge_raw_features = torch.from_numpy(edge_raw_features.astype(np.float32)).to(device).contiguous()
        self.node_feat_dim = self.node_raw_features.shape[1]
        self.edge_feat_dim = self.edge_raw_features.shape[1]
        self.time_feat_dim = time_feat_dim
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.dropout = dropout
        self.device = device
        self.src_node_mean_time_shift = src_node_mean_time_shift
        self.src_node_std_time_shift = src_node_std_time_shift
        self.dst_node_mean_time_shift_dst = dst_node_mean_time_shift_dst
        self.dst_node_std_time_shift = dst_node_std_time_shift
        self.model_name = model_name
        self.num_nodes = self.node_raw_features.shape[0]
        self.memory_dim = self.node_feat_dim
        self.message_dim = self.memory_dim + self.memory_dim + self.time_feat_dim + self.edge_feat_dim
        self.time_encoder = TimeEncoder(time_dim=time_feat_dim)
        self.message_aggregator = MessageAggregator()
        self.memory_bank = MemoryBank(num_nodes=self.num_nodes, memory_dim=self.memory_dim, message_dim=self.message_dim)
        if self.model_name == 'TGN':
            self.memory_updater = GRUMemoryUpdater(self.memory_bank, self.message_dim, self.memory_dim)
        elif self.model_name in ['DyRep', 'JODIE']:
            self.memory_updater = RNNMemoryUpdater(self.memory_bank, self.message_dim, self.memory_dim)
        else:
            raise ValueError(f'Not implemented error for model_name {self.model_name}!')
        if self.model_name == 'JODIE':
            self.embedding_module = TimeProjectionEmbedding(memory_dim=self.memory_dim, dropout=self.dropout)
        else:
            self.embedding_module = GraphAttentionEmbedding(self.node_raw_features, self.edge_raw_features, neighbor_sampler,
                                                            self.time_encoder, self.node_feat_dim, self.edge_feat_dim,
                                                            self.time_feat_dim, self.num_layers, self.num_heads, self.dropout)
        self.check_time_order = True
        self.to(device)

    def compute_src_dst_node_temporal_embeddings(self, src_node_ids: np.ndarray, dst_node_ids: np.ndarray, node_interact_times: np.ndarray,
                                                 edge_ids: np.ndarray, edges_are_positive: bool = True, num_neighbors: int = 20):
        """``compute_src_dst_node_temporal_embeddings`` (``models/MemoryModel.py:87-168``)."""
        if self.training and torch.is_grad_enabled():
            self._view_stale = True
            return self._forward_train(src_node_ids, dst_node_ids, node_interact_times, edge_ids, edges_are_positive, num_neighbors)
        self._refresh_view_if_stale()
        dev = self.node_raw_features.device
        src = _as_dev(src_node_ids, torch.int64, dev)
        dst = _as_dev(dst_node_ids, torch.int64, dev)
        tq = _as_dev(node_interact_times, torch.float64, dev)
        B = src.numel()
        emb, ret = self._embed_eval([src, dst], tq, num_neighbors)
        if edges_are_positive:
            assert edge_ids is not None
            self._advance(src, dst, tq, _as_dev(edge_ids, torch.int64, dev), emb[B:], emb[:B], None)
        return ret[:B], ret[B:]

    def compute_pos_neg_temporal_embeddings(self, src_node_ids, dst_node_ids, neg_dst_node_ids, node_interact_times, edge_ids,
                                            num_neighbors: int = 20):
        """Addition to the reference API: the two calls its loops make per batch (``train_link_prediction.py:236-247``,
        ``evaluate_models_utils.py:60-80``) -- ``compute(src, neg, t, None, False)`` then ``compute(src, dst, t, edge_ids, True)``
        -- in one.  Both read the same memory state, so the 4 B roots go through ONE embedding pass (half the kernel launches
        of a latency-bound step), then the positive batch advances the memory.
        Returns (neg_src_emb, neg_dst_emb, pos_src_emb, pos_dst_emb), equal to the two calls' results."""
        if self.training and torch.is_grad_enabled():
            a, b = self.compute_src_dst_node_temporal_embeddings(src_node_ids, neg_dst_node_ids, node_interact_times, None, False, num_neighbors)
            c, d = self.compute_src_dst_node_temporal_embeddings(src_node_ids, dst_node_ids, node_interact_times, edge_ids, True, num_neighbors)
            return a, b, c, d
        self._refresh_view_if_stale()
        dev = self.node_raw_features.device
        src = _as_dev(src_node_ids, torch.int64, dev)
        dst = _as_dev(dst_node_ids, torch.int64, dev)
        neg = _as_dev(neg_dst_node_ids, torch.int64, dev)
        tq = _as_dev(node_interact_times, torch.float64, dev)
        B = src.numel()
        emb, ret = self._embed_eval([src, neg, src, dst], tq, num_neighbors)
        assert edge_ids is not None
        self._advance(src, dst, tq, _as_dev(edge_ids, torch.int64, dev), emb[3 * B:], emb[2 * B:3 * B], None)
        return ret[:B], ret[B:2 * B], ret[2 * B:3 * B], ret[3 * B:]

    def _refresh_view_if_stale(self):
        if getattr(self, '_view_stale', False):
            # the look-ahead view of the pending nodes was built by earlier weights: rebuild it with the current ones
            with torch.no_grad():
                mem_upd, lu_upd = self._updated_memories_train()
                st0 = self.memory_bank._ensure()
                st0['mem_view'].copy_(mem_upd)
                st0['lu_view'].copy_(lu_upd)
            self._view_stale = False

    def _embed_eval(self, roles, tq, num_neighbors):
        """Embeddings of the concatenated root sets ``roles`` (alternating src / dst roles, each B ids at times ``tq``) on the
        look-ahead view == ``get_updated_memories`` of all nodes (``models/MemoryModel.py:108-136``).  Returns (embeddings,
        what the model returns for them: DyRep returns the look-ahead memories instead, ``:163-166``)."""
        lib = _native.load()
        dev = self.node_raw_features.device
        st = self.memory_bank._ensure()
        D = self.memory_dim
        B = tq.numel()
        node_ids = torch.cat(roles)
        mem_view, lu_view = st['mem_view'], st['lu_view']
        if self.model_name == 'JODIE':
            ll = self.embedding_module.linear_layer
            emb = torch.empty((len(roles) * B, D), dtype=torch.float32, device=dev)
            wv, bv = ll.weight.detach().reshape(-1), ll.bias.detach()
            for r, ids in enumerate(roles):
                mean, std = ((self.src_node_mean_time_shift, self.src_node_std_time_shift) if r % 2 == 0 else
                             (self.dst_node_mean_time_shift_dst, self.dst_node_std_time_shift))
                _native.check(lib.dyg_jodie_project(_p(mem_view), D, _p(lu_view), _p(ids), _p(tq), B, D, float(mean), float(std),
                                                    _p(wv), _p(bv), _p(emb[r * B:(r + 1) * B]), D, _stream()))
                ops._count()
        else:
            emb = self.embedding_module.compute_node_temporal_embeddings(mem_view, node_ids, torch.cat([tq] * len(roles)),
                                                                         self.num_layers, num_neighbors)
        ret = ops.gather_rows(mem_view, node_ids) if self.model_name == 'DyRep' else emb
        return emb, ret

    def _advance(self, src, dst, tq, eid, dst_emb, src_emb, recomputed):
        """update_memories + clear + new raw messages of a positive batch (``models/MemoryModel.py:139-161``).
        ``recomputed``: training mode only, the (num_nodes, D) table of memories recomputed with the current weights; its rows
        replace what ``dyg_tgn_persist`` copies from the look-ahead view (which earlier weights produced)."""
        lib = _native.load()
        dev = self.node_raw_features.device
        bank = self.memory_bank
        st = bank._ensure()
        D, T, E = self.memory_dim, self.time_feat_dim, self.edge_feat_dim
        B = src.numel()
        node_ids = torch.cat([src, dst])
        lu_view = st['lu_view']
        mem_view = st['mem_view']
        mem, lu = bank.node_memories.data, bank.node_last_updated_times.data
        if recomputed is not None:
            was_pending = node_ids[st['pending'][node_ids].bool()]
            mem_view[was_pending] = recomputed[was_pending]
        if self.check_time_order:
            _native.check(lib.dyg_tgn_check_time(_p(node_ids), 2 * B, _p(lu), _p(lu_view), _p(st['pending']), _p(st['flag']), _stream()))
            ops._count()
        # update_memories + clear_node_raw_messages for the batch's nodes (:142-145)
        _native.check(lib.dyg_tgn_persist(_p(node_ids), 2 * B, _p(mem), _p(mem_view), _p(lu), _p(lu_view), _p(st['pending']), D, _stream()))
        ops._count(2)
        # new raw messages, src role then dst role (:148-161); last message per node wins
        _native.check(lib.dyg_tgn_select_last(_p(src), _p(dst), B, _p(st['winner']), _stream()))
        other = torch.cat([dst_emb, src_emb]).detach().contiguous() if self.model_name == 'DyRep' else None
        w, b = self.time_encoder.wb()
        msg = torch.empty((2 * B, self.message_dim), dtype=torch.float32, device=dev)
        _native.check(lib.dyg_tgn_build_messages(_p(src), _p(dst), _p(tq), _p(eid), B, _p(mem), _p(lu), D, _p(other), D,
                                                 _p(self.edge_raw_features), self.edge_raw_features.stride(0), E,
                                                 _p(w), _p(b), T, _p(msg), self.message_dim, _stream()))
        cell = self.memory_updater.memory_updater
        G = self.memory_updater.gates
        gi = ops.linear([ops.seg_rows(msg)], 2 * B, cell.weight_ih.detach(), cell.bias_ih.detach())
        gh = ops.linear([ops.seg_rows(mem, D, node_ids)], 2 * B, cell.weight_hh.detach(), cell.bias_hh.detach())
        _native.check(lib.dyg_tgn_cell_commit(_p(gi), _p(gh), G, _p(src), _p(dst), _p(tq), B, _p(st['winner']), _p(mem),
                                              _p(mem_view), _p(lu_view), _p(st['pending']), D, _p(msg), self.message_dim,
                                              self.message_dim, _p(st['msg_store']), _p(st['msg_time']), _stream()))
        ops._count(3)

    def _updated_memories_train(self):
        """``get_updated_memories`` over all nodes (``models/MemoryModel.py:170-191, 461-487``) with autograd: the recurrent
        cell over every pending node's last message (gate GEMMs on ``autograd.linear``, gate math elementwise)."""
        from .. import autograd as ag
        bank = self.memory_bank
        st = bank._ensure()
        mem, lu = bank.node_memories.data, bank.node_last_updated_times.data
        pend = torch.nonzero(st['pending']).reshape(-1)
        if pend.numel() == 0:
            return mem, lu
        cell = self.memory_updater.memory_updater
        h = mem[pend]
        gi = ag.linear(st['msg_store'][pend], cell.weight_ih, cell.bias_ih)
        gh = ag.linear(h, cell.weight_hh, cell.bias_hh)
        if self.memory_updater.gates == 3:      # nn.GRUCell: gates r | z | n
            i_r, i_z, i_n = gi.chunk(3, dim=1)
            h_r, h_z, h_n = gh.chunk(3, dim=1)
            r = torch.sigmoid(i_r + h_r)
            z = torch.sigmoid(i_z + h_z)
            n = torch.tanh(i_n + r * h_n)
            new = (1.0 - z) * n + z * h
        else:                                   # nn.RNNCell (tanh)
            new = torch.tanh(gi + gh)
        return mem.index_copy(0, pend, new), lu.index_copy(0, pend, st['msg_time'][pend].float())

    def _forward_train(self, src_node_ids, dst_node_ids, node_interact_times, edge_ids, edges_are_positive, num_neighbors):
        """Training-mode ``compute_src_dst_node_temporal_embeddings``."""
        dev = self.node_raw_features.device
        src = _as_dev(src_node_ids, torch.int64, dev)
        dst = _as_dev(dst_node_ids, torch.int64, dev)
        tq = _as_dev(node_interact_times, torch.float64, dev)
        B = src.numel()
        node_ids = torch.cat([src, dst])
        mem_upd, lu_upd = self._updated_memories_train()
        if self.model_name == 'JODIE':
            em = self.embedding_module
            tt = tq.float()
            s_iv = (tt - lu_upd[src] - self.src_node_mean_time_shift) / self.src_node_std_time_shift
            d_iv = (tt - lu_upd[dst] - self.dst_node_mean_time_shift_dst) / self.dst_node_std_time_shift
            iv = torch.cat([s_iv, d_iv]).unsqueeze(1)
            emb = em.dropout(mem_upd[node_ids] * (1 + em.linear_layer(iv)))      # models/MemoryModel.py:543
        else:
            emb = self.embedding_module.embed_train(mem_upd, node_ids, torch.cat([tq, tq]), self.num_layers, num_neighbors)
        src_emb, dst_emb = emb[:B], emb[B:]
        if edges_are_positive:
            assert edge_ids is not None
            self._advance(src, dst, tq, _as_dev(edge_ids, torch.int64, dev), dst_emb, src_emb, mem_upd.detach())
        if self.model_name == 'DyRep':
            return mem_upd[src], mem_upd[dst]
        return src_emb, dst_emb

    def assert_time_order(self):
        """Raises if any update went backwards in time (the reference asserts per call, ``:448-449``; here the
        device flag is read on demand to avoid a host sync per batch)."""
        assert int(self.memory_bank._ensure()['flag'].item()) == 0, 'Trying to update memory to time in the past!'

    def set_neighbor_sampler(self, neighbor_sampler: NeighborSampler):
        """``set_neighbor_sampler`` (``models/MemoryModel.py:253-263``)."""
        assert self.model_name in ['TGN', 'DyRep'], f'Neighbor sampler is not defined in model {self.model_name}!'
        self.embedding_module.neighbor_sampler = neighbor_sampler
        if self.embedding_module.neighbor_sampler.sample_neighbor_strategy in ['uniform', 'time_interval_aware']:
            assert self.embedding_module.neighbor_sampler.seed is not None
            self.embedding_module.neighbor_sampler.reset_random_state()


def compute_src_dst_node_time_shifts(src_node_ids: np.ndarray, dst_node_ids: np.ndarray, node_interact_times: np.ndarray):
    """``compute_src_dst_node_time_shifts`` (``models/MemoryModel.py:667-698``): mean / std of the gaps between
    consecutive interactions of each src (resp. dst) node, first gap measured from time 0.  Host preprocessing,
    vectorised with a stable sort instead of the reference's python dict loop."""
    def shifts(ids, t):
        order = np.argsort(ids, kind='stable')
        si, st = ids[order], t[order]
        prev = np.empty_like(st)
        prev[1:] = st[:-1]
        first = np.ones(len(si), dtype=bool)
        first[1:] = si[1:] != si[:-1]
        prev[first] = 0
        out = np.empty_like(st)
        out[order] = st - prev
        return out
    t = np.asarray(node_interact_times, dtype=np.float64)
    s = shifts(np.asarray(src_node_ids), t)
    d = shifts(np.asarray(dst_node_ids), t)
    return np.mean(s), np.std(s), np.mean(d), np.std(d)
