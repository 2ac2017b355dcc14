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

import ctypes
import os
from collections import defaultdict
from collections.abc import MutableMapping

import numpy as np
import torch
import torch.nn as nn

from .. import _native, ops
from ..ops import _p, _stream
from ..utils.utils import NeighborSampler, _as_dev
from .modules import TimeEncoder, MergeLayer, MultiHeadAttention
from ._temporal import temporal_conv, temporal_conv_train, zero_time_features, query_constant


def gru_update(cell, gates, msg, msg_idx, hid, hid_idx, out, out_idx, winner=None, save_gates=None):
    """One launch of dyg_gru_update_fwd: rows of ``out`` = cell(msg rows, hid rows) (``nn.GRUCell`` / ``nn.RNNCell``)."""
    P = int(hid_idx.numel() if hid_idx is not None else (msg_idx.numel() if msg_idx is not None else msg.shape[0]))
    if P == 0:
        return out
    if out.data_ptr() == hid.data_ptr():
        raise ValueError('gru_update: out must not alias hid')
    D = hid.shape[1]
    w_ih, w_hh = cell.weight_ih.detach(), cell.weight_hh.detach()
    with ops._Timed('gru_update_kernel', 2.0 * P * gates * D * (msg.shape[1] + D), 4.0 * P * (msg.shape[1] + 2 * D)):
        _native.check(_native.load().dyg_gru_update_fwd(
            _p(msg), int(msg.stride(0)), _p(msg_idx), int(msg.shape[1]), _p(hid), int(hid.stride(0)), _p(hid_idx), int(D),
            _p(w_ih), _p(cell.bias_ih.detach()), _p(w_hh), _p(cell.bias_hh.detach()), int(gates), _p(winner), _p(out), int(out.stride(0)),
            _p(out_idx), _p(save_gates), P, _stream()))
    ops._count()
    return out


class MessageAggregator(nn.Module):
    """``MessageAggregator`` (``models/MemoryModel.py:267-300``): keep the last message of every node."""

    def __init__(self):
        super().__init__()

    def aggregate_messages(self, node_ids: np.ndarray, node_raw_messages: dict):
        """Compat API over the reference's dict-of-lists format; with the bank's own device-backed mapping the selection is one
        gather (the fused model path never calls this: it selects last messages on the device with dyg_tgn_select_last)."""
        unique_node_ids = np.unique(np.asarray(node_ids))
        if isinstance(node_raw_messages, RawMessageStore):
            st = node_raw_messages._bank._ensure()
            dev = st['pending'].device
            cand = torch.from_numpy(unique_node_ids.astype(np.int64)).to(dev)
            ids = cand[st['pending'][cand].bool()]
            if ids.numel() == 0:
                return np.array([]), torch.Tensor([]), np.array([])
            return ids.cpu().numpy(), ops.gather_rows(st['msg_store'], ids), st['msg_time'][ids].cpu().numpy()
        msgs, ts, ids = [], [], []
        for v in unique_node_ids:
            lst = node_raw_messages.get(v, []) if isinstance(node_raw_messages, dict) else node_raw_messages[v]
            if len(lst) > 0:
                ids.append(v)
                msgs.append(lst[-1][0])
                ts.append(lst[-1][1])
        return (np.array(ids), torch.stack(msgs, dim=0) if msgs else torch.Tensor([]), np.array(ts))


class _MessageList(list):
    """``node_raw_messages[node]``: the (message, time) tuples of one node; mutations write through to the device store."""

    def __init__(self, store, node, items):
        super().__init__(items)
        self._store, self._node = store, node

    def append(self, item):
        super().append(item)
        self._store[self._node] = list(self)

    def extend(self, items):
        super().extend(items)
        self._store[self._node] = list(self)

    def clear(self):
        super().clear()
        self._store[self._node] = []


class RawMessageStore(MutableMapping):
    """``MemoryBank.node_raw_messages`` (``models/MemoryModel.py:321-322``: ``{node_id: [(message, time), ...]}``) as a
    mutable mapping over the device tables.  Only a node's LAST message ever reaches the memory updater
    (``MessageAggregator.aggregate_messages``, ``:287-291``) and ``clear_node_raw_messages`` drops the whole list, so the
    store keeps exactly that one row per node: reading gives a one-element list, assigning a list keeps its last entry.
    It pickles as the reference's plain ``defaultdict(list)`` (``torch.save`` in ``utils/EarlyStopping.py:73-75``)."""

    def __init__(self, bank):
        self._bank = bank

    def _node(self, key):
        v = int(key)
        if not 0 <= v < self._bank.num_nodes:
            raise KeyError(key)
        return v

    def __getitem__(self, key):
        v = self._node(key)
        st = self._bank._ensure()
        if not bool(st['pending'][v].item()):
            return _MessageList(self, v, [])          # defaultdict(list) semantics
        return _MessageList(self, v, [(st['msg_store'][v].clone(), np.float64(st['msg_time'][v].item()))])

    def __setitem__(self, key, messages):
        self._bank._load_messages({self._node(key): messages})

    def __delitem__(self, key):
        self._bank._load_messages({self._node(key): []})

    def _pending(self):
        return torch.nonzero(self._bank._ensure()['pending']).reshape(-1)

    def __iter__(self):
        return iter(self._pending().tolist())

    def __len__(self):
        return int(self._pending().numel())

    def __contains__(self, key):
        try:
            v = self._node(key)
        except (KeyError, TypeError, ValueError):
            return False
        return bool(self._bank._ensure()['pending'][v].item())

    def items(self):
        """One gather for all pending nodes (the loops of ``evaluate_link_prediction.py:152-156`` / ``backup_memory_bank``)."""
        st = self._bank._ensure()
        ids = self._pending()
        rows = ops.gather_rows(st['msg_store'], ids) if ids.numel() else None
        times = st['msg_time'][ids].cpu().numpy()
        return [(v, _MessageList(self, v, [(rows[i], np.float64(times[i]))])) for i, v in enumerate(ids.tolist())]

    def values(self):
        return [m for _, m in self.items()]

    def to_dict(self):
        out = defaultdict(list)
        for v, m in self.items():
            out[v] = list(m)
        return out

    def __reduce__(self):
        return (defaultdict, (list,), None, None, iter([(v, list(m)) for v, m in self.items()]))

    def __repr__(self):
        return f'RawMessageStore({len(self)} pending nodes)'


_ANY_WEIGHTS = 'any'    # view key of a bank without pending messages: the look-ahead view equals the persisted state


class MemoryBank(nn.Module):

    def __init__(self, num_nodes: int, memory_dim: int, message_dim: int = 0):
        """``MemoryBank`` (``models/MemoryModel.py:304-422``); ``node_memories`` and ``node_last_updated_times`` are
        non-grad Parameters so they land in the state_dict like the reference's.  Device state next to them
        (``_state``): the look-ahead view ``mem_view`` / ``lu_view`` (== ``get_updated_memories`` of every node), the
        ``pending`` flags and the last raw message of every node (``msg_store`` / ``msg_time``).  ``_view_key`` names the
        recurrent-cell weights the view was computed with (None: unknown, rebuild before use)."""
        super().__init__()
        self.num_nodes = num_nodes
        self.memory_dim = memory_dim
        self.message_dim = message_dim
        self.node_memories = nn.Parameter(torch.zeros((num_nodes, memory_dim)), requires_grad=False)
        self.node_last_updated_times = nn.Parameter(torch.zeros(num_nodes), requires_grad=False)
        self._state = None
        self._view_key = _ANY_WEIGHTS

    def _ensure(self):
        dev = self.node_memories.device
        if self._state is None or self._state['mem_view'].device != dev:
            N, D = self.num_nodes, self.memory_dim
            old = self._state
            self._state = dict(
                mem_view=self.node_memories.data.clone(),
                lu_view=self.node_last_updated_times.data.clone(),
                pending=torch.zeros(N, dtype=torch.uint8, device=dev),
                winner=torch.full((N,), -1, dtype=torch.int32, device=dev),
                msg_store=torch.zeros((N, max(self.message_dim, 1)), dtype=torch.float32, device=dev),
                msg_time=torch.zeros(N, dtype=torch.float64, device=dev),
                flag=torch.zeros(1, dtype=torch.int32, device=dev))
            if old is not None:      # the module moved to another device: the pending messages move with it
                for k in ('pending', 'msg_store', 'msg_time'):
                    self._state[k].copy_(old[k])
                self._view_key = None
        return self._state

    def mark_view_stale(self):
        self._view_key = None

    def _load_from_state_dict(self, *args, **kwargs):
        # loaded memories replace the persisted state: the look-ahead view has to be rebuilt from them
        super()._load_from_state_dict(*args, **kwargs)
        self._view_key = None

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
        self._view_key = _ANY_WEIGHTS

    def get_memories(self, node_ids):
        ids = _as_dev(node_ids, torch.int64, self.node_memories.device)
        return ops.gather_rows(self.node_memories.data, ids)

    def set_memories(self, node_ids, updated_node_memories: torch.Tensor):
        ids = _as_dev(node_ids, torch.int64, self.node_memories.device)
        self.node_memories.data[ids] = updated_node_memories.detach()
        self._view_key = None

    def get_node_last_updated_times(self, unique_node_ids):
        ids = _as_dev(unique_node_ids, torch.int64, self.node_memories.device)
        return self.node_last_updated_times.data[ids]

    # ---- raw messages: the reference's dict-of-lists attribute, backed by the device tables
    @property
    def node_raw_messages(self):
        """``{node_id: [(message, time)]}`` view of the pending messages (``models/MemoryModel.py:321-322``); assignable
        (``utils/EarlyStopping.py:86``) and item-assignable (``evaluate_link_prediction.py:152-156``)."""
        return RawMessageStore(self)

    @node_raw_messages.setter
    def node_raw_messages(self, messages):
        st = self._ensure()
        if isinstance(messages, RawMessageStore) and messages._bank is self:
            return
        st['pending'].zero_()
        self._view_key = None
        self._load_messages(messages.to_dict() if isinstance(messages, RawMessageStore) else messages)

    def _load_messages(self, messages):
        """Install ``{node: [(message, time), ...]}``: the last entry of every non-empty list becomes the node's pending message,
        an empty list clears it."""
        st = self._ensure()
        dev = st['pending'].device
        keep_ids, rows, times, drop = [], [], [], []
        for v, lst in messages.items():
            v = int(v)
            if not 0 <= v < self.num_nodes:
                raise KeyError(v)
            if len(lst) == 0:
                drop.append(v)
            else:
                keep_ids.append(v)
                rows.append(torch.as_tensor(lst[-1][0]).detach().to(device=dev, dtype=torch.float32).reshape(-1))
                times.append(float(lst[-1][1]))
        if drop:
            st['pending'][torch.tensor(drop, dtype=torch.int64, device=dev)] = 0
        if keep_ids:
            ids = torch.tensor(keep_ids, dtype=torch.int64, device=dev)
            block = torch.stack(rows)
            if block.shape[1] != st['msg_store'].shape[1]:
                raise ValueError(f'raw messages have {block.shape[1]} columns, the bank stores {st["msg_store"].shape[1]}')
            st['msg_store'][ids] = block
            st['msg_time'][ids] = torch.tensor(times, dtype=torch.float64, device=dev)
            st['pending'][ids] = 1
        if drop or keep_ids:
            self._view_key = None

    def store_node_raw_messages(self, node_ids, new_node_raw_messages: dict):
        """``store_node_raw_messages`` (``models/MemoryModel.py:389-398``): append the new messages of ``node_ids``."""
        self._load_messages({int(v): new_node_raw_messages[v] for v in np.asarray(node_ids).reshape(-1).tolist()
                             if len(new_node_raw_messages[v]) > 0})

    def clear_node_raw_messages(self, node_ids):
        """``clear_node_raw_messages`` (``models/MemoryModel.py:400-407``)."""
        st = self._ensure()
        ids = _as_dev(node_ids, torch.int64, self.node_memories.device)
        if ids.numel():
            st['pending'][ids] = 0
            self._view_key = None

    def backup_memory_bank(self):
        """``backup_memory_bank`` (``models/MemoryModel.py:351-360``): flat device copies (the third element carries the
        pending messages, the look-ahead view and the weights key that view belongs to)."""
        st = self._ensure()
        third = {k: v.clone() for k, v in st.items()}
        third['_view_key'] = self._view_key
        return (self.node_memories.data.clone(), self.node_last_updated_times.data.clone(), third)

    def reload_memory_bank(self, backup_memory_bank: tuple):
        """``reload_memory_bank`` (``models/MemoryModel.py:362-372``).  The restored look-ahead view is only trusted if it was
        built by the weights the model has now (``MemoryModel._refresh_view_if_stale`` compares the key)."""
        self.node_memories.data.copy_(backup_memory_bank[0])
        self.node_last_updated_times.data.copy_(backup_memory_bank[1])
        st = self._ensure()
        third = backup_memory_bank[2]
        if '_view_key' in third or 'mem_view' in third:
            for k, v in third.items():   # in place (graph-safe)
                if k != '_view_key':
                    st[k].copy_(v)
            self._view_key = third.get('_view_key')
        else:                            # a backup in the reference's format: {node: [(message, time)]}
            self.node_raw_messages = third

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

    def _operands(self, unique_node_ids, unique_node_messages, unique_node_timestamps):
        bank = self.memory_bank
        dev = bank.node_memories.device
        ids = _as_dev(unique_node_ids, torch.int64, dev)
        msg = unique_node_messages.detach().to(device=dev, dtype=torch.float32).contiguous()
        ts = _as_dev(unique_node_timestamps, torch.float64, dev).float()
        assert bool((bank.node_last_updated_times.data[ids] <= ts).all().item()), 'Trying to update memory to time in the past!'
        return ids, msg, ts

    def update_memories(self, unique_node_ids, unique_node_messages: torch.Tensor, unique_node_timestamps):
        """``update_memories`` (``models/MemoryModel.py:435-459``): persist cell(message, memory) and the message time for
        ``unique_node_ids`` (one fused launch, dyg_gru_update_fwd)."""
        if len(unique_node_ids) <= 0:
            return
        bank = self.memory_bank
        ids, msg, ts = self._operands(unique_node_ids, unique_node_messages, unique_node_timestamps)
        new = torch.empty((ids.numel(), bank.memory_dim), dtype=torch.float32, device=ids.device)
        gru_update(self.memory_updater, self.gates, msg, None, bank.node_memories.data, ids, new, None)
        bank.node_memories.data[ids] = new
        bank.node_last_updated_times.data[ids] = ts
        bank.mark_view_stale()

    def get_updated_memories(self, unique_node_ids, unique_node_messages: torch.Tensor, unique_node_timestamps):
        """``get_updated_memories`` (``models/MemoryModel.py:461-487``): copies of the whole memory / last-update tables with
        the rows of ``unique_node_ids`` advanced by their messages; nothing is persisted."""
        bank = self.memory_bank
        mem, lu = bank.node_memories.data.clone(), bank.node_last_updated_times.data.clone()
        if len(unique_node_ids) <= 0:
            return mem, lu
        ids, msg, ts = self._operands(unique_node_ids, unique_node_messages, unique_node_timestamps)
        gru_update(self.memory_updater, self.gates, msg, None, bank.node_memories.data, ids, mem, ids)
        lu[ids] = ts
        return mem, lu


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

    def compute_node_temporal_embeddings(self, node_memories: torch.Tensor, node_ids, node_time_intervals: torch.Tensor):
        """``compute_node_temporal_embeddings`` (``models/MemoryModel.py:534-545``):
        dropout(memories[ids] * (1 + linear(intervals))); eval mode runs dyg_jodie_project."""
        dev = node_memories.device
        ids = _as_dev(node_ids, torch.int64, dev)
        if self.training and torch.is_grad_enabled():
            return self.dropout(node_memories[ids] * (1 + self.linear_layer(node_time_intervals.unsqueeze(dim=1))))
        mem = node_memories.detach().float().contiguous()
        iv = node_time_intervals.detach().to(device=dev, dtype=torch.float64).contiguous()
        M, D = ids.numel(), self.memory_dim
        out = torch.empty((M, D), dtype=torch.float32, device=dev)
        zeros = torch.zeros(mem.shape[0], dtype=torch.float32, device=dev)   # "last update" 0: the interval is given directly
        _native.check(_native.load().dyg_jodie_project(_p(mem), int(mem.stride(0)), _p(zeros), _p(ids), _p(iv), M, D, 0.0, 1.0,
                                                       _p(self.linear_layer.weight.detach().reshape(-1)), _p(self.linear_layer.bias.detach()),
                                                       _p(out), D, _stream()))
        ops._count()
        return out


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
        # eval-mode batches of TGN with one attention layer run as ONE cooperative launch (dyg_tgn_step); False selects the
        # kernel-per-step path (the same results up to fp32 summation order; tests compare the two)
        self.fused_step = os.environ.get('DYG_TGN_FUSED_STEP', '1') != '0'
        self._step_scratch = {}
        self.to(device)

    def compute_src_dst_node_temporal_embeddings(self, src_node_ids: np.ndarray, dst_node_ids: np.ndarray, node_interact_times: np.ndarray,
                                                 edge_ids: np.ndarray, edges_are_positive: bool = True, num_neighbors: int = 20):
        """``compute_src_dst_node_temporal_embeddings`` (``models/MemoryModel.py:87-168``)."""
        if self.training and torch.is_grad_enabled():
            self.memory_bank.mark_view_stale()      # the weights move between training calls
            return self._forward_train(src_node_ids, dst_node_ids, node_interact_times, edge_ids, edges_are_positive, num_neighbors)
        self._refresh_view_if_stale()
        dev = self.node_raw_features.device
        src = _as_dev(src_node_ids, torch.int64, dev)
        dst = _as_dev(dst_node_ids, torch.int64, dev)
        tq = _as_dev(node_interact_times, torch.float64, dev)
        B = src.numel()
        if edges_are_positive and self._fused_step_ok(B):
            assert edge_ids is not None
            emb, _ = self._fused_step(None, src, dst, tq, _as_dev(edge_ids, torch.int64, dev), num_neighbors)
            return emb[:B], emb[B:]
        emb, ret = self._embed_eval([src, dst], tq, num_neighbors)
        if edges_are_positive:
            assert edge_ids is not None
            self._advance(src, dst, tq, _as_dev(edge_ids, torch.int64, dev), emb[B:], emb[:B], None)
        return ret[:B], ret[B:]

    def compute_pos_neg_temporal_embeddings(self, src_node_ids, dst_node_ids, neg_dst_node_ids, node_interact_times, edge_ids,
                                            num_neighbors: int = 20, link_predictor=None):
        """Addition to the reference API: the two calls its loops make per batch (``train_link_prediction.py:236-247``,
        ``evaluate_models_utils.py:60-80``) -- ``compute(src, neg, t, None, False)`` then ``compute(src, dst, t, edge_ids, True)``
        -- in one.  Both read the same memory state, so the 4 B roots go through ONE embedding pass (half the kernel launches
        of a latency-bound step), then the positive batch advances the memory.
        Returns (neg_src_emb, neg_dst_emb, pos_src_emb, pos_dst_emb), equal to the two calls' results.
        ``link_predictor`` (a ``MergeLayer``, ``model[1]`` of the reference's ``nn.Sequential``): additionally returns
        ``(pos_prob, neg_prob)`` = ``link_predictor(src_emb, dst_emb).squeeze(-1).sigmoid()`` of the two pairs
        (``train_link_prediction.py:243-244``); on the fused TGN path they come out of the same launch."""
        def with_probs(a, b, c, d):
            if link_predictor is None:
                return a, b, c, d
            lp = link_predictor
            pr = ops.mlp2([torch.cat([c, a]), torch.cat([d, b])], lp.fc1.weight.detach(), lp.fc1.bias.detach(), lp.fc2.weight.detach(),
                          lp.fc2.bias.detach(), act2=ops.ACT_SIGMOID).reshape(-1)
            n = c.shape[0]
            return a, b, c, d, pr[:n], pr[n:]
        if self.training and torch.is_grad_enabled():
            a, b = self.compute_src_dst_node_temporal_embeddings(src_node_ids, neg_dst_node_ids, node_interact_times, None, False, num_neighbors)
            c, d = self.compute_src_dst_node_temporal_embeddings(src_node_ids, dst_node_ids, node_interact_times, edge_ids, True, num_neighbors)
            return with_probs(a, b, c, d)
        self._refresh_view_if_stale()
        dev = self.node_raw_features.device
        src = _as_dev(src_node_ids, torch.int64, dev)
        dst = _as_dev(dst_node_ids, torch.int64, dev)
        neg = _as_dev(neg_dst_node_ids, torch.int64, dev)
        tq = _as_dev(node_interact_times, torch.float64, dev)
        B = src.numel()
        assert edge_ids is not None
        if self._fused_step_ok(B):
            # the src embedding of the negative pair equals that of the positive pair (same memory state, same times): 3 B roots
            emb, prob = self._fused_step(neg, src, dst, tq, _as_dev(edge_ids, torch.int64, dev), num_neighbors,
                                         link_predictor=link_predictor, pairs='pos_neg')
            out = (emb[:B], emb[B:2 * B], emb[:B], emb[2 * B:])
            return out if link_predictor is None else out + (prob[:B], prob[B:])
        emb, ret = self._embed_eval([src, neg, src, dst], tq, num_neighbors)
        self._advance(src, dst, tq, _as_dev(edge_ids, torch.int64, dev), emb[3 * B:], emb[2 * B:3 * B], None)
        return with_probs(ret[:B], ret[B:2 * B], ret[2 * B:3 * B], ret[3 * B:])

    # ------------------------------------------------------------------ one launch per batch (dyg_tgn_step)
    def _fused_step_ok(self, B):
        em = self.embedding_module
        return (self.fused_step and self.model_name == 'TGN' and self.num_layers == 1 and self.num_heads == 2 and B > 0 and
                em.neighbor_sampler.sample_neighbor_strategy == 'recent' and self.node_feat_dim % 4 == 0 and self.edge_feat_dim % 4 == 0 and
                self.time_feat_dim % 4 == 0 and self.node_feat_dim + self.edge_feat_dim <= 384 and self.time_feat_dim <= 128 and
                (self.node_feat_dim + self.time_feat_dim) % 8 == 0 and self.message_dim % 8 == 0 and
                (self.node_feat_dim + self.edge_feat_dim + self.time_feat_dim) % 4 == 0 and
                self.node_raw_features.stride(0) % 4 == 0 and self.edge_raw_features.stride(0) % 4 == 0)

    @staticmethod
    def _pack_planes(weight, widths):
        """BF16x3 operand planes (2, N, sum of padded widths) of ``weight[:, :sum(widths)]`` with every K segment zero-padded to a
        multiple of 8 columns (dyg_tgn_step's weight layout)."""
        w = weight.detach().float()
        N = w.shape[0]
        pads = [(x + 7) // 8 * 8 for x in widths]
        packed = torch.zeros((N, sum(pads)), dtype=torch.float32, device=w.device)
        src = dst = 0
        for x, xp in zip(widths, pads):
            packed[:, dst:dst + x] = w[:, src:src + x]
            src += x
            dst += xp
        hi = packed.to(torch.bfloat16)
        mid = (packed - hi.float()).to(torch.bfloat16)
        return torch.stack([hi, mid]).contiguous()

    def _step_weights(self, link_predictor):
        """Weight-derived operands of dyg_tgn_step, rebuilt when a parameter changes."""
        em = self.embedding_module
        attn, merge, cell = em.temporal_conv_layers[0], em.merge_layers[0], self.memory_updater.memory_updater
        params = [attn.query_projection.weight, attn.key_projection.weight, attn.value_projection.weight, attn.residual_fc.weight,
                  merge.fc1.weight, merge.fc2.weight, cell.weight_ih, cell.weight_hh, self.time_encoder.w.weight, self.time_encoder.w.bias]
        if link_predictor is not None:
            params += [link_predictor.fc1.weight, link_predictor.fc1.bias, merge.fc2.bias]
        key = (ops.WEIGHTS_EPOCH,) + tuple((q.data_ptr(), q._version) for q in params)
        ent = getattr(self, '_step_w', None)
        if ent is None or ent[0] != key:
            F_, T_ = self.node_feat_dim, self.time_feat_dim
            Dq = F_ + T_
            dev = self.node_raw_features.device
            t0 = zero_time_features(self.time_encoder, dev)
            wqk, wvr = attn.folded()
            w = dict(t0=t0, cq=query_constant(attn, self.time_encoder, t0),
                     wqk=self._pack_planes(wqk, [F_]), wvr=self._pack_planes(wvr, [wvr.shape[1]]),
                     m1=self._pack_planes(merge.fc1.weight, [Dq, F_]), m2=self._pack_planes(merge.fc2.weight, [F_]),
                     w_ih=self._pack_planes(cell.weight_ih, [self.message_dim]), w_hh=self._pack_planes(cell.weight_hh, [F_]))
            if link_predictor is not None:
                # fc1([emb_a | emb_b]) with emb = W2 h + b2 folded onto h (float64 once per weight version, like MultiHeadAttention.folded)
                wp, bp = link_predictor.fc1.weight.detach().double(), link_predictor.fc1.bias.detach().double()
                w2, b2 = merge.fc2.weight.detach().double(), merge.fc2.bias.detach().double()
                wa, wb = wp[:, :F_], wp[:, F_:2 * F_]
                w['p1'] = self._pack_planes(torch.cat([wa @ w2, wb @ w2], dim=1).float(), [F_, F_])
                w['p1_b'] = (bp + wa @ b2 + wb @ b2).float().contiguous()
            ent = self._step_w = (key, w, params)
        return ent[1]

    def _fused_step(self, neg, src, dst, tq, eid, k, link_predictor=None, pairs=None):
        """Embeddings of the roots ``[src | neg | dst]`` (``[src | dst]`` when ``neg`` is None; root r at time tq[r % B]) on the
        look-ahead view, then the memory update of the positive batch (src, dst, tq, eid), then optionally the link probabilities
        -- one cooperative launch (csrc/tgn_step.cu), which reads the three id lists where they lie."""
        lib = _native.load()
        dev = self.node_raw_features.device
        bank, em, sampler = self.memory_bank, self.embedding_module, self.embedding_module.neighbor_sampler
        st = bank._ensure()
        attn, merge = em.temporal_conv_layers[0], em.merge_layers[0]
        cell = self.memory_updater.memory_updater
        F_, E_, T_ = self.node_feat_dim, self.edge_feat_dim, self.time_feat_dim
        H = self.num_heads
        Dk, Dq, MD, Fp = F_ + E_ + T_, F_ + T_, self.message_dim, (F_ + 7) // 8 * 8
        B = src.numel()
        R = (3 if neg is not None else 2) * B
        P_ = 2 * B if link_predictor is not None else 0
        key = (R, B, P_, int(k), str(dev))
        sc = self._step_scratch.get(key)
        if sc is None:
            f32 = dict(dtype=torch.float32, device=dev)

            def pl(rows, cols):       # operand planes written by the kernel; the padding columns stay zero
                return torch.zeros((2, rows, cols), dtype=torch.bfloat16, device=dev)
            sc = dict(nbr_ids=torch.empty((R, k), dtype=torch.int64, device=dev), nbr_eids=torch.empty((R, k), dtype=torch.int64, device=dev),
                      nbr_t=torch.empty((R, k), **f32), feat=torch.empty((R, F_), **f32), qk=torch.empty((R, H * Dk), **f32),
                      o=torch.empty((R, Dq), **f32), msg=torch.empty((2 * B, MD), **f32), hnew=torch.empty((2 * B, F_), **f32),
                      ph=torch.empty((max(P_, 1), F_), **f32), barrier=torch.zeros(2, dtype=torch.int32, device=dev),
                      feat_pl=pl(R, Fp), msg_pl=pl(2 * B, MD), s_pl=pl(R, H * Dk), y_pl=pl(R, Dq), h1_pl=pl(R, Fp))
            if P_:
                ar = torch.arange(B, dtype=torch.int64, device=dev)
                sc['pair_a'] = torch.cat([ar, ar])                       # pos pairs (src, dst) then neg pairs (src, neg)
                sc['pair_b'] = torch.cat([ar + 2 * B, ar + B])
            if len(self._step_scratch) > 8:
                self._step_scratch.clear()
            self._step_scratch[key] = sc
        emb = torch.empty((R, F_), dtype=torch.float32, device=dev)
        prob = torch.empty(max(P_, 1), dtype=torch.float32, device=dev)
        wts = self._step_weights(link_predictor)
        w, b = self.time_encoder.wb()
        p = _native.TgnStep()
        keep = []

        def ptr(x):
            keep.append(x)
            return _p(x).value if x is not None else None

        def planes(field, x):
            keep.append(x)
            field.hi, field.mid, field.ld = x[0].data_ptr(), x[1].data_ptr(), x.shape[2]
        p.he, p.indptr, p.num_nodes = ptr(sampler.halfedges), ptr(sampler.indptr), sampler.num_nodes
        p.src, p.dst, p.t, p.eid, p.neg, p.roots = ptr(src), ptr(dst), ptr(tq), ptr(eid), ptr(neg), None
        p.B, p.R, p.k, p.H, p.G, p.check_time = B, R, int(k), H, self.memory_updater.gates, int(self.check_time_order)
        p.node_raw, p.ld_node = ptr(self.node_raw_features), self.node_raw_features.stride(0)
        p.edge_raw, p.ld_edge = ptr(self.edge_raw_features), self.edge_raw_features.stride(0)
        p.F, p.E, p.T = F_, E_, T_
        p.memory, p.last_update = ptr(bank.node_memories.data), ptr(bank.node_last_updated_times.data)
        p.mem_view, p.lu_view, p.pending, p.winner = ptr(st['mem_view']), ptr(st['lu_view']), ptr(st['pending']), ptr(st['winner'])
        p.msg_store, p.msg_time, p.flag = ptr(st['msg_store']), ptr(st['msg_time']), ptr(st['flag'])
        p.time_w, p.time_b, p.t0 = ptr(w), ptr(b), ptr(wts['t0'])
        for name in ('wqk', 'wvr', 'm1', 'm2', 'w_ih', 'w_hh'):
            planes(getattr(p, name), wts[name])
        p.cq, p.rbias = ptr(wts['cq']), ptr(attn.residual_fc.bias.detach())
        p.ln_g, p.ln_b, p.ln_eps = ptr(attn.layer_norm.weight.detach()), ptr(attn.layer_norm.bias.detach()), float(attn.layer_norm.eps)
        p.m1_b, p.m2_b = ptr(merge.fc1.bias.detach()), ptr(merge.fc2.bias.detach())
        p.b_ih, p.b_hh = ptr(cell.bias_ih.detach()), ptr(cell.bias_hh.detach())
        if P_:
            lp = link_predictor
            planes(p.p1, wts['p1'])
            p.p1_b = ptr(wts['p1_b'])
            p.p2_w, p.p2_b = ptr(lp.fc2.weight.detach().reshape(-1)), ptr(lp.fc2.bias.detach())
            p.pair_a, p.pair_b, p.P = ptr(sc['pair_a']), ptr(sc['pair_b']), P_
        for name in ('nbr_ids', 'nbr_eids', 'nbr_t', 'feat', 'qk', 'o', 'msg', 'hnew', 'ph', 'barrier'):
            setattr(p, name, ptr(sc[name]))
        for name in ('feat_pl', 'msg_pl', 's_pl', 'y_pl', 'h1_pl'):
            planes(getattr(p, name), sc[name])
        p.emb, p.prob = ptr(emb), ptr(prob)
        if getattr(self, 'phase_ns', None) is not None:      # profiling hook: per-phase globaltimer stamps (scripts/tgn_phases.py)
            p.phase_ns = ptr(self.phase_ns)
        flops = 2.0 * (R * (F_ * H * Dk + H * Dk * Dq + (Dq + F_) * F_ + F_ * F_) + 2 * B * self.memory_updater.gates * F_ * (MD + F_) +
                       P_ * 2 * F_ * F_)
        with ops._Timed('tgn_step_kernel', flops, 4.0 * (R * k * (2 * F_ + E_) + 2 * B * (2 * MD + 4 * F_))):
            _native.check(lib.dyg_tgn_step(ctypes.byref(p), _stream()))
        ops._count()
        return emb, prob

    def _cell_key(self):
        """Identity of the recurrent cell's weights (the look-ahead view of the pending nodes is a function of them)."""
        return (ops.WEIGHTS_EPOCH,) + tuple((q.data_ptr(), q._version) for q in self.memory_updater.memory_updater.parameters())

    def _refresh_view_if_stale(self):
        """Make ``mem_view`` / ``lu_view`` equal ``get_updated_memories`` of all nodes under the CURRENT weights
        (``models/MemoryModel.py:108-109``).  The view is maintained incrementally by ``_advance``; it is rebuilt from the
        persisted memories and the stored last messages when the bank says it is unknown (after ``load_state_dict``, a
        raw-message assignment, a compat mutator or a training step) or was built by other weights (``reload_memory_bank``
        of a backup taken before the weights moved)."""
        bank = self.memory_bank
        key = self._cell_key()
        if bank._view_key == _ANY_WEIGHTS:
            bank._view_key = key
        elif bank._view_key != key:
            self._rebuild_view()
            bank._view_key = key

    def _rebuild_view(self):
        bank = self.memory_bank
        st = bank._ensure()
        with torch.no_grad():
            st['mem_view'].copy_(bank.node_memories.data)
            st['lu_view'].copy_(bank.node_last_updated_times.data)
            pend = torch.nonzero(st['pending']).reshape(-1)
            if pend.numel():
                gru_update(self.memory_updater.memory_updater, self.memory_updater.gates, st['msg_store'], pend,
                           bank.node_memories.data, pend, st['mem_view'], pend)
                st['lu_view'][pend] = st['msg_time'][pend].float()

    # ------------------------------------------------------------------ the reference's sub-API (compat surface)
    def get_updated_memories(self, node_ids, node_raw_messages):
        """``get_updated_memories`` (``models/MemoryModel.py:170-191``): (memories, last-update times) of ALL nodes with the
        rows of ``node_ids`` advanced by their last pending message; nothing is persisted.  With the bank's own message store
        this is a masked copy of the look-ahead view; with any other ``{node: [(message, time)]}`` dict the messages are
        aggregated on the host and run through the fused cell."""
        bank = self.memory_bank
        if isinstance(node_raw_messages, RawMessageStore) and node_raw_messages._bank is bank:
            self._refresh_view_if_stale()
            st = bank._ensure()
            dev = bank.node_memories.device
            sel = torch.zeros(self.num_nodes, dtype=torch.bool, device=dev)
            sel[_as_dev(node_ids, torch.int64, dev)] = True
            sel &= st['pending'].bool()
            return (torch.where(sel.unsqueeze(1), st['mem_view'], bank.node_memories.data),
                    torch.where(sel, st['lu_view'], bank.node_last_updated_times.data))
        ids, msgs, ts = self.message_aggregator.aggregate_messages(node_ids=node_ids, node_raw_messages=node_raw_messages)
        return self.memory_updater.get_updated_memories(unique_node_ids=ids, unique_node_messages=msgs, unique_node_timestamps=ts)

    def update_memories(self, node_ids, node_raw_messages):
        """``update_memories`` (``models/MemoryModel.py:193-210``): persist the update of ``node_ids`` from their last messages."""
        ids, msgs, ts = self.message_aggregator.aggregate_messages(node_ids=node_ids, node_raw_messages=node_raw_messages)
        self.memory_updater.update_memories(unique_node_ids=ids, unique_node_messages=msgs, unique_node_timestamps=ts)

    def compute_new_node_raw_messages(self, src_node_ids, dst_node_ids, dst_node_embeddings, node_interact_times, edge_ids):
        """``compute_new_node_raw_messages`` (``models/MemoryModel.py:212-251``): one message per event for the nodes in
        ``src_node_ids``: [memory[src] | memory[dst] (DyRep: dst embedding) | time_enc(t - last_update[src]) | edge feature]
        (dyg_tgn_build_messages), returned as (unique node ids, {node: [(message row, time), ...]})."""
        dev = self.node_raw_features.device
        bank = self.memory_bank
        src = _as_dev(src_node_ids, torch.int64, dev)
        dst = _as_dev(dst_node_ids, torch.int64, dev)
        tq = _as_dev(node_interact_times, torch.float64, dev)
        eid = _as_dev(edge_ids, torch.int64, dev)
        B = src.numel()
        D, T, E = self.memory_dim, self.time_feat_dim, self.edge_feat_dim
        other = None
        if self.model_name == 'DyRep':       # the kernel builds both roles; only the src-role half is asked for here
            other = torch.zeros((2 * B, D), dtype=torch.float32, device=dev)
            other[:B] = dst_node_embeddings.detach()
        w, b = self.time_encoder.wb()
        msg = torch.empty((2 * B, self.message_dim), dtype=torch.float32, device=dev)
        _native.check(_native.load().dyg_tgn_build_messages(
            _p(src), _p(dst), _p(tq), _p(eid), B, _p(bank.node_memories.data), _p(bank.node_last_updated_times.data), D, _p(other), D,
            _p(self.edge_raw_features), self.edge_raw_features.stride(0), E, _p(w), _p(b), T, _p(msg), self.message_dim, _stream()))
        ops._count()
        src_h = np.asarray(src_node_ids) if not isinstance(src_node_ids, torch.Tensor) else src_node_ids.cpu().numpy()
        t_h = np.asarray(node_interact_times) if not isinstance(node_interact_times, torch.Tensor) else node_interact_times.cpu().numpy()
        new_node_raw_messages = defaultdict(list)
        for i in range(B):
            new_node_raw_messages[src_h[i]].append((msg[i], t_h[i]))
        return np.unique(src_h), new_node_raw_messages

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
        # the recurrent cell of the winning candidates in one launch (gate GEMMs + gates + scatter into the look-ahead view),
        # then the bookkeeping of the same candidates (message store, times, pending flags)
        G = self.memory_updater.gates
        gru_update(self.memory_updater.memory_updater, G, msg, None, mem, node_ids, mem_view, node_ids, winner=st['winner'])
        _native.check(lib.dyg_tgn_cell_commit(None, None, G, _p(src), _p(dst), _p(tq), B, _p(st['winner']), _p(mem),
                                              _p(mem_view), _p(lu_view), _p(st['pending']), D, _p(msg), self.message_dim,
                                              self.message_dim, _p(st['msg_store']), _p(st['msg_time']), _stream()))
        ops._count(2)

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
