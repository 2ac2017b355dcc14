"""Drop-in for the reference's ``TGAT`` (``models/TGAT.py``) on the device-resident sampler and fused kernels."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..utils.utils import NeighborSampler, _as_dev
from .modules import TimeEncoder, MergeLayer, MultiHeadAttention
from ._temporal import temporal_conv, temporal_conv_train, zero_time_features


class TGAT(nn.Module):

    def __init__(self, node_raw_features: np.ndarray, edge_raw_features: np.ndarray, neighbor_sampler: NeighborSampler,
                 time_feat_dim: int, num_layers: int = 2, num_heads: int = 2, dropout: float = 0.1, device: str = 'cuda'):
        """Same arguments as ``TGAT.__init__`` (``models/TGAT.py:11-46``); feature tables live in HBM."""
        super().__init__()
        self.node_raw_features = torch.from_numpy(node_raw_features.astype(np.float32)).to(device).contiguous()
        self.edge_raw_features = torch.from_numpy(edge_raw_features.astype(np.float32)).to(device).contiguous()
        self._zero_row0 = ops.zero_row0_flags(self.node_raw_features, self.edge_raw_features)
        self.neighbor_sampler = neighbor_sampler
        self.node_feat_dim = self.node_raw_features.shape[1]
        self.edge_feat_dim = self.edge_raw_features.shape[1]
        self.time_feat_dim = time_feat_dim
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.dropout = dropout
        self.time_encoder = TimeEncoder(time_dim=time_feat_dim)
        self.temporal_conv_layers = nn.ModuleList([
            MultiHeadAttention(self.node_feat_dim, self.edge_feat_dim, self.time_feat_dim, self.num_heads, self.dropout)
            for _ in range(num_layers)])
        self.merge_layers = nn.ModuleList([
            MergeLayer(self.node_feat_dim + self.time_feat_dim, self.node_feat_dim, self.node_feat_dim, self.node_feat_dim)
            for _ in range(num_layers)])
        self.to(device)

    def compute_src_dst_node_temporal_embeddings(self, src_node_ids: np.ndarray, dst_node_ids: np.ndarray,
                                                 node_interact_times: np.ndarray, num_neighbors: int = 20):
        """``compute_src_dst_node_temporal_embeddings`` (``models/TGAT.py:48-64``): src recursion first, then dst
        (the order matters for the random strategies' RNG stream)."""
        src = self.compute_node_temporal_embeddings(src_node_ids, node_interact_times, self.num_layers, num_neighbors)
        dst = self.compute_node_temporal_embeddings(dst_node_ids, node_interact_times, self.num_layers, num_neighbors)
        return src, dst

    def compute_node_temporal_embeddings(self, node_ids, node_interact_times, current_layer_num: int, num_neighbors: int = 20):
        """``compute_node_temporal_embeddings`` (``models/TGAT.py:66-136``); accepts numpy arrays or CUDA tensors."""
        assert current_layer_num >= 0
        dev = self.node_raw_features.device
        ids = _as_dev(node_ids, torch.int64, dev)
        tq = _as_dev(node_interact_times, torch.float64, dev)
        if self.training and torch.is_grad_enabled():
            return self._embed_train(ids, tq, current_layer_num, num_neighbors)
        t0 = zero_time_features(self.time_encoder, dev)
        return self._embed(ids, tq, current_layer_num, num_neighbors, t0)

    def _embed_train(self, ids, tq, layer, k):
        """The recursion of ``_embed`` with autograd (training mode)."""
        raw = ops.gather_rows(self.node_raw_features, ids)
        if layer == 0:
            return raw
        conv = raw if layer == 1 else self._embed_train(ids, tq, layer - 1, k)
        nbr, eid, nt = self.neighbor_sampler.get_historical_neighbors_device(ids, tq, k)
        nbr_dense = self._embed_train(nbr.reshape(-1), nt.reshape(-1).double(), layer - 1, k) if layer > 1 else None
        return temporal_conv_train(self.temporal_conv_layers[layer - 1], self.merge_layers[layer - 1], self.time_encoder, conv, raw,
                                   self.node_raw_features, nbr, nbr_dense, self.edge_raw_features, eid, tq, nt, k,
                                   zero_row0=self._zero_row0)

    def _embed(self, ids, tq, layer, k, t0):
        raw = ops.gather_rows(self.node_raw_features, ids)
        if layer == 0:
            return raw
        conv = raw if layer == 1 else self._embed(ids, tq, layer - 1, k, t0)
        nbr, eid, nt = self.neighbor_sampler.get_historical_neighbors_device(ids, tq, k)
        nbr_dense = None
        if layer > 1:
            # hop >= 2 queries run at the float32-rounded neighbour times (models/TGAT.py:107-110)
            nbr_dense = self._embed(nbr.reshape(-1), nt.reshape(-1).double(), layer - 1, k, t0)
        return temporal_conv(self.temporal_conv_layers[layer - 1], self.merge_layers[layer - 1], self.time_encoder, t0,
                             conv, raw, self.node_raw_features, None, nbr, nbr_dense, self.edge_raw_features, eid, tq, nt, k,
                             zero_row0=self._zero_row0)

    def set_neighbor_sampler(self, neighbor_sampler: NeighborSampler):
        """``set_neighbor_sampler`` (``models/TGAT.py:138-147``)."""
        self.neighbor_sampler = neighbor_sampler
        if self.neighbor_sampler.sample_neighbor_strategy in ['uniform', 'time_interval_aware']:
            assert self.neighbor_sampler.seed is not None
            self.neighbor_sampler.reset_random_state()
