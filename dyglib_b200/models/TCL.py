"""Drop-in for the reference's ``TCL`` (``models/TCL.py``): a caller of the hot path.

Per root it builds the sequence [self, k neighbours] from the device sampler (``get_historical_neighbors``), encodes it with
node rows, edge rows, the trainable ``TimeEncoder`` and a depth embedding, and runs self- and cross-attention transformer
blocks with shared weights between the source and destination sequences.  Sampling, the time encoding (forward
``dyg_time_encode``, backward ``dyg_time_encode_bwd``) and every dense layer (``autograd.linear`` on the sm_100a GEMMs) come
from the path; the 21-position masked softmax, LayerNorm, ReLU and dropout are library elementwise ops.
Same constructor arguments, parameter names and ``state_dict`` as the reference.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import autograd as ag
from ..utils.utils import NeighborSampler, _as_dev
from .modules import TimeEncoder, TransformerEncoder


class TCL(nn.Module):

    def __init__(self, node_raw_features: np.ndarray, edge_raw_features: np.ndarray, neighbor_sampler: NeighborSampler,
                 time_feat_dim: int, num_layers: int = 2, num_heads: int = 2, num_depths: int = 20, dropout: float = 0.1,
                 device: str = 'cuda'):
        """Same arguments as ``TCL.__init__`` (``models/TCL.py:11-54``)."""
        super().__init__()
        self.node_raw_features = torch.from_numpy(node_raw_features.astype(np.float32)).to(device).contiguous()
        self.edge_raw_features = torch.from_numpy(edge_raw_features.astype(np.float32)).to(device).contiguous()
        self.neighbor_sampler = neighbor_sampler
        self.node_feat_dim = self.node_raw_features.shape[1]
        self.edge_feat_dim = self.edge_raw_features.shape[1]
        self.time_feat_dim = time_feat_dim
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.num_depths = num_depths
        self.dropout = dropout
        self.device = device
        self.time_encoder = TimeEncoder(time_dim=time_feat_dim)
        self.depth_embedding = nn.Embedding(num_embeddings=num_depths, embedding_dim=self.node_feat_dim)
        self.projection_layer = nn.ModuleDict({
            'node': nn.Linear(self.node_feat_dim, self.node_feat_dim, bias=True),
            'edge': nn.Linear(self.edge_feat_dim, self.node_feat_dim, bias=True),
            'time': nn.Linear(self.time_feat_dim, self.node_feat_dim, bias=True)})
        self.transformers = nn.ModuleList([TransformerEncoder(self.node_feat_dim, self.num_heads, self.dropout) for _ in range(self.num_layers)])
        self.output_layer = nn.Linear(self.node_feat_dim, self.node_feat_dim, bias=True)
        self.to(device)

    def _sequences(self, ids, tq, k):
        """[self | k sampled neighbours]: ids, edge ids (0 for self) and times (the query time for self) (``models/TCL.py:84-99``)."""
        nbr, eid, nt = self.neighbor_sampler.get_historical_neighbors_device(ids, tq, k)
        n = ids.numel()
        seq_ids = torch.cat([ids.reshape(n, 1), nbr], dim=1)
        seq_eids = torch.cat([torch.zeros((n, 1), dtype=torch.int64, device=ids.device), eid], dim=1)
        # the reference concatenates float64 query times with the float32 neighbour times (-> float64) before the subtraction
        seq_t = torch.cat([tq.reshape(n, 1), nt.double()], dim=1)
        return seq_ids, seq_eids, seq_t

    def get_features(self, node_interact_times, nodes_neighbor_ids, nodes_edge_ids, nodes_neighbor_times, time_encoder: TimeEncoder = None):
        """``get_features`` (``models/TCL.py:156-183``) on device tensors (numpy inputs are accepted)."""
        dev = self.node_raw_features.device
        ids = _as_dev(nodes_neighbor_ids, torch.int64, dev)
        eids = _as_dev(nodes_edge_ids, torch.int64, dev)
        tq = _as_dev(node_interact_times, torch.float64, dev)
        nt = _as_dev(nodes_neighbor_times, torch.float64, dev)
        te = time_encoder or self.time_encoder
        dt = (tq.reshape(-1, 1) - nt).float()
        time_feat = ag.time_encode(dt, te.w.weight, te.w.bias)
        assert ids.shape[1] == self.depth_embedding.weight.shape[0]
        return self.node_raw_features[ids], self.edge_raw_features[eids], time_feat, self.depth_embedding.weight

    def compute_src_dst_node_temporal_embeddings(self, src_node_ids: np.ndarray, dst_node_ids: np.ndarray, node_interact_times: np.ndarray,
                                                 num_neighbors: int = 20):
        """``compute_src_dst_node_temporal_embeddings`` (``models/TCL.py:56-154``): src neighbours are sampled first, then dst."""
        dev = self.node_raw_features.device
        src = _as_dev(src_node_ids, torch.int64, dev)
        dst = _as_dev(dst_node_ids, torch.int64, dev)
        tq = _as_dev(node_interact_times, torch.float64, dev)
        k = int(num_neighbors)
        s_ids, s_eids, s_t = self._sequences(src, tq, k)
        d_ids, d_eids, d_t = self._sequences(dst, tq, k)
        pl = self.projection_layer
        feats = []
        for ids, eids, t in ((s_ids, s_eids, s_t), (d_ids, d_eids, d_t)):
            nf, ef, tf, depth = self.get_features(tq, ids, eids, t, self.time_encoder)
            n, L, F_ = nf.shape
            x = (ag.linear(nf.reshape(n * L, F_), pl['node'].weight, pl['node'].bias) +
                 ag.linear(ef.reshape(n * L, -1), pl['edge'].weight, pl['edge'].bias) +
                 ag.linear(tf.reshape(n * L, -1), pl['time'].weight, pl['time'].bias)).reshape(n, L, F_)
            feats.append(x + depth)
        s, d = feats
        for tr in self.transformers:
            s = tr(inputs_query=s, inputs_key=s, inputs_value=s, neighbor_masks=s_ids)
            d = tr(inputs_query=d, inputs_key=d, inputs_value=d, neighbor_masks=d_ids)
            s_new = tr(inputs_query=s, inputs_key=d, inputs_value=d, neighbor_masks=d_ids)
            d_new = tr(inputs_query=d, inputs_key=s, inputs_value=s, neighbor_masks=s_ids)
            s, d = s_new, d_new
        return (ag.linear(s[:, 0, :].contiguous(), self.output_layer.weight, self.output_layer.bias),
                ag.linear(d[:, 0, :].contiguous(), self.output_layer.weight, self.output_layer.bias))

    def set_neighbor_sampler(self, neighbor_sampler: NeighborSampler):
        """``set_neighbor_sampler`` (``models/TCL.py:185-194``)."""
        self.neighbor_sampler = neighbor_sampler
        if self.neighbor_sampler.sample_neighbor_strategy in ['uniform', 'time_interval_aware']:
            assert self.neighbor_sampler.seed is not None
            self.neighbor_sampler.reset_random_state()
