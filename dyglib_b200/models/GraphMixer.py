"""Drop-in for the reference's ``GraphMixer`` (``models/GraphMixer.py``): a caller of the hot path.

It consumes exactly the pieces the path provides -- ``recent`` neighbour sampling (k link-encoder neighbours and a
``time_gap``-deep window for the node encoder), the (here fixed) ``TimeEncoder``, gathers of edge / node rows -- and adds an
MLP-Mixer over the k tokens.  Sampling runs on the device CSR; the link encoder's input projection is the fused gather-GEMM
(edge rows and time encodings are never materialised in eval mode); every dense layer runs forward on the sm_100a GEMMs
through ``autograd.linear`` (which also gives the training path); LayerNorm / GELU / dropout over the (n, k, 172) mixer
activations are library elementwise ops.  Same constructor arguments, parameter names and ``state_dict`` as the reference.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import autograd as ag
from .. import ops
from ..utils.utils import NeighborSampler, _as_dev
from .modules import TimeEncoder


class FeedForwardNet(nn.Module):

    def __init__(self, input_dim: int, dim_expansion_factor: float, dropout: float = 0.0):
        """``FeedForwardNet`` (``models/GraphMixer.py:164-190``): Linear, GELU, Dropout, Linear, Dropout."""
        super().__init__()
        self.input_dim = input_dim
        self.dim_expansion_factor = dim_expansion_factor
        self.dropout = dropout
        hidden = int(dim_expansion_factor * input_dim)
        self.ffn = nn.Sequential(nn.Linear(input_dim, hidden), nn.GELU(), nn.Dropout(dropout), nn.Linear(hidden, input_dim), nn.Dropout(dropout))

    def forward(self, x: torch.Tensor):
        shape = x.shape
        h = ag.linear(x.reshape(-1, shape[-1]), self.ffn[0].weight, self.ffn[0].bias)
        h = F.dropout(F.gelu(h), self.dropout, self.training)
        h = ag.linear(h, self.ffn[3].weight, self.ffn[3].bias)
        return F.dropout(h, self.dropout, self.training).reshape(shape)


class MLPMixer(nn.Module):

    def __init__(self, num_tokens: int, num_channels: int, token_dim_expansion_factor: float = 0.5,
                 channel_dim_expansion_factor: float = 4.0, dropout: float = 0.0):
        """``MLPMixer`` (``models/GraphMixer.py:193-238``)."""
        super().__init__()
        self.token_norm = nn.LayerNorm(num_tokens)
        self.token_feedforward = FeedForwardNet(num_tokens, token_dim_expansion_factor, dropout)
        self.channel_norm = nn.LayerNorm(num_channels)
        self.channel_feedforward = FeedForwardNet(num_channels, channel_dim_expansion_factor, dropout)

    def forward(self, input_tensor: torch.Tensor):
        """Token mixing over the k neighbours, then channel mixing, both with residuals (``:218-238``)."""
        hidden = self.token_norm(input_tensor.permute(0, 2, 1))
        hidden = self.token_feedforward(hidden.contiguous()).permute(0, 2, 1)
        output = hidden + input_tensor
        hidden = self.channel_feedforward(self.channel_norm(output))
        return hidden + output


class GraphMixer(nn.Module):

    def __init__(self, node_raw_features: np.ndarray, edge_raw_features: np.ndarray, neighbor_sampler: NeighborSampler,
                 time_feat_dim: int, num_tokens: int, num_layers: int = 2, token_dim_expansion_factor: float = 0.5,
                 channel_dim_expansion_factor: float = 4.0, dropout: float = 0.1, device: str = 'cuda'):
        """Same arguments as ``GraphMixer.__init__`` (``models/GraphMixer.py:11-56``)."""
        super().__init__()
        self.node_raw_features = torch.from_numpy(node_raw_features.astype(np.float32)).to(device).contiguous()
        self.edge_raw_features = torch.from_numpy(edge_raw_features.astype(np.float32)).to(device).contiguous()
        self.neighbor_sampler = neighbor_sampler
        self.node_feat_dim = self.node_raw_features.shape[1]
        self.edge_feat_dim = self.edge_raw_features.shape[1]
        self.time_feat_dim = time_feat_dim
        self.num_tokens = num_tokens
        self.num_layers = num_layers
        self.token_dim_expansion_factor = token_dim_expansion_factor
        self.channel_dim_expansion_factor = channel_dim_expansion_factor
        self.dropout = dropout
        self.device = device
        self.num_channels = self.edge_feat_dim
        self.time_encoder = TimeEncoder(time_dim=time_feat_dim, parameter_requires_grad=False)   # not trainable in GraphMixer (:42-43)
        self.projection_layer = nn.Linear(self.edge_feat_dim + time_feat_dim, self.num_channels)
        self.mlp_mixers = nn.ModuleList([
            MLPMixer(self.num_tokens, self.num_channels, self.token_dim_expansion_factor, self.channel_dim_expansion_factor, self.dropout)
            for _ in range(self.num_layers)])
        self.output_layer = nn.Linear(self.num_channels + self.node_feat_dim, self.node_feat_dim, bias=True)
        self.to(device)

    def compute_src_dst_node_temporal_embeddings(self, src_node_ids: np.ndarray, dst_node_ids: np.ndarray, node_interact_times: np.ndarray,
                                                 num_neighbors: int = 20, time_gap: int = 2000):
        """``compute_src_dst_node_temporal_embeddings`` (``models/GraphMixer.py:58-76``): src first, then dst."""
        src = self.compute_node_temporal_embeddings(src_node_ids, node_interact_times, num_neighbors, time_gap)
        dst = self.compute_node_temporal_embeddings(dst_node_ids, node_interact_times, num_neighbors, time_gap)
        return src, dst

    def compute_node_temporal_embeddings(self, node_ids, node_interact_times, num_neighbors: int = 20, time_gap: int = 2000):
        """``compute_node_temporal_embeddings`` (``models/GraphMixer.py:78-151``)."""
        dev = self.node_raw_features.device
        ids = _as_dev(node_ids, torch.int64, dev)
        tq = _as_dev(node_interact_times, torch.float64, dev)
        n, k = ids.numel(), int(num_neighbors)
        E, T = self.edge_feat_dim, self.time_feat_dim
        # ---- link encoder: k most recent interactions, [edge row | masked time encoding] -> channels -> MLP-Mixer -> mean
        nbr, eid, nt = self.neighbor_sampler.get_historical_neighbors_device(ids, tq, k)
        w, b = self.time_encoder.wb()
        pl = self.projection_layer
        if torch.is_grad_enabled() and self.training:
            dt = (tq.reshape(n, 1) - nt.double()).float()
            te = ops.time_encode(dt.reshape(-1), w, b).reshape(n, k, T) * (nbr != 0).unsqueeze(-1)
            x = torch.cat([self.edge_raw_features[eid], te], dim=-1).reshape(n * k, E + T)
            x = ag.linear(x, pl.weight, pl.bias)
        else:
            x = ops.linear([ops.seg_rows(self.edge_raw_features, E, eid.reshape(-1)),
                            ops.seg_time(nt.reshape(-1), w, b, mask_ids=nbr.reshape(-1), t_query=tq, tq_div=k)],
                           n * k, pl.weight.detach(), pl.bias.detach())
        x = x.reshape(n, k, self.num_channels)
        for mixer in self.mlp_mixers:
            x = mixer(x)
        link = torch.mean(x, dim=1)
        # ---- node encoder: mean over the `time_gap` most recent neighbours' raw rows, weighted by softmax of the validity mask
        # and divided by time_gap once more, exactly as written in the reference (:127-142)
        gap_ids, _, _ = self.neighbor_sampler.get_historical_neighbors_device(ids, tq, int(time_gap))
        mask = (gap_ids > 0).float()
        mask[mask == 0] = -1e10
        scores = torch.softmax(mask, dim=1)
        agg = torch.einsum('nj,njf->nf', scores, self.node_raw_features[gap_ids]) / float(time_gap)
        node = agg + self.node_raw_features[ids]
        return ag.linear([link, node], self.output_layer.weight, self.output_layer.bias)

    def set_neighbor_sampler(self, neighbor_sampler: NeighborSampler):
        """``set_neighbor_sampler`` (``models/GraphMixer.py:153-161``)."""
        self.neighbor_sampler = neighbor_sampler
        if self.neighbor_sampler.sample_neighbor_strategy in ['uniform', 'time_interval_aware']:
            assert self.neighbor_sampler.seed is not None
            self.neighbor_sampler.reset_random_state()
