"""On-disk formats and index loaders around the hot path: drop-in for ``utils/DataLoader.py``.

``get_link_prediction_data`` reads the reference's processed files (``processed_data/<name>/ml_<name>.csv`` with columns
``u, i, ts, label, idx``; ``ml_<name>.npy`` edge features, ``ml_<name>_node.npy`` node features) and returns the same eight
objects (``utils/DataLoader.py:66-176``).  The train / val / test / new-node splits are reproduced with the reference's
``random.seed(2020)`` draw; the reference passes a ``set`` to ``random.sample`` (``:115``), which Python >= 3.11 rejects -- the
draw here is over ``tuple(the_set)``, exactly what Python <= 3.10 did with a set population.  Masks are vectorised
(``np.isin``) instead of per-row python lambdas.  Host plumbing only: nothing here touches the GPU.
"""
from __future__ import annotations

import os
import random

import numpy as np
from torch.utils.data import DataLoader, Dataset

NODE_FEAT_DIM = EDGE_FEAT_DIM = 172


class CustomizedDataset(Dataset):
    """``CustomizedDataset`` (``utils/DataLoader.py:7-26``): a dataset of indices."""

    def __init__(self, indices_list: list):
        super().__init__()
        self.indices_list = indices_list

    def __getitem__(self, idx: int):
        return self.indices_list[idx]

    def __len__(self):
        return len(self.indices_list)


def get_idx_data_loader(indices_list: list, batch_size: int, shuffle: bool):
    """``get_idx_data_loader`` (``utils/DataLoader.py:29-43``)."""
    return DataLoader(dataset=CustomizedDataset(indices_list=indices_list), batch_size=batch_size, shuffle=shuffle, drop_last=False)


class Data:
    """``Data`` (``utils/DataLoader.py:46-64``)."""

    def __init__(self, src_node_ids: np.ndarray, dst_node_ids: np.ndarray, node_interact_times: np.ndarray, edge_ids: np.ndarray,
                 labels: np.ndarray):
        self.src_node_ids = src_node_ids
        self.dst_node_ids = dst_node_ids
        self.node_interact_times = node_interact_times
        self.edge_ids = edge_ids
        self.labels = labels
        self.num_interactions = len(src_node_ids)
        self.unique_node_ids = set(src_node_ids) | set(dst_node_ids)
        self.num_unique_nodes = len(self.unique_node_ids)


def _pad_features(x, dim, what, dataset_name):
    assert dim >= x.shape[1], f'{what} feature dimension in dataset {dataset_name} is bigger than {dim}!'
    if x.shape[1] < dim:
        x = np.concatenate([x, np.zeros((x.shape[0], dim - x.shape[1]))], axis=1)
    return x


def _load(dataset_name, root):
    import pandas as pd
    d = os.path.join(root, dataset_name)
    graph_df = pd.read_csv(os.path.join(d, f'ml_{dataset_name}.csv'))
    edge_raw_features = _pad_features(np.load(os.path.join(d, f'ml_{dataset_name}.npy')), EDGE_FEAT_DIM, 'Edge', dataset_name)
    node_raw_features = _pad_features(np.load(os.path.join(d, f'ml_{dataset_name}_node.npy')), NODE_FEAT_DIM, 'Node', dataset_name)
    return graph_df, node_raw_features, edge_raw_features


def _subset(full: Data, mask):
    return Data(full.src_node_ids[mask], full.dst_node_ids[mask], full.node_interact_times[mask], full.edge_ids[mask], full.labels[mask])


def get_link_prediction_data(dataset_name: str, val_ratio: float, test_ratio: float, root: str = './processed_data', verbose: bool = True):
    """``get_link_prediction_data`` (``utils/DataLoader.py:66-176``): (node_raw_features, edge_raw_features, full_data, train_data,
    val_data, test_data, new_node_val_data, new_node_test_data)."""
    graph_df, node_raw_features, edge_raw_features = _load(dataset_name, root)
    val_time, test_time = list(np.quantile(graph_df.ts, [(1 - val_ratio - test_ratio), (1 - test_ratio)]))
    src = graph_df.u.values.astype(np.longlong)
    dst = graph_df.i.values.astype(np.longlong)
    t = graph_df.ts.values.astype(np.float64)
    full_data = Data(src, dst, t, graph_df.idx.values.astype(np.longlong), graph_df.label.values)

    random.seed(2020)                                            # "the setting of seed follows previous works" (:104)
    node_set = set(src) | set(dst)
    test_node_set = set(src[t > val_time]).union(set(dst[t > val_time]))
    # 10 % of all nodes, drawn among the nodes seen after the validation time, are held out as new nodes (:112-115)
    new_test_nodes = np.array(random.sample(tuple(test_node_set), int(0.1 * len(node_set))), dtype=np.longlong)
    observed = ~np.isin(src, new_test_nodes) & ~np.isin(dst, new_test_nodes)
    train_data = _subset(full_data, (t <= val_time) & observed)
    train_nodes = np.union1d(train_data.src_node_ids, train_data.dst_node_ids)
    assert not np.isin(train_nodes, new_test_nodes).any()
    new_nodes = np.setdiff1d(np.fromiter(node_set, dtype=np.longlong, count=len(node_set)), train_nodes)   # never seen in training
    val_mask = (t <= test_time) & (t > val_time)
    test_mask = t > test_time
    has_new = np.isin(src, new_nodes) | np.isin(dst, new_nodes)
    val_data, test_data = _subset(full_data, val_mask), _subset(full_data, test_mask)
    new_node_val_data, new_node_test_data = _subset(full_data, val_mask & has_new), _subset(full_data, test_mask & has_new)
    if verbose:
        for name, d in (('dataset', full_data), ('training dataset', train_data), ('validation dataset', val_data), ('test dataset', test_data),
                        ('new node validation dataset', new_node_val_data), ('new node test dataset', new_node_test_data)):
            print(f'The {name} has {d.num_interactions} interactions, involving {d.num_unique_nodes} different nodes')
        print(f'{len(new_test_nodes)} nodes were used for the inductive testing, i.e. are never seen during training')
    return node_raw_features, edge_raw_features, full_data, train_data, val_data, test_data, new_node_val_data, new_node_test_data


def get_node_classification_data(dataset_name: str, val_ratio: float, test_ratio: float, root: str = './processed_data'):
    """``get_node_classification_data`` (``utils/DataLoader.py:179-229``): chronological split, no held-out nodes."""
    graph_df, node_raw_features, edge_raw_features = _load(dataset_name, root)
    val_time, test_time = list(np.quantile(graph_df.ts, [(1 - val_ratio - test_ratio), (1 - test_ratio)]))
    t = graph_df.ts.values.astype(np.float64)
    full_data = Data(graph_df.u.values.astype(np.longlong), graph_df.i.values.astype(np.longlong), t,
                     graph_df.idx.values.astype(np.longlong), graph_df.label.values)
    return (node_raw_features, edge_raw_features, full_data, _subset(full_data, t <= val_time),
            _subset(full_data, (t <= test_time) & (t > val_time)), _subset(full_data, t > test_time))
