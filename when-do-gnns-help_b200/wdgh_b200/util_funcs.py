"""Drop-in mirror of the graph-statistics parts of the reference's `utils/util_funcs.py`.

Normalisers run on the GPU through libwdgh_b200.so; the split / accuracy helpers are the
reference's host-side RNG logic (they must consume torch's global RNG identically so that
seeded runs reproduce).  Dataset loaders (load_data, full_load_data*, utils/datasets.py) are
out of scope of the hot path -- keep using the reference's.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch

from . import graph as G
from .graph import CSRGraph


def _to_graph(adj, binary=False) -> CSRGraph:
    if isinstance(adj, CSRGraph):
        return adj
    if isinstance(adj, torch.Tensor):
        return CSRGraph.from_torch_sparse(adj if adj.is_sparse else adj.to_sparse(), binary=binary)
    return CSRGraph.from_scipy(sp.coo_matrix(adj), binary=binary)


def normalize_tensor(mx, symmetric=0):
    """Row-normalise (symmetric=0) or D^-1/2 M D^-1/2 (symmetric=1) a dense matrix.  util_funcs.py:365-380."""
    return G.normalize_dense(mx, 1 if symmetric else 0)


def sys_normalized_adjacency(adj):
    """D^-1/2 (A + I) D^-1/2 as a resident CSRGraph.  util_funcs.py:418-426."""
    return _to_graph(adj).with_self_loops().normalized(G.NORM_SYM)


def row_normalized_adjacency(adj):
    """D^-1 (A + I) (l1 row normalisation) as a resident CSRGraph.  util_funcs.py:383-390."""
    return _to_graph(adj).with_self_loops().normalized(G.NORM_RW)


def sparse_mx_to_torch_sparse_tensor(sparse_mx):
    """util_funcs.py:400-407; also accepts the CSRGraph the normalisers above return."""
    if isinstance(sparse_mx, CSRGraph):
        return sparse_mx.to_torch_sparse()
    sparse_mx = sparse_mx.tocoo().astype(np.float32)
    indices = torch.from_numpy(np.vstack((sparse_mx.row, sparse_mx.col)).astype(np.int64))
    values = torch.from_numpy(sparse_mx.data)
    return torch.sparse_coo_tensor(indices, values, torch.Size(sparse_mx.shape))


def propagate(adj, features, symmetric=1, add_self_loop=True):
    """SGC-1 / GCN propagation A_hat X without materialising A_hat (on-the-fly D^-1/2 or D^-1)."""
    g = _to_graph(adj, binary=True)
    return G.spmm(g, features, G.NORM_SYM if symmetric else G.NORM_RW, add_self_loop)


def accuracy(labels, output):
    """util_funcs.py:393-397."""
    preds = output.max(1)[1].type_as(labels)
    correct = preds.eq(labels).double()
    return correct.sum() / len(labels)


def index_to_mask(index, size):
    """util_funcs.py:478-481."""
    mask = torch.zeros(size, dtype=torch.bool, device=index.device)
    mask[index] = 1
    return mask


def random_disassortative_splits(labels, num_classes, training_percentage=0.6):
    """util_funcs.py:454-475: class-balanced 60/20/20 masks; same torch.randperm call sequence."""
    labels = torch.as_tensor(labels).cpu()
    num_classes = int(num_classes)
    indices = []
    for i in range(num_classes):
        index = torch.nonzero((labels == i)).view(-1)
        index = index[torch.randperm(index.size(0))]
        indices.append(index)
    percls_trn = int(round(training_percentage * (labels.size()[0] / num_classes)))
    val_lb = int(round(0.2 * labels.size()[0]))
    train_index = torch.cat([i[:percls_trn] for i in indices], dim=0)
    rest_index = torch.cat([i[percls_trn:] for i in indices], dim=0)
    rest_index = rest_index[torch.randperm(rest_index.size(0))]
    train_mask = index_to_mask(train_index, size=labels.size()[0])
    val_mask = index_to_mask(rest_index[:val_lb], size=labels.size()[0])
    test_mask = index_to_mask(rest_index[val_lb:], size=labels.size()[0])
    return train_mask, val_mask, test_mask


def rand_train_test_idx(label, train_prop=.6, valid_prop=.2, ignore_negative=True):
    """util_funcs.py:484-508."""
    labeled_nodes = torch.where(label != -1)[0] if ignore_negative else label
    n = labeled_nodes.shape[0]
    train_num, valid_num = int(n * train_prop), int(n * valid_prop)
    perm = torch.as_tensor(np.random.permutation(n))
    train_indices, val_indices, test_indices = perm[:train_num], perm[train_num:train_num + valid_num], perm[train_num + valid_num:]
    if not ignore_negative:
        return train_indices, val_indices, test_indices
    return labeled_nodes[train_indices], labeled_nodes[val_indices], labeled_nodes[test_indices]
