"""Helpers that rebuild reference inputs from the committed golden fixtures (tests/golden/*.npz)."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(prefix=""):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def cora_dense_features(z):
    x = np.zeros((int(z["in_n"]), int(z["in_feat_dim"])), dtype=np.float32)
    x[z["in_feat_row"], z["in_feat_col"]] = z["in_feat_val"]
    return x


def dense_normalized_with_self_loops(z, symmetric):
    """homophily_tests.py:83-85 -- normalize_tensor(eye + A.to_dense(), symmetric).to_sparse(), as COO."""
    n = int(z["in_n"])
    ei = z["in_edge_index"].astype(np.int64)
    a = torch.zeros(n, n)
    a[ei[0], ei[1]] = 1.0
    a = a + torch.eye(n)
    rowsum = a.sum(1)
    if symmetric == 0:
        r = rowsum.pow(-1)
        r[torch.isinf(r)] = 0
        a = r[:, None] * a
    else:
        r = rowsum.pow(-0.5)
        r[torch.isinf(r)] = 0
        a = (r[:, None] * a) * r[None, :]
    s = a.to_sparse().coalesce()
    return s.indices()[0].numpy(), s.indices()[1].numpy(), s.values().numpy()


def proj_matrix(d, seed=5):
    """The column subset + random projection make_golden.spmm_projection() used."""
    g = torch.Generator().manual_seed(seed)
    cols = torch.randperm(d, generator=g)[: min(24, d)].sort().values
    proj = torch.randn(d, 16, generator=g)
    return cols.numpy(), proj.double().numpy()


def plot_flow_adjacency(z):
    """synthetic_plot.py:94 -- normalize(adj + eye): dense row-normalised adjacency with self-loops (float32)."""
    n = int(z["in_n"])
    ei = z["in_edge_index"].astype(np.int64)
    a = torch.zeros(n, n)
    a[ei[0], ei[1]] = 1.0
    a = a + torch.eye(n)
    r = 1.0 / a.sum(1)
    r[torch.isinf(r)] = 0
    return r[:, None] * a


def linkx_graph(z):
    """Symmetric facebook100 adjacency of a linkx_* fixture as coalesced (row, col), rebuilt from the stored triangle."""
    up = z["in_edges_upper"].astype(np.int64)
    row, col = np.concatenate([up[0], up[1]]), np.concatenate([up[1], up[0]])
    order = np.lexsort((col, row))
    return row[order], col[order]


def l1_normalize(x):
    """torch.nn.functional.normalize(x, p=1, dim=1) as homophily_tests.py:95 calls it (eps = 1e-12)."""
    x = torch.as_tensor(x, dtype=torch.float32)
    return x / x.abs().sum(1, keepdim=True).clamp_min(1e-12)
