"""CPU oracle for the graph-statistics hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a CPU restatement (numpy + torch-CPU + scipy) of the reference's
algorithms for the path named in BASELINE.json `north_star`:

    /root/reference/utils/homophily_metrics.py   (metrics)
    /root/reference/utils/util_funcs.py          (normalisers, splits, accuracy)

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it; the product package
(`when-do-gnns-help_b200/wdgh_b200`) never does and fails loudly when its CUDA
library is missing.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function
here against outputs of the UNMODIFIED reference, captured by
`tests/golden/make_golden.py` (Cora, four `data_synthesis` graphs, eight
edge-case graphs) and committed under `tests/golden/*.npz`.

Conventions: a graph is the coalesced COO triple `(row, col, val)` of a torch
sparse tensor -- entries sorted row-major, duplicates already summed -- plus
the node count `n`; that is exactly what `A.coalesce().indices()/.values()`
hands the reference.  Every function cites the reference lines it restates.
Each float result follows the reference's precision flow (float64 scipy
normalisers cast to float32, float32 torch arithmetic afterwards); each
integer quantity (degrees, match counts, class-pair histograms) is returned
exactly so the CUDA path can be compared bit-for-bit.
"""
from __future__ import annotations

import math
import random

import numpy as np
import scipy.sparse as sp
import torch

__all__ = [
    "coalesce", "csr_from_coo", "sys_normalized_adjacency", "row_normalized_adjacency",
    "normalize_tensor", "normalize", "preprocess_features", "normalize_adj",
    "dataset_edge_balance", "spmm", "structure_counts", "edge_homophily", "node_homophily",
    "compat_matrix", "class_homophily", "class_distribution", "adjusted_homo",
    "label_informativeness", "edge_cosine", "generalized_edge_homophily", "similarity",
    "gntk_kernels", "random_disassortative_splits", "accuracy", "kr_metric",
]


# ---------------------------------------------------------------------------
# graph containers
# ---------------------------------------------------------------------------
def coalesce(row, col, val, n):
    """torch `sparse_coo_tensor(...).coalesce()`: row-major sort + duplicate sum.

    Mirrors what every metric does first (homophily_metrics.py:50,63,127,165).
    """
    row = np.asarray(row, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    val = np.ones(row.shape[0], np.float32) if val is None else np.asarray(val, dtype=np.float32)
    key = row * np.int64(n) + col
    order = np.argsort(key, kind="stable")
    key, val = key[order], val[order]
    uniq, start = np.unique(key, return_index=True)
    summed = np.add.reduceat(val, start).astype(np.float32) if key.size else val
    return (uniq // n).astype(np.int64), (uniq % n).astype(np.int64), summed


def csr_from_coo(row, n):
    """Row pointer (int64, n+1) of a row-major sorted COO."""
    counts = np.bincount(np.asarray(row, dtype=np.int64), minlength=n)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    return rowptr


# ---------------------------------------------------------------------------
# util_funcs.py normalisers
# ---------------------------------------------------------------------------
def _with_self_loops(row, col, val, n):
    a = sp.coo_matrix((np.asarray(val, dtype=np.float64), (row, col)), shape=(n, n))
    return (a + sp.eye(n)).tocsr()


def sys_normalized_adjacency(row, col, val, n):
    """D^-1/2 (A+I) D^-1/2 in float64, stored float32.  util_funcs.py:418-426, 400-407."""
    a = _with_self_loops(row, col, val, n)
    deg = np.asarray(a.sum(1)).ravel()
    deg = np.where(deg == 0, 1.0, deg)            # :422  (row_sum == 0) * 1 + row_sum
    with np.errstate(divide="ignore"):
        dis = np.power(deg, -0.5)                 # :423
    dis[np.isinf(dis)] = 0.0                      # :424
    m = sp.diags(dis).dot(a).dot(sp.diags(dis)).tocoo()   # :425-426
    return coalesce(m.row, m.col, m.data.astype(np.float32), n)   # sparse_mx_to_torch_sparse_tensor :402


def row_normalized_adjacency(row, col, val, n):
    """D^-1 (A+I) (sklearn l1 row normalisation) in float64, stored float32.  util_funcs.py:383-390."""
    a = _with_self_loops(row, col, val, n).tocsr()
    norms = np.asarray(abs(a).sum(1)).ravel()     # sk_normalize(norm='l1') divides by sum |a_ij|
    norms[norms == 0] = 1.0
    m = sp.diags(1.0 / norms).dot(a).tocoo()
    return coalesce(m.row, m.col, m.data.astype(np.float32), n)


def normalize_tensor(mx, symmetric=0):
    """Dense torch row / symmetric normalisation.  util_funcs.py:365-380."""
    mx = torch.as_tensor(mx)
    rowsum = mx.sum(1)
    if symmetric == 0:
        r = rowsum.pow(-1).flatten()
        r[torch.isinf(r)] = 0.0
        return r[:, None] * mx                    # diag(r) @ mx
    r = rowsum.pow(-0.5).flatten()
    r[torch.isinf(r)] = 0.0
    return (r[:, None] * mx) * r[None, :]         # diag(r) @ mx @ diag(r)


def normalize(mx):
    """Row-normalise a matrix: diag(1 / rowsum) @ mx, rows with a zero sum stay zero.  util_funcs.py:29-36.
    (`preprocess_features`, util_funcs.py:39-46, is the same arithmetic applied to the feature matrix.)

    Dense input (numpy / torch, the synthetic_plot.py:82,92 flow) -> dense numpy array of the input's dtype;
    scipy sparse input (the full_load_data flow) -> scipy CSR in the dtype scipy promotes to."""
    if sp.issparse(mx):
        rowsum = np.asarray(mx.sum(1))
        with np.errstate(divide="ignore"):
            r_inv = (1 / rowsum).flatten()
        r_inv[np.isinf(r_inv)] = 0.0
        return sp.diags(r_inv).dot(mx)
    if isinstance(mx, torch.Tensor):               # the row sum is torch's (its float32 summation order), :31
        m, rowsum = mx.detach().cpu().numpy(), mx.detach().cpu().sum(1).numpy()
    else:
        m = np.asarray(mx)
        rowsum = m.sum(1)
    with np.errstate(divide="ignore"):
        r_inv = (1 / rowsum).flatten()
    r_inv[np.isinf(r_inv)] = 0.0
    return sp.diags(r_inv).dot(m)


preprocess_features = normalize


def normalize_adj(row, col, val, n):
    """(A D^-1/2)^T D^-1/2 = D^-1/2 A^T D^-1/2 with D = diag(row sums of A), float64.  util_funcs.py:429-436.
    Returns the coalesced COO triple of the result (values float64)."""
    a = sp.coo_matrix((np.asarray(val, dtype=np.float64), (row, col)), shape=(n, n))
    rowsum = np.asarray(a.sum(1))
    with np.errstate(divide="ignore"):
        dis = np.power(rowsum, -0.5).flatten()
    dis[np.isinf(dis)] = 0.0
    d = sp.diags(dis)
    m = a.dot(d).transpose().dot(d).tocoo()
    key = m.row.astype(np.int64) * n + m.col
    order = np.argsort(key, kind="stable")
    return m.row[order].astype(np.int64), m.col[order].astype(np.int64), m.data[order]


def dataset_edge_balance(row, col, val, labels, n):
    """Per class: node count, adjacency mass inside the class, adjacency mass leaving it.  util_funcs.py:439-451."""
    labels = np.asarray(labels)
    c = int(labels.max()) + 1
    a = sp.coo_matrix((np.asarray(val, dtype=np.float64), (row, col)), shape=(n, n)).tocsr()
    nodes = np.zeros(c)
    balance = np.zeros((c, 2))
    for i in range(c):
        idx = np.where(labels == i)[0]
        rest = np.delete(np.arange(n), idx)
        nodes[i] = idx.shape[0]
        balance[i, 0] = a[idx, :][:, idx].sum()
        balance[i, 1] = a[idx, :][:, rest].sum()
    return nodes, balance


def spmm(row, col, val, n, x, threads=None):
    """`torch.spmm(adj, features)` on the CPU in float32 (homophily_metrics.py:192,199,234)."""
    x = torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32)
    idx = torch.from_numpy(np.vstack([row, col]).astype(np.int64))
    a = torch.sparse_coo_tensor(idx, torch.as_tensor(np.asarray(val, dtype=np.float32)), (n, n))
    a = a.coalesce()
    return torch.sparse.mm(a, x).numpy()


# ---------------------------------------------------------------------------
# exact integer statistics every label metric is built from
# ---------------------------------------------------------------------------
def structure_counts(row, col, labels, n, num_classes=None):
    """All integer statistics of one (graph, labels) pair.

    Returns a dict with
      deg_all   [n]   entries per row INCLUDING self-loops      (hm.py:129 unique counts)
      deg_nsl   [n]   entries per row EXCLUDING self-loops      (hm.py:75 bincount)
      match_nsl [n]   per-row entries with equal endpoint labels, self-loops excluded (hm.py:76-77)
      match_all       number of stored entries with equal endpoint labels (hm.py:51)
      match_lab / n_lab   same, restricted to entries whose endpoints are both labelled >=0 (hm.py:52-54)
      hist      [C,C] class-pair counts over non-self-loop entries with both labels >= 0 (hm.py:91-100)
      hist_any  [C,C] class-pair counts over non-self-loop entries, labels taken as-is   (hm.py:144)
      class_count [C] nodes per class (labels >= 0)
    """
    row = np.asarray(row, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    lab = np.asarray(labels, dtype=np.int64).reshape(-1)
    c = int(lab.max()) + 1 if num_classes is None else int(num_classes)
    ls, lt = lab[row], lab[col]
    nsl = row != col
    same = ls == lt
    both = (ls >= 0) & (lt >= 0)
    out = {
        "deg_all": np.bincount(row, minlength=n).astype(np.int64),
        "deg_nsl": np.bincount(row[nsl], minlength=n).astype(np.int64),
        "match_nsl": np.bincount(row[nsl & same], minlength=n).astype(np.int64),
        "match_all": int(same.sum()),
        "match_lab": int((same & both).sum()),
        "n_lab": int(both.sum()),
        "nnz": int(row.shape[0]),
        "n_selfloop": int((~nsl).sum()),
    }
    k = nsl & both
    out["hist"] = np.bincount(ls[k] * c + lt[k], minlength=c * c).reshape(c, c).astype(np.int64)
    # class_distribution never masks negative labels: `labels[src] == i` simply never matches them
    out["hist_any"] = out["hist"].copy()
    out["class_count"] = np.bincount(lab[lab >= 0], minlength=c).astype(np.int64)
    return out


# ---------------------------------------------------------------------------
# homophily_metrics.py : label metrics
# ---------------------------------------------------------------------------
def edge_homophily(row, col, labels, ignore_negative=False):
    """homophily_metrics.py:43-57.  `labels` may be 1-D ints or a 2-D (one-hot) matrix.

    With a 2-D label matrix the reference compares label ROWS elementwise and
    averages over all nnz*C booleans (that is what homophily_tests.py:115-116
    feeds it), so a mismatching one-hot pair still scores (C-2)/C.
    """
    lab = np.asarray(labels)
    row = np.asarray(row, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    matching = lab[row] == lab[col]
    if ignore_negative:
        mask = (lab[row] >= 0) * (lab[col] >= 0)          # :52
        return float(np.mean(matching[mask]))             # :54
    return np.float32(np.mean(matching.astype(np.float32), dtype=np.float64))   # :56


def node_homophily(row, col, labels, n):
    """homophily_metrics.py:60-78 (self-loops removed, mean over nodes with a neighbour).

    Raises RuntimeError exactly where the reference does: `bincount` yields
    max(src)+1 bins, and `hs.scatter_add(...) / degs` needs that to equal n.
    """
    s = structure_counts(row, col, labels, n)
    rows_nsl = np.asarray(row)[np.asarray(row) != np.asarray(col)]
    nbins = int(rows_nsl.max()) + 1 if rows_nsl.size else 0
    if nbins != n and nbins != 1:
        raise RuntimeError(f"The size of tensor a ({n}) must match the size of tensor b ({nbins})")
    deg = s["deg_nsl"].astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        hs = s["match_nsl"].astype(np.float32) / deg      # :77
    keep = deg != 0
    return np.float32(np.mean(hs[keep], dtype=np.float64)) if keep.any() else np.float32("nan")


def compat_matrix(row, col, labels):
    """Row-normalised class compatibility matrix H.  homophily_metrics.py:81-102."""
    lab = np.asarray(labels, dtype=np.int64).reshape(-1)
    s = structure_counts(row, col, lab, lab.shape[0])
    h = s["hist"].astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return h / h.sum(1, keepdims=True)                # :101


def class_homophily(row, col, labels):
    """`our_measure`, the class-insensitive edge homophily.  homophily_metrics.py:105-123.

    `proportions` is indexed by class id although it is built from
    `unique(return_counts=True)`, i.e. only from classes that occur -- kept.
    """
    lab = np.asarray(labels, dtype=np.int64).reshape(-1)
    c = int(lab.max()) + 1
    h = compat_matrix(row, col, lab)
    nz = lab[lab >= 0]
    counts = np.unique(nz, return_counts=True)[1]
    prop = counts.astype(np.float32) / np.float32(nz.shape[0])
    val = np.float32(0)
    for k in range(c):
        add = np.float32(max(h[k, k] - prop[k], 0)) if not np.isnan(h[k, k]) else np.float32("nan")
        if not np.isnan(add):
            val = np.float32(val + add)
    return np.float32(val / np.float32(c - 1))


def class_distribution(row, col, labels, n):
    """(p, p_bar, pc) of homophily_metrics.py:126-147.

    deg := (entries per present row) - 1, i.e. the code assumes one self-loop
    per node; rows without any entry shift `deg` against the node ids and the
    reference then dies with IndexError -- raised here too.
    """
    lab = np.asarray(labels, dtype=np.int64).reshape(-1)
    c = int(lab.max()) + 1
    s = structure_counts(row, col, lab, n)
    if (s["deg_all"] == 0).any():
        raise IndexError("class_distribution: a node without any stored entry misaligns `deg` (hm.py:129,141)")
    deg = s["deg_all"] - 1                                 # :132
    p = np.unique(lab, return_counts=True)[1] / lab.shape[0]   # :137 (float64 true division -> float32 in torch)
    p_bar = np.zeros(c, dtype=np.float64)
    for i in range(c):
        p_bar[i] = deg[lab == i].sum()                     # :141
    pc = s["hist_any"].astype(np.float64)                  # :144
    tot = np.float32(deg.sum())
    p_bar = (p_bar.astype(np.float32) / tot).astype(np.float32)   # :145 (float32 tensors)
    pc = (pc.astype(np.float32) / tot).astype(np.float32)
    p_bar[p_bar == 0] = 1e-8                               # :146
    pc[pc == 0] = 1e-8
    return p.astype(np.float32), p_bar, pc


def adjusted_homo(row, col, labels, n):
    """homophily_metrics.py:150-155."""
    _, p_bar, _ = class_distribution(row, col, labels, n)
    eh = edge_homophily(row, col, labels)
    s2 = np.float32(np.sum(p_bar.astype(np.float32) ** 2, dtype=np.float32))
    return np.float32((eh - s2) / (np.float32(1) - s2))


def label_informativeness(row, col, labels, n):
    """homophily_metrics.py:158-161."""
    _, p_bar, pc = class_distribution(row, col, labels, n)
    num = np.sum(pc * np.log(pc), dtype=np.float32)
    den = np.sum(p_bar * np.log(p_bar), dtype=np.float32)
    return np.float32(2 - num / den)


# ---------------------------------------------------------------------------
# generalised edge homophily (feature cosine similarity over edges)
# ---------------------------------------------------------------------------
def edge_cosine(src, dst, x):
    """cos(x[src], x[dst]) per edge in float32; 0/0 -> NaN -> 0 (homophily_metrics.py:182-185)."""
    x = torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32)
    s = torch.as_tensor(np.asarray(src, dtype=np.int64))
    t = torch.as_tensor(np.asarray(dst, dtype=np.int64))
    sim = (x[s] * x[t]).sum(1) / (x[s].norm(dim=1, p=2) * x[t].norm(dim=1, p=2))
    sim[torch.isnan(sim)] = 0
    return sim.numpy()


def generalized_edge_homophily(row, col, x, n, sample_max=75000, iteration=10):
    """homophily_metrics.py:164-187.

    nnz <  sample_max: sum_{stored i!=j} cos(x_i,x_j) / #{stored i!=j}  (dense n x n in the
                       reference; evaluated edge-wise here -- same sum, hm.py:167-172).
                       The stored value must be > 0 to count (:171).
    nnz >= sample_max: `iteration` draws of `sample_max` stored entries (self-loops
                       included) with python's `random.sample`, mean cosine each (:178-187).
    """
    row = np.asarray(row, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    nnz = row.shape[0]
    if nnz < sample_max:
        keep = row != col
        sim = edge_cosine(row[keep], col[keep], x).astype(np.float64)
        return np.float32(sim.sum() / keep.sum())
    g = np.zeros(iteration)
    for i in range(iteration):
        pick = np.asarray(random.sample(range(nnz), int(sample_max)), dtype=np.int64)   # :179-180
        g[i] = np.float32(np.mean(edge_cosine(row[pick], col[pick], x), dtype=np.float64))
    return float(np.mean(g))


# ---------------------------------------------------------------------------
# aggregation homophily (similarity) and the GNTK / KR kernels
# ---------------------------------------------------------------------------
def similarity(features, row, col, val, n, label_onehot, hard=None, LP=1, ifsum=1, idx_train=None):
    """Aggregation similarity score.  homophily_metrics.py:190-229.

    features: [n,d] float32; label_onehot: [n,C]; idx_train: optional bool mask [n].
    """
    z = torch.from_numpy(spmm(row, col, val, n, features))
    label = torch.as_tensor(np.asarray(label_onehot), dtype=torch.float32)
    if idx_train is None:
        labels = label.argmax(1)
    else:
        m = torch.as_tensor(np.asarray(idx_train)).bool()
        labels = label.argmax(1)[m]
        label = label[m]
        z = z[m]
    gram = z @ z.T                                                    # :192 / :199
    c = int(labels.max()) + 1
    w = torch.zeros(z.shape[0], c)
    for i in range(c):
        cols = gram[:, labels == i]
        w[:, i] = cols.sum(1) if ifsum == 1 else cols.mean(1)         # :203-206
    own = w[torch.arange(labels.shape[0]), labels]
    if hard is None:
        if ifsum == 1:
            nnodes = labels.shape[0]
            degs_label = (label @ label.T).sum(1)                     # :210
        else:
            nnodes = c
            degs_label = 1
        if LP == 1:
            ratio = (own / degs_label) / ((w.sum(1) - own) / (nnodes - degs_label))   # :216-218
            ratio[torch.isnan(ratio)] = 0
            return np.float32((ratio >= 1).float().mean().item())
        return np.float32((((w - w * label).sum(1) <= 0) & ((w * label).sum(1) >= 0)).float().mean().item())
    if LP == 1:
        return np.float32(w.argmax(1).eq(labels).float().mean().item())               # :226
    return np.float32((((w - w * label).max(1)[0] <= 0.0) & ((w * label).sum(1) >= 0)).float().mean().item())



def similarity_near_ties(features, row, col, val, n, label_onehot, rel=1e-6):
    """Checker helper (not in the reference): number of nodes whose soft-LAS decision of `similarity` / `plot_similarity`
    (hard=None, LP=1, ifsum=1: ratio >= 1, hm.py:216-218 / hp.py:226-228) is a TIE in exact arithmetic -- |ratio - 1|
    <= rel with everything evaluated in float64.  For such a node the float32 outcome is decided by the summation order
    of the Gram and of the class sums (torch's CPU sgemm / vectorised sum vs any other order), so a comparison of two
    float32 implementations may differ by up to this many indicator flips; every other node must agree."""
    x = torch.as_tensor(np.ascontiguousarray(features), dtype=torch.float64)
    idx = torch.from_numpy(np.vstack([row, col]).astype(np.int64))
    a = torch.sparse_coo_tensor(idx, torch.as_tensor(np.asarray(val, dtype=np.float64)), (n, n)).coalesce().to_dense()
    z = a @ x
    inner = z @ z.T
    lab = torch.as_tensor(np.asarray(label_onehot)).argmax(1)
    c = int(lab.max()) + 1
    w = torch.stack([inner[:, lab == i].sum(1) for i in range(c)], 1)
    own = w[torch.arange(n), lab]
    degs = torch.bincount(lab, minlength=c)[lab].double()
    ratio = (own / degs) / ((w.sum(1) - own) / (n - degs))
    return int(((ratio - 1).abs() <= rel).sum().item())


def _arccos_kernel(gram, n_layers, eps=1e-8):
    """One half of gntk_homophily_ (homophily_metrics.py:236-244 / 247-255)."""
    d = torch.sqrt(torch.diag(gram))
    norm = d.reshape(-1, 1) * d.reshape(1, -1)
    norm = (norm > eps) * norm + eps * (norm <= eps)
    if n_layers == 1:
        arccos = torch.acos(gram / norm)
        root = torch.sqrt(norm.square() - gram.square())
        arccos[torch.isnan(arccos)] = 0
        root[torch.isnan(root)] = 0
        return 1 / math.pi * (gram * (math.pi - arccos) + root)
    return gram


def gntk_kernels(features, row, col, val, n, sample, n_layers):
    """(K_G/2, K_X/2) of homophily_metrics.py:232-257 for the node ids in `sample`."""
    x = torch.as_tensor(np.ascontiguousarray(features), dtype=torch.float32)
    z = torch.from_numpy(spmm(row, col, val, n, features))
    sample = torch.as_tensor(np.asarray(sample))
    zs, xs = z[sample], x[sample]
    kg = _arccos_kernel(zs @ zs.T, n_layers)
    kx = _arccos_kernel(xs @ xs.T, n_layers)
    return (kg / 2).numpy(), (kx / 2).numpy()


# ---------------------------------------------------------------------------
# util_funcs.py helpers used by the KR metric
# ---------------------------------------------------------------------------
def random_disassortative_splits(labels, num_classes, training_percentage=0.6):
    """util_funcs.py:454-475: per-class shuffled 60/20/20 split; consumes torch's global RNG."""
    labels = torch.as_tensor(labels)
    num_classes = int(num_classes)
    indices = []
    for i in range(num_classes):
        index = torch.nonzero(labels == i).view(-1)
        indices.append(index[torch.randperm(index.size(0))])
    per_class = int(round(training_percentage * (labels.size(0) / num_classes)))
    val_lb = int(round(0.2 * labels.size(0)))
    train_index = torch.cat([i[:per_class] for i in indices], dim=0)
    rest = torch.cat([i[per_class:] for i in indices], dim=0)
    rest = rest[torch.randperm(rest.size(0))]

    def mask(ix):
        m = torch.zeros(labels.size(0), dtype=torch.bool)
        m[ix] = True
        return m

    return mask(train_index), mask(rest[:val_lb]), mask(rest[val_lb:])


def accuracy(labels, output):
    """util_funcs.py:393-397."""
    preds = output.max(1)[1].type_as(labels)
    return preds.eq(labels).double().sum() / len(labels)


def kr_metric(features, row, col, val, n, labels, sample_max, base_classifier="kernel_reg1", epochs=100, trace=None,
              gram_fn=None, z=None):
    """`classifier_based_performance_metric` p-value.  homophily_metrics.py:260-349.

    Kernel-regression ('kernel_reg0' / 'kernel_reg1') and Gaussian naive Bayes
    ('gnb') classifiers; consumes torch's global RNG exactly like the reference.

    trace: optional list; one dict per epoch is appended (accuracies, validation ids, arg-max predictions and, for
    the kernel classifiers, the two kernel matrices with the train / validation index sets) -- what the parity tests
    compare instead of the p-value alone.  gram_fn: replaces `z @ z.T` (float32 torch.mm as in the reference) to
    probe how the predictions react to the rounding of the Gram matrix.  z: the propagated features A X when the
    caller forms them differently (homophily_plot.py:278-368 multiplies a DENSE adjacency: `plot_kr_metric`).
    """
    from scipy.stats import ttest_ind
    from sklearn.naive_bayes import GaussianNB

    gram_fn = gram_fn or (lambda m: m @ m.T)
    labels = torch.as_tensor(np.asarray(labels)).flatten().long()
    x = torch.as_tensor(np.ascontiguousarray(features), dtype=torch.float32)
    if z is None:
        z = torch.from_numpy(spmm(row, col, val, n, features))        # recomputed per epoch upstream
    z = torch.as_tensor(z, dtype=torch.float32)
    g_res, x_res, diff = torch.zeros(epochs), torch.zeros(epochs), torch.zeros(epochs)
    c = int(labels.max()) + 1
    for j in range(epochs):
        if n <= sample_max:
            sample = torch.arange(n)
            labels_sample = labels
        else:
            m, _, _ = random_disassortative_splits(labels, c, sample_max / n)          # :274
            sample = m
            labels_sample = labels[m]
        onehot = torch.eye(c)[labels][sample]
        tr, va, te = random_disassortative_splits(labels_sample, int(labels_sample.max()) + 1)   # :278
        va = va + te                                                                   # :279
        rec = {"va": va.clone(), "tr": tr.clone(), "labels_val": labels_sample[va].clone()}
        if base_classifier in ("kernel_reg0", "kernel_reg1"):
            nl = 0 if base_classifier == "kernel_reg0" else 1
            zs, xs = z[sample], x[sample]
            gram_g, gram_x = gram_fn(zs), gram_fn(xs)
            kg = _arccos_kernel(gram_g, nl) / 2
            kx = _arccos_kernel(gram_x, nl) / 2
            preds = []
            for k in (kg, kx):
                ktt = k[tr][:, tr]
                kvt = k[va][:, tr]
                alpha = torch.tensor(np.linalg.pinv(ktt.numpy())) @ onehot[tr]        # :286-290
                preds.append(kvt @ alpha)
            acc_g, acc_x = accuracy(labels_sample[va], preds[0]), accuracy(labels_sample[va], preds[1])
            rec.update(kg=kg, kx=kx, gram_g=gram_g, gram_x=gram_x, n_layers=nl, onehot_tr=onehot[tr].clone(),
                       pred_g=preds[0].max(1)[1], pred_x=preds[1].max(1)[1])
        elif base_classifier in ("gnb", "svm_rbf", "svm_poly", "svm_linear"):
            from sklearn import svm
            xs, zs = x[sample], z[sample]
            make = {"gnb": lambda: GaussianNB(),
                    "svm_rbf": lambda: svm.SVC(kernel="rbf", gamma=0.5, C=0.1),          # hm.py:317-319
                    "svm_poly": lambda: svm.SVC(kernel="poly", degree=3, C=1),           # hm.py:320-322
                    "svm_linear": lambda: svm.SVC(kernel="linear")}[base_classifier]    # hm.py:323-325
            cx, cg = make(), make()
            cx.fit(xs[tr], labels_sample[tr])
            cg.fit(zs[tr], labels_sample[tr])
            px, pg = torch.tensor(cx.predict(xs[va])), torch.tensor(cg.predict(zs[va]))
            acc_x = px.eq(labels_sample[va]).float().mean()
            acc_g = pg.eq(labels_sample[va]).float().mean()
            rec.update(pred_g=pg, pred_x=px)
        else:
            raise ValueError(base_classifier)
        diff[j] = float(acc_g > acc_x)
        g_res[j], x_res[j] = acc_g, acc_x
        rec.update(acc_g=float(acc_g), acc_x=float(acc_x))
        if trace is not None:
            trace.append(rec)
    _, p = ttest_ind(x_res.numpy(), g_res.numpy(), axis=0, equal_var=False, nan_policy="propagate")   # :340
    return float(p / 2) if diff.mean() <= 0.5 else float(1 - p / 2)                     # :343-347


def plot_kr_metric(features, adj_dense, labels, sample_max, base_classifier="kernel_reg1", epochs=100, trace=None):
    """hp.py:278-368: the synthetic_plot.py variant -- same epochs, splits, kernels and t-test as `kr_metric`; the
    adjacency is a dense float32 matrix, so A X is a dense torch.mm (`torch.spmm(adj, features)` on dense input)."""
    x = torch.as_tensor(np.ascontiguousarray(features), dtype=torch.float32)
    a = torch.as_tensor(adj_dense, dtype=torch.float32)
    n = int(a.shape[0])
    return kr_metric(features, None, None, None, n, labels, sample_max, base_classifier, epochs, trace=trace,
                     z=torch.mm(a, x))


def kr_unstable_nodes(gram, n_layers, tr, va, onehot_tr, rel=5e-7, rel_k=2e-6, trials=16, seed=0):
    """Validation nodes whose kernel-regression arg-max is decided by rounding noise.

    gram: the float32 Gram matrix of one epoch (Z Z^T of the sampled rows, before the arccos transform); tr / va:
    boolean masks; onehot_tr: one-hot train labels.  The prediction is
    `k[va][:, tr] @ pinv(k[tr][:, tr]) @ onehot` with k = arccos_kernel(gram) / 2 (hm.py:236-257, 286-290) and numpy's
    default rcond = 1e-15, i.e. singular values down to 1e-15 of the largest are inverted: rounding noise of the Gram
    matrix passes through `sqrt(norm^2 - g^2)` (which cancels wherever two rows are nearly parallel) and is then
    amplified by the condition number of the train block.  A node is reported unstable when its arg-max changes under
    any of `trials` symmetric perturbations of either kind:
      * G_ij += rel * sqrt(G_ii G_jj) * N(0, 1) -- the size of the float32 dot-product error bound (a few ulps of
        |z_i| |z_j|), which is what a different summation order of A X or of the Gram GEMM produces;
      * K_ij *= 1 + rel_k * N(0, 1) on the transformed kernel -- what a differently rounded acos / sqrt produces
        (the reference's own float32 transform is ~3e-5 away from its float64 evaluation).
    Returns a boolean tensor over the validation nodes.  The set is a Monte-Carlo LOWER bound: a node that flips under
    a few percent of the perturbations can be missed by 16 draws (in an epoch where 80% of the nodes are flagged, most of
    the rest are borderline too) -- `kr_flips_outside_unstable` therefore re-draws with more trials before it calls a
    differing prediction a violation.
    """
    gen = torch.Generator().manual_seed(seed)
    gram = gram.to(torch.float32)
    d = torch.sqrt(torch.diag(gram).clamp(min=0))
    scale = d.reshape(-1, 1) * d.reshape(1, -1)

    def predict_k(km):
        ktt, kvt = km[tr][:, tr], km[va][:, tr]
        return (kvt @ (torch.tensor(np.linalg.pinv(ktt.numpy())) @ onehot_tr)).max(1)[1]

    k0 = _arccos_kernel(gram, n_layers) / 2
    base = predict_k(k0)
    unstable = torch.zeros(base.shape[0], dtype=torch.bool)
    for _ in range(trials):
        e = torch.randn(gram.shape, generator=gen)
        e = (e + e.T) / 2 ** 0.5
        unstable |= predict_k(_arccos_kernel(gram + rel * scale * e, n_layers) / 2) != base
        unstable |= predict_k(k0 * (1 + rel_k * e)) != base
    return unstable


def kr_flips_outside_unstable(changed, gram, n_layers, tr, va, onehot_tr, escalate=(128,)):
    """(number of differing predictions that are NOT noise-decided, the unstable mask used, escalations).

    changed: boolean tensor over the validation nodes (this implementation's prediction != the reference's).  First the
    standard 16-trial unstable set; only if a differing prediction lies outside it, the analysis is repeated with more
    trials and another seed (the perturbation model is unchanged, only its sampling is denser) and the sets are united."""
    unstable = kr_unstable_nodes(gram, n_layers, tr, va, onehot_tr)
    n_esc = 0
    for trials in escalate:
        if not bool((changed & ~unstable).any()):
            break
        unstable = unstable | kr_unstable_nodes(gram, n_layers, tr, va, onehot_tr, trials=trials, seed=1 + n_esc)
        n_esc += 1
    return int((changed & ~unstable).sum()), unstable, n_esc


# ---------------------------------------------------------------------------
# utils/homophily_plot.py : the dense-adjacency variants used by synthetic_plot.py
# (restated on the stored entries (row, col, val) of the dense matrix; cited as hp.py:LINE)
# ---------------------------------------------------------------------------
def plot_edge_homophily(row, col, val, label_matrix):
    """hp.py:43-54: sum_{i!=j, a_ij>0} <l_i, l_j> / #{i!=j, a_ij>0}."""
    row, col = np.asarray(row, dtype=np.int64), np.asarray(col, dtype=np.int64)
    keep = (row != col) & (np.asarray(val) > 0)
    lab = np.asarray(label_matrix, dtype=np.float64)
    return np.float32((lab[row[keep]] * lab[col[keep]]).sum() / keep.sum())


def plot_node_homophily(row, col, labels, n):
    """hp.py:81-100: like node_homophily but self-loops are NOT removed (every stored nonzero counts)."""
    row, col = np.asarray(row, dtype=np.int64), np.asarray(col, dtype=np.int64)
    lab = np.asarray(labels, dtype=np.int64)
    deg = np.bincount(row, minlength=n).astype(np.float32)
    if row.size == 0 or int(row.max()) + 1 != n:
        raise RuntimeError("bincount length differs from the number of nodes")
    m = np.bincount(row[lab[row] == lab[col]], minlength=n).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        hs = m / deg
    return np.float32(np.mean(hs[deg != 0], dtype=np.float64))


def plot_compat_matrix(src, dst, labels):
    """hp.py:103-125: class compatibility matrix of an edge list, self-loops INCLUDED."""
    lab = np.asarray(labels, dtype=np.int64).reshape(-1)
    c = int(lab.max()) + 1
    ls, lt = lab[np.asarray(src, dtype=np.int64)], lab[np.asarray(dst, dtype=np.int64)]
    k = (ls >= 0) & (lt >= 0)
    h = np.bincount(ls[k] * c + lt[k], minlength=c * c).reshape(c, c).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return h / h.sum(1, keepdims=True)


def plot_class_homophily(row, col, val, labels, n):
    """hp.py:128-148 `our_measure(A, label)`: diagonal dropped, isolated nodes get a self-loop."""
    row, col = np.asarray(row, dtype=np.int64), np.asarray(col, dtype=np.int64)
    lab = np.asarray(labels, dtype=np.int64).reshape(-1)
    keep = (row != col) & (np.asarray(val) != 0)
    r, q = row[keep], col[keep]
    rs = np.bincount(r, weights=np.asarray(val, dtype=np.float64)[keep], minlength=n)
    iso = np.nonzero(rs == 0)[0]
    src, dst = np.concatenate([r, iso]), np.concatenate([q, iso])
    h = plot_compat_matrix(src, dst, lab)
    c = int(lab.max()) + 1
    nz = lab[lab >= 0]
    prop = np.unique(nz, return_counts=True)[1].astype(np.float32) / np.float32(nz.shape[0])
    v = np.float32(0)
    for k in range(c):
        add = h[k, k] - prop[k]
        if not np.isnan(add):
            v = np.float32(v + max(add, 0))
    return np.float32(v / np.float32(c - 1))


def plot_similarity(features, row, col, val, n, label_onehot, NTK=None, hard=None, LP=1, ifsum=1, idx_train=None):
    """hp.py:189-241: as `similarity`, optional NTK feature kernel, idx_train is an INDEX list applied after the Gram."""
    x = torch.as_tensor(np.ascontiguousarray(features), dtype=torch.float32)
    idx = torch.from_numpy(np.vstack([row, col]).astype(np.int64))
    a = torch.sparse_coo_tensor(idx, torch.as_tensor(np.asarray(val, dtype=np.float32)), (n, n)).coalesce().to_dense()
    label = torch.as_tensor(np.asarray(label_onehot), dtype=torch.float32)
    if NTK:
        k = torch.clamp(x @ x.T, 0, 1)
        k = (k * (torch.pi - torch.acos(k))) / (2 * torch.pi)
        inner = a @ k @ a.T
    else:
        z = a @ x
        inner = z @ z.T
    if idx_train is None:
        labels = label.argmax(1)
    else:
        it = torch.as_tensor(np.asarray(idx_train)).long()
        labels = label.argmax(1)[it]
        label = label[it]
        inner = inner[it][:, it]
    c = int(labels.max()) + 1
    w = torch.zeros(inner.shape[0], c)
    for i in range(c):
        cols = inner[:, labels == i]
        w[:, i] = cols.sum(1) if ifsum == 1 else cols.mean(1)
    own = w[torch.arange(labels.shape[0]), labels]
    if hard is None:
        nnodes, degs = (labels.shape[0], (label @ label.T).sum(1)) if ifsum == 1 else (c, 1)
        if LP == 1:
            ratio = (own / degs) / ((w.sum(1) - own) / (nnodes - degs))
            ratio[torch.isnan(ratio)] = 0
            return np.float32((ratio >= 1).float().mean().item())
        return np.float32((((w - w * label).sum(1) <= 0) & ((w * label).sum(1) >= 0)).float().mean().item())
    if LP == 1:
        return np.float32(w.argmax(1).eq(labels).float().mean().item())
    return np.float32((((w - w * label).max(1)[0] <= 0.0) & ((w * label).sum(1) >= 0)).float().mean().item())
