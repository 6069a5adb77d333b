#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference (utils/homophily_metrics.py, utils/util_funcs.py) is imported
from /root/reference as-is.  Its hard dependencies that are absent from this
image and are NOT on the graph-statistics path (torch_scatter.scatter_add,
dgl, torch_geometric, ogb, google_drive_downloader) are replaced by tiny stub
modules so that the import succeeds; `scatter_add` is the only stub that is
ever called and it forwards to torch's own Tensor.scatter_add_.

Every fixture stores the inputs (`in_*`) handed to the reference functions and
the outputs (`out_*`) those functions returned, so tests can replay the same
inputs through oracle/ (CPU, `-m "not gpu"`) and through the CUDA path
(`-m gpu`) without the reference being present.
"""
import os
import random
import sys
import types
import warnings

import numpy as np
import scipy.sparse as sp
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


# --------------------------------------------------------------------------
# the unmodified reference, imported through oracle/ref_shim.py (stubs for the absent third-party imports only)
# --------------------------------------------------------------------------
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_shim  # noqa: E402

_cwd = os.getcwd()
os.chdir(REF)  # the reference opens "data/ind.cora.x" relative to its root
hm, uf, hp = ref_shim.load(REF)
np.int = int  # utils/datasets.py:109 still says `np.int` (removed in numpy 1.24); an alias, not a patch of the reference
import utils.datasets as ds  # noqa: E402  (load_fb100_dataset for the LINKX fixtures; resolved by the shim's sys.modules)
from sklearn.preprocessing import label_binarize as _label_binarize  # noqa: E402
ds.label_binarize = _label_binarize  # datasets.py:118 calls it without importing it (LINKX's own file does import it)


class capture_kr_epochs:
    """Records the per-epoch accuracy vectors the reference hands to scipy's t-test (hm.py:340): the p-value is a
    function of them, and they are what a flipped validation prediction changes."""

    def __init__(self, mod):
        self.mod, self.real, self.x, self.g = mod, mod.ttest_ind, None, None

    def __enter__(self):
        def spy(a, b, *args, **kw):
            self.x, self.g = np.asarray(a, dtype=np.float64).copy(), np.asarray(b, dtype=np.float64).copy()
            return self.real(a, b, *args, **kw)
        self.mod.ttest_ind = spy
        return self

    def __exit__(self, *exc):
        self.mod.ttest_ind = self.real


def seed_all(s):
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)


def t2n(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def try_call(fn):
    """Return (value, '') or (nan, ExceptionTypeName)."""
    try:
        return fn(), ""
    except Exception as e:  # the reference's own error behaviour is part of the contract
        return None, type(e).__name__


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def coo_from_edges(row, col, n, vals=None):
    idx = torch.from_numpy(np.vstack([row, col]).astype(np.int64))
    v = torch.ones(idx.shape[1]) if vals is None else torch.from_numpy(vals.astype(np.float32))
    return torch.sparse_coo_tensor(idx, v, (n, n)).coalesce()


# --------------------------------------------------------------------------
# the label/structure metrics (homophily_metrics.py) on one sparse adjacency
# --------------------------------------------------------------------------
def structure_metrics(adj, labels, out, onehot=True):
    """adj: coalesced torch sparse COO (values irrelevant); labels: 1-D long."""
    c = int(labels.max()) + 1
    ei = adj.coalesce().indices()
    v, e = try_call(lambda: hm.edge_homophily(adj, labels))
    out["out_edge_homo"], out["err_edge_homo"] = (t2n(v) if e == "" else np.nan), e
    if onehot:
        oh = torch.eye(c)[labels.clamp(min=0)]
        out["out_edge_homo_onehot"] = t2n(hm.edge_homophily(adj, oh))
    v, e = try_call(lambda: hm.edge_homophily(adj, labels.numpy(), ignore_negative=True))
    out["out_edge_homo_ignore_negative"], out["err_edge_homo_ignore_negative"] = (
        (np.float64(v) if e == "" else np.nan), e)
    v, e = try_call(lambda: hm.node_homophily(adj, labels))
    out["out_node_homo"], out["err_node_homo"] = (t2n(v) if e == "" else np.nan), e
    v, e = try_call(lambda: hm.compact_matrix_edge_idx(ei, labels))
    out["out_compat"], out["err_compat"] = (t2n(v) if e == "" else np.zeros((c, c)) * np.nan), e
    v, e = try_call(lambda: hm.our_measure(ei, labels))
    out["out_class_homo"], out["err_class_homo"] = (t2n(v) if e == "" else np.nan), e
    v, e = try_call(lambda: hm.class_distribution(adj, labels))
    if e == "":
        out["out_p"], out["out_p_bar"], out["out_pc"] = t2n(v[0]), t2n(v[1]), t2n(v[2])
    out["err_class_distribution"] = e
    v, e = try_call(lambda: hm.adjusted_homo(adj, labels))
    out["out_adj_homo"], out["err_adj_homo"] = (t2n(v) if e == "" else np.nan), e
    v, e = try_call(lambda: hm.label_informativeness(adj, labels))
    out["out_label_info"], out["err_label_info"] = (t2n(v) if e == "" else np.nan), e


def gram_metrics(adj_spmm, features, labels, out, sample_n, tag=""):
    """similarity / gntk kernels / KR on `adj_spmm` (torch sparse, float32 values)."""
    c = int(labels.max()) + 1
    n = labels.shape[0]
    oh = torch.eye(c)[labels]
    out[f"out_soft_las{tag}"] = t2n(hm.similarity(oh, adj_spmm, oh, hard=None, LP=1))
    out[f"out_hard_las{tag}"] = t2n(hm.similarity(oh, adj_spmm, oh, hard=1, LP=1))
    out[f"out_soft_las_lp0{tag}"] = t2n(hm.similarity(oh, adj_spmm, oh, hard=None, LP=0))
    out[f"out_hard_las_lp0{tag}"] = t2n(hm.similarity(oh, adj_spmm, oh, hard=1, LP=0))
    out[f"out_soft_las_mean{tag}"] = t2n(hm.similarity(oh, adj_spmm, oh, hard=None, LP=1, ifsum=0))
    # idx_train branch (bool mask, as random_disassortative_splits returns)
    seed_all(7)
    mask = torch.zeros(n, dtype=torch.bool)
    mask[torch.randperm(n)[: max(8, n // 3)]] = True
    out[f"in_idx_train{tag}"] = t2n(mask)
    out[f"out_soft_las_idx{tag}"] = t2n(hm.similarity(oh, adj_spmm, oh, hard=None, LP=1, idx_train=mask))
    out[f"out_hard_las_idx{tag}"] = t2n(hm.similarity(oh, adj_spmm, oh, hard=1, LP=1, idx_train=mask))
    # GNTK / kernel-regression Gram matrices on a fixed sample
    seed_all(11)
    sample = np.sort(np.random.choice(n, size=min(sample_n, n), replace=False))
    out[f"in_gntk_sample{tag}"] = sample
    for nl in (0, 1):
        kg, kx = hm.gntk_homophily_(features, adj_spmm, sample, nl)
        out[f"out_gntk_KG_l{nl}{tag}"] = t2n(kg)
        out[f"out_gntk_KX_l{nl}{tag}"] = t2n(kx)


def kr_metric(adj_spmm, features, labels, out, sample_max, epochs, tag=""):
    for clf in ("kernel_reg0", "kernel_reg1", "gnb"):
        seed_all(2023)
        with capture_kr_epochs(hm) as cap:
            p, _ = hm.classifier_based_performance_metric(features, adj_spmm, labels, sample_max,
                                                          base_classifier=clf, epochs=epochs)
        out[f"out_kr_p_{clf}{tag}"] = np.float64(p)
        out[f"out_kr_acc_x_{clf}{tag}"], out[f"out_kr_acc_g_{clf}{tag}"] = cap.x, cap.g
    out[f"in_kr_sample_max{tag}"] = np.int64(sample_max)
    out[f"in_kr_epochs{tag}"] = np.int64(epochs)
    out[f"in_kr_seed{tag}"] = np.int64(2023)


def spmm_projection(adj, features, out, tag):
    """Store AX compactly: 24 seeded columns + row sums + a seeded 16-col random projection."""
    ax = torch.spmm(adj, features)
    g = torch.Generator().manual_seed(5)
    d = features.shape[1]
    cols = torch.randperm(d, generator=g)[: min(24, d)].sort().values
    proj = torch.randn(d, 16, generator=g)
    out[f"in_proj_cols_{tag}"] = t2n(cols)
    out[f"in_proj_mat_seed_{tag}"] = np.int64(5)
    out[f"out_ax_cols_{tag}"] = t2n(ax[:, cols])
    out[f"out_ax_rowsum_{tag}"] = t2n(ax.double().sum(1))
    out[f"out_ax_proj_{tag}"] = t2n(ax.double() @ proj.double())


# --------------------------------------------------------------------------
# case 1: Cora through the homophily_tests.py small-dataset flow
# --------------------------------------------------------------------------
def _small_dataset_flow(name, adj_sp, feats_sp, features_raw, labels, kr_sample_max, kr_epochs):
    """homophily_tests.py small-dataset flow (:78-87 normalisation, :108-137 metric dispatch) on one loaded dataset."""
    A = uf.sparse_mx_to_torch_sparse_tensor(adj_sp).coalesce()  # raw, binary, no normalisation
    n = labels.shape[0]
    out = {}
    out["in_n"] = np.int64(n)
    out["in_edge_index"] = t2n(A.indices()).astype(np.int32)
    out["in_labels"] = t2n(labels).astype(np.int64)
    fcoo = feats_sp.tocoo()
    out["in_feat_row"], out["in_feat_col"] = fcoo.row.astype(np.int32), fcoo.col.astype(np.int32)
    out["in_feat_val"] = fcoo.data.astype(np.float32)
    out["in_feat_dim"] = np.int64(features_raw.shape[1])

    features = uf.normalize_tensor(features_raw)  # homophily_tests.py:80
    out["out_features_rownorm_rowsum"] = t2n(features.double().sum(1))
    for sym in (0, 1):
        adjn = uf.normalize_tensor(torch.eye(n) + A.to_dense(), symmetric=sym).to_sparse().coalesce()  # :83-85
        o = {}
        structure_metrics(adjn, labels, o)
        o["out_adj_values"] = t2n(adjn.values())
        o["out_gen_edge_homo"] = t2n(hm.generalized_edge_homophily(adjn, features, labels))
        spmm_projection(adjn, features, o, "norm")
        for k, v in o.items():
            out[f"{k}__sym{sym}"] = v
    # aggregation homophily + KR are computed on the RAW adjacency (homophily_tests.py:120,135)
    gram_metrics(A, features_raw, labels, out, sample_n=96)
    kr_metric(A, features_raw, labels, out, sample_max=kr_sample_max, epochs=kr_epochs)
    # scipy-side normalisers of the LINKX flow (homophily_tests.py:99-104)
    a_sym = uf.sparse_mx_to_torch_sparse_tensor(uf.sys_normalized_adjacency(adj_sp)).coalesce()
    a_rw = uf.sparse_mx_to_torch_sparse_tensor(uf.row_normalized_adjacency(adj_sp)).coalesce()
    out["out_sys_norm_values"] = t2n(a_sym.values())
    out["out_row_norm_values"] = t2n(a_rw.values())
    out["out_sys_norm_index"] = t2n(a_sym.indices()).astype(np.int32)
    spmm_projection(a_sym, features, out, "sys")
    spmm_projection(a_rw, features, out, "rw")
    save(name, **out)


def case_cora():
    adj_sp, feats, labels = uf.load_data("cora")  # util_funcs.py:49
    labels = torch.LongTensor(np.argmax(labels, axis=-1))
    _small_dataset_flow("cora", adj_sp, sp.csr_matrix(feats), torch.FloatTensor(np.asarray(feats.todense())), labels,
                        kr_sample_max=500, kr_epochs=6)


# --------------------------------------------------------------------------
# case 1b: the other datasets shipped with the reference, loaded by its own loaders: citeseer (Planetoid files,
# util_funcs.py:49) and the heterophilous WebKB / actor graphs (new_data/*, full_load_data_large :291-338)
# --------------------------------------------------------------------------
def case_datasets():
    import networkx as nx  # the reference's loader needs it (full_load_data_large)
    assert nx is not None
    for name, sample_max in (("citeseer", 500), ("texas", 60), ("cornell", 60), ("wisconsin", 80), ("film", 300)):
        if name == "citeseer":
            adj_sp, feats, onehot = uf.load_data("citeseer")
            labels = torch.LongTensor(np.argmax(onehot, axis=-1))
            features_raw = torch.FloatTensor(np.asarray(feats.todense()))
        else:
            adj_t, features_raw, labels = uf.full_load_data_large(name)
            adj_t = adj_t.coalesce()
            idx = t2n(adj_t.indices())
            adj_sp = sp.coo_matrix((t2n(adj_t.values()).astype(np.float64), (idx[0], idx[1])), shape=tuple(adj_t.shape))
            features_raw, labels = features_raw.float().cpu(), labels.long().cpu()
        _small_dataset_flow(f"ds_{name}", adj_sp, sp.csr_matrix(features_raw.numpy()), features_raw, labels,
                            kr_sample_max=sample_max, kr_epochs=4)


# --------------------------------------------------------------------------
# case 2: data_synthesis graphs (ACM-GNN generator output shipped with the reference)
# --------------------------------------------------------------------------
def case_synthetic():
    for nedge, h, s in ((800, 0.05, 0), (800, 0.5, 3), (4000, 0.2, 1), (4000, 0.9, 2)):
        base = f"{REF}/data_synthesis/{nedge}/{h}"
        adj = torch.load(f"{base}/adj_{h}_{s}.pt", weights_only=False).coalesce()
        lab = torch.load(f"{base}/label_{h}_{s}.pt", weights_only=False).to_dense()
        labels = torch.argmax(lab, 1)
        n, c = lab.shape
        # the shipped feature files are empty placeholders: draw class-conditional Gaussians instead
        g = torch.Generator().manual_seed(100 + s)
        centers = torch.randn(c, 32, generator=g)
        features = (centers[labels] + 1.5 * torch.randn(n, 32, generator=g)).float()
        A = coo_from_edges(t2n(adj.indices())[0], t2n(adj.indices())[1], n)  # float32 values
        out = {"in_n": np.int64(n), "in_edge_index": t2n(A.indices()).astype(np.int32),
               "in_labels": t2n(labels).astype(np.int64), "in_features": t2n(features)}
        # sparse-side normalisation with self-loops, LINKX flow
        a_sp = sp.coo_matrix((np.ones(A._nnz()), (t2n(A.indices())[0], t2n(A.indices())[1])), shape=(n, n))
        for sym, fn in ((1, uf.sys_normalized_adjacency), (0, uf.row_normalized_adjacency)):
            adjn = uf.sparse_mx_to_torch_sparse_tensor(fn(a_sp)).coalesce()
            o = {}
            structure_metrics(adjn, labels, o)
            o["out_adj_values"] = t2n(adjn.values())
            seed_all(3)  # the >=75000-entry branch draws edges with random.sample
            o["out_gen_edge_homo"] = np.float64(hm.generalized_edge_homophily(adjn, features, labels))
            o["in_gen_seed"] = np.int64(3)
            spmm_projection(adjn, features, o, "norm")
            gram_metrics(adjn, features, labels, o, sample_n=64, tag="_norm")
            for k, v in o.items():
                out[f"{k}__sym{sym}"] = v
        gram_metrics(A, features, labels, out, sample_n=64)
        kr_metric(A, features, labels, out, sample_max=300, epochs=4)
        save(f"syn_{nedge}_{h}_{s}", **out)


# --------------------------------------------------------------------------
# case 3: small random graphs exercising the edge cases
# --------------------------------------------------------------------------
def case_edge_cases():
    rng = np.random.default_rng(42)

    def rand_graph(n, m, c, self_loops, neg_frac, undirected=True, isolated=0):
        r = rng.integers(0, n - isolated, m)
        q = rng.integers(0, n - isolated, m)
        keep = r != q
        r, q = r[keep], q[keep]
        if undirected:
            r, q = np.concatenate([r, q]), np.concatenate([q, r])
        if self_loops == "all":
            r, q = np.concatenate([r, np.arange(n)]), np.concatenate([q, np.arange(n)])
        elif self_loops == "some":
            k = np.arange(0, n, 3)
            r, q = np.concatenate([r, k]), np.concatenate([q, k])
        labels = rng.integers(0, c, n)
        labels[rng.permutation(n)[:c]] = np.arange(c)  # every class present
        if neg_frac > 0:
            labels[rng.random(n) < neg_frac] = -1
        return r, q, labels

    specs = {
        # name: (n, m, c, self_loops, neg_frac, undirected, isolated)
        "ec_selfloops_all": (300, 900, 4, "all", 0.0, True, 0),
        "ec_no_selfloops": (300, 900, 4, "none", 0.0, True, 0),
        "ec_some_selfloops": (257, 700, 3, "some", 0.0, True, 0),
        "ec_negative_labels": (400, 1500, 2, "all", 0.25, True, 0),
        "ec_directed": (200, 800, 5, "all", 0.0, False, 0),
        "ec_isolated_tail": (128, 300, 3, "none", 0.0, True, 9),
        "ec_skewed": (500, 0, 6, "all", 0.0, True, 0),
        "ec_many_classes": (600, 4000, 40, "all", 0.0, True, 0),
    }
    for name, (n, m, c, sl, neg, und, iso) in specs.items():
        if name == "ec_skewed":  # star-heavy graph: a few hubs hold most edges
            hubs = rng.integers(0, 5, 3000)
            leaves = rng.integers(5, n, 3000)
            r, q = np.concatenate([hubs, leaves, np.arange(n)]), np.concatenate([leaves, hubs, np.arange(n)])
            labels = rng.integers(0, c, n)
            labels[:c] = np.arange(c)
        else:
            r, q, labels = rand_graph(n, m, c, sl, neg, und, iso)
        vals = rng.random(r.shape[0]).astype(np.float32) + 0.1
        A = coo_from_edges(r, q, n, vals)  # coalesce() sums duplicate entries
        labels_t = torch.from_numpy(labels.astype(np.int64))
        features = torch.from_numpy(rng.standard_normal((n, 19)).astype(np.float32))
        features[::17] = 0  # zero rows: cosine similarity NaN/0 handling
        out = {"in_n": np.int64(n), "in_edge_index": t2n(A.indices()).astype(np.int32),
               "in_edge_values": t2n(A.values()), "in_labels": labels.astype(np.int64),
               "in_features": t2n(features)}
        structure_metrics(A, labels_t, out, onehot=(neg == 0))
        out["out_gen_edge_homo"] = t2n(hm.generalized_edge_homophily(A, features, labels_t))
        seed_all(3)
        v = hm.generalized_edge_homophily(A, features, labels_t, sample_max=200, iteration=3)
        out["out_gen_edge_homo_sampled"] = np.float64(v)
        out["in_gen_sampled_args"] = np.array([3, 200, 3], dtype=np.int64)  # seed, sample_max, iteration
        spmm_projection(A, features, out, "w")
        if neg == 0:
            gram_metrics(A, features, labels_t, out, sample_n=48)
        save(name, **out)


# --------------------------------------------------------------------------
# case 4: the dense-adjacency variants of utils/homophily_plot.py through the synthetic_plot.py flow
# --------------------------------------------------------------------------
def case_plot_variants():
    for nedge, h, s in ((800, 0.3, 1), (4000, 0.6, 4)):
        base = f"{REF}/data_synthesis/{nedge}/{h}"
        adj_raw = torch.load(f"{base}/adj_{h}_{s}.pt", weights_only=False).to_dense().clone().detach().float()
        label = torch.load(f"{base}/label_{h}_{s}.pt", weights_only=False).to_dense().clone().detach().float()
        labels = torch.argmax(label, 1)
        n, c = label.shape
        g = torch.Generator().manual_seed(300 + s)
        centers = torch.rand(c, 24, generator=g)
        feats_raw = (centers[labels] + 0.8 * torch.rand(n, 24, generator=g)).float()        # non-negative, like bag-of-words
        features = torch.tensor(uf.preprocess_features(feats_raw.clone())).clone().detach()    # synthetic_plot.py:81-82
        adj = torch.tensor(uf.normalize(adj_raw + torch.eye(n)))                              # synthetic_plot.py:94
        raw_ei = adj_raw.to_sparse().coalesce().indices()
        out = {"in_n": np.int64(n), "in_edge_index": t2n(raw_ei).astype(np.int32), "in_labels": t2n(labels),
               "in_features_raw": t2n(feats_raw), "out_features": t2n(features), "out_adj_rowsum": t2n(adj.double().sum(1)),
               "out_adj_diag": t2n(torch.diag(adj))}
        out["out_edge_homo"] = t2n(hp.edge_homophily(adj, label))
        out["out_node_homo"] = t2n(hp.node_homophily(adj, labels))
        out["out_class_homo"] = t2n(hp.our_measure(adj, labels))
        out["out_soft_las"] = t2n(hp.similarity(label, adj, label, NTK=None, hard=None, LP=1))
        out["out_hard_las"] = t2n(hp.similarity(label, adj, label, NTK=None, hard=1, LP=1))
        idx = torch.arange(0, n, 3)
        out["in_idx_train"] = t2n(idx)
        out["out_soft_las_idx"] = t2n(hp.similarity(label, adj, label, NTK=None, hard=None, LP=1, idx_train=idx))
        out["out_soft_las_ntk"] = t2n(hp.similarity(features / features.norm(dim=1, keepdim=True), adj, label, NTK=True,
                                                    hard=None, LP=1))
        out["out_adj_homo"] = t2n(hp.adjusted_homo(adj, label))
        out["out_label_info"] = t2n(hp.label_informativeness(adj, label))
        out["out_gen_edge_homo"] = t2n(hp.generalized_edge_homophily(adj, features, label))
        p, p_bar, pc = hp.class_distribution(adj, labels)
        out["out_p"], out["out_p_bar"], out["out_pc"] = t2n(p), t2n(p_bar), t2n(pc)
        ei_t = adj.nonzero()
        out["out_compat"] = t2n(hp.compact_matrix_edge_idx(ei_t, labels))
        for clf in ("kernel_reg0", "kernel_reg1", "gnb"):
            seed_all(77)
            with capture_kr_epochs(hp) as cap:
                out[f"out_kr_p_{clf}"] = np.float64(hp.classifier_based_performance_metric(
                    features, adj, labels, sample_max=300, base_classifier=clf, epochs=4))
            out[f"out_kr_acc_x_{clf}"], out[f"out_kr_acc_g_{clf}"] = cap.x, cap.g
        out["in_kr"] = np.array([77, 300, 4], dtype=np.int64)  # seed, sample_max, epochs
        save(f"plot_syn_{nedge}_{h}_{s}", **out)


# --------------------------------------------------------------------------
# case 5: the remaining normalisers / edge statistics of util_funcs.py that sit next to the path
# (normalize :29, preprocess_features :39, normalize_adj :429, dataset_edge_balance :439)
# --------------------------------------------------------------------------
def case_util_norm():
    rng = np.random.default_rng(2024)
    n, c = 350, 5
    r, q = rng.integers(0, n - 6, 2600), rng.integers(0, n - 6, 2600)      # directed, last 6 rows/cols empty
    vals = (rng.random(2600) + 0.25).astype(np.float64)
    a = sp.coo_matrix((vals, (r, q)), shape=(n, n)).tocsr()                # duplicates summed
    a.sort_indices()
    coo = a.tocoo()
    labels = rng.integers(0, c, n)
    labels[:c] = np.arange(c)
    feats = sp.random(n, 40, density=0.15, random_state=7, format="csr", dtype=np.float64)
    feats.data = np.ceil(feats.data * 4)                                   # bag-of-words like counts, some empty rows
    out = {"in_n": np.int64(n), "in_row": coo.row.astype(np.int32), "in_col": coo.col.astype(np.int32),
           "in_val": coo.data.astype(np.float64), "in_labels": labels.astype(np.int64)}
    fc = feats.tocoo()
    out["in_feat_row"], out["in_feat_col"], out["in_feat_val"] = fc.row.astype(np.int32), fc.col.astype(np.int32), fc.data
    out["in_feat_dim"] = np.int64(40)
    # scipy-sparse flow (full_load_data: util_funcs.py:189-190)
    m = sp.csr_matrix(uf.normalize(a)); m.sort_indices()
    out["out_normalize_indptr"], out["out_normalize_indices"], out["out_normalize_data"] = m.indptr, m.indices, m.data
    f = sp.csr_matrix(uf.preprocess_features(feats)); f.sort_indices()
    out["out_preprocess_indptr"], out["out_preprocess_indices"], out["out_preprocess_data"] = f.indptr, f.indices, f.data
    # dense torch flow (synthetic_plot.py:82,92)
    dense = torch.from_numpy(a.toarray()).float()
    out["out_normalize_dense"] = np.asarray(uf.normalize(dense + torch.eye(n)))
    out["out_preprocess_dense"] = np.asarray(uf.preprocess_features(torch.from_numpy(feats.toarray()).float()))
    na = uf.normalize_adj(a).tocsr(); na.sort_indices()
    out["out_normalize_adj_indptr"], out["out_normalize_adj_indices"], out["out_normalize_adj_data"] = (
        na.indptr, na.indices, na.data)
    nodes, bal = uf.dataset_edge_balance(a.toarray(), torch.from_numpy(labels))
    out["out_balance_nodes"], out["out_balance"] = nodes, bal
    b = (a != 0).astype(np.float64)
    nodes_b, bal_b = uf.dataset_edge_balance(b.toarray(), torch.from_numpy(labels))
    out["out_balance_binary"] = bal_b
    save("util_norm", **out)


# --------------------------------------------------------------------------
# case 6: the LINKX-family graphs the reference ships (data/facebook100/*.mat) through ITS loader
# (utils/datasets.py:105-128 load_fb100_dataset: gender label, 0 -> -1 = unlabelled; one-hot features) and the
# large-dataset flow of homophily_tests.py:88-137 (f.normalize(p=1), sys_/row_normalized_adjacency, metric dispatch,
# 10 x similarity on class-balanced samples, KR).  `to_undirected` (torch_geometric, util_funcs.py:239) is a no-op on
# these symmetric matrices up to coalescing, which `sparse_coo_tensor(...).coalesce()` does here.
# --------------------------------------------------------------------------
def case_linkx():
    import torch.nn.functional as f
    for fname, kr_sample_max, num_sample in (("Reed98", 300, 10000), ("Amherst41", 400, 1000),
                                             ("Johns Hopkins55", 500, 2000), ("Cornell5", 500, 10000)):
        dataset = ds.load_fb100_dataset(fname)
        ei = dataset.graph["edge_index"]
        n = int(dataset.graph["num_nodes"])
        und = torch.sparse_coo_tensor(torch.cat([ei, ei.flip(0)], 1), torch.ones(2 * ei.shape[1]), (n, n)).coalesce()
        row, col = t2n(und.indices())
        adj_sp = sp.coo_matrix((np.ones(row.shape[0]), (row, col)), shape=(n, n))          # uf.py:242
        features_raw, labels = dataset.graph["node_feat"], dataset.label.long()
        A = uf.sparse_mx_to_torch_sparse_tensor(adj_sp).coalesce()                         # uf.py:357
        big = fname == "Cornell5"
        out = {"in_n": np.int64(n), "in_labels": t2n(labels).astype(np.int64),
               "in_features": t2n(features_raw).astype(np.uint8), "in_num_sample": np.int64(num_sample)}
        up = row < col                                                                     # symmetric: store one triangle
        out["in_edges_upper"] = np.vstack([row[up], col[up]]).astype(np.int32)
        assert 2 * int(up.sum()) == row.shape[0]
        features = f.normalize(features_raw, p=1, dim=1)                                   # homophily_tests.py:95
        out["out_features_l1"] = t2n(features[:: max(1, n // 64)])
        for sym, fn in ((1, uf.sys_normalized_adjacency), (0, uf.row_normalized_adjacency)):
            adjn = uf.sparse_mx_to_torch_sparse_tensor(fn(adj_sp)).coalesce()              # :98-104
            o = {}
            structure_metrics(adjn, labels, o)                                             # :108-116 metric dispatch
            o["out_adj_values_sum"] = np.float64(adjn.values().double().sum())
            o["out_adj_values_head"] = t2n(adjn.values()[:4096])
            seed_all(3)
            o["out_gen_edge_homo"] = np.float64(hm.generalized_edge_homophily(adjn, features, labels))  # :117-118
            spmm_projection(adjn, features, o, "norm")                                     # SGC-1 propagation
            for k, v in o.items():
                out[f"{k}__sym{sym}"] = v
        # aggregation homophily: 10 x similarity on samples (homophily_tests.py:119-132)
        label_onehot = torch.eye(int(labels.max()) + 1)[labels]
        for hard, key in ((None, "soft"), (1, "hard")):
            seed_all(19)
            las = np.zeros(10)
            for i in range(10):
                if n >= num_sample:
                    idx_train, _, _ = uf.random_disassortative_splits(labels, labels.max() + 1, num_sample / n)
                else:
                    idx_train = None
                las[i] = 2 * float(hm.similarity(label_onehot, A, label_onehot, hard=hard, LP=1, idx_train=idx_train)) - 1
            out[f"out_agg_homo_{key}_las"] = las
        out["in_agg_seed"] = np.int64(19)
        # KR (homophily_tests.py:133-137), few epochs
        if not big:
            kr_metric(A, features_raw, labels, out, sample_max=kr_sample_max, epochs=4)
        save("linkx_" + fname.replace(" ", "_"), **out)


# --------------------------------------------------------------------------
# case 7: the SVM base classifiers of the KR metric (hm.py:312-333) on texas; inputs are those of ds_texas.npz
# --------------------------------------------------------------------------
def case_svm():
    adj_t, features_raw, labels = uf.full_load_data_large("texas")
    adj_t = adj_t.coalesce()
    features_raw, labels = features_raw.float().cpu(), labels.long().cpu()
    out = {"in_kr_seed": np.int64(404), "in_kr_sample_max": np.int64(120), "in_kr_epochs": np.int64(3)}
    for clf in ("svm_rbf", "svm_poly", "svm_linear"):
        seed_all(404)
        with capture_kr_epochs(hm) as cap:
            p, _ = hm.classifier_based_performance_metric(features_raw, adj_t, labels, 120, base_classifier=clf, epochs=3)
        out[f"out_kr_p_{clf}"] = np.float64(p)
        out[f"out_kr_acc_x_{clf}"], out[f"out_kr_acc_g_{clf}"] = cap.x, cap.g
    save("svm_texas", **out)


if __name__ == "__main__":
    cases = {"plot": case_plot_variants, "cora": case_cora, "datasets": case_datasets, "synthetic": case_synthetic,
             "edge": case_edge_cases, "util_norm": case_util_norm, "linkx": case_linkx, "svm": case_svm}
    for name in (sys.argv[1:] or list(cases)):     # no argument = regenerate everything
        cases[name]()
    os.chdir(_cwd)
