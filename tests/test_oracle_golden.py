"""Pin oracle/ref_port.py against outputs of the UNMODIFIED reference (tests/golden/*.npz).

CPU only.  Integer-derived quantities must agree to float32 round-off; float
paths within 1e-5 relative (north_star asks 1e-4 of the CUDA path; the oracle
is held tighter).
"""
import random

import numpy as np
import pytest
import torch

import _golden as G
from oracle import ref_port as O

RTOL = 1e-5
ATOL = 1e-6


def close(a, b, rtol=RTOL, atol=ATOL):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64),
                               rtol=rtol, atol=atol, equal_nan=True)


def check_structure(z, row, col, labels, n, suffix=""):
    g = lambda k: z[k + suffix]  # noqa: E731
    e = lambda k: str(z[k + suffix])  # noqa: E731
    close(O.edge_homophily(row, col, labels), g("out_edge_homo"))
    if "out_edge_homo_onehot" + suffix in z.files:
        c = int(labels.max()) + 1
        close(O.edge_homophily(row, col, np.eye(c, dtype=np.float32)[labels]), g("out_edge_homo_onehot"))
    close(O.edge_homophily(row, col, labels, ignore_negative=True), g("out_edge_homo_ignore_negative"))
    if e("err_node_homo"):
        with pytest.raises(RuntimeError):
            O.node_homophily(row, col, labels, n)
    else:
        close(O.node_homophily(row, col, labels, n), g("out_node_homo"))
    close(O.compat_matrix(row, col, labels), g("out_compat"))
    close(O.class_homophily(row, col, labels), g("out_class_homo"))
    if e("err_class_distribution"):
        for fn in (O.class_distribution, O.adjusted_homo, O.label_informativeness):
            with pytest.raises(IndexError):
                fn(row, col, labels, n)
    else:
        p, p_bar, pc = O.class_distribution(row, col, labels, n)
        close(p, g("out_p"))
        close(p_bar, g("out_p_bar"))
        close(pc, g("out_pc"))
        close(O.adjusted_homo(row, col, labels, n), g("out_adj_homo"), rtol=1e-4)
        close(O.label_informativeness(row, col, labels, n), g("out_label_info"), rtol=1e-4, atol=1e-5)


def check_spmm(z, row, col, val, n, x, tag):
    ax = O.spmm(row, col, val, n, x)
    cols, proj = G.proj_matrix(x.shape[1])
    assert np.array_equal(cols, z[f"in_proj_cols_{tag}"])
    close(ax[:, cols], z[f"out_ax_cols_{tag}"])
    close(ax.astype(np.float64).sum(1), z[f"out_ax_rowsum_{tag}"], rtol=1e-5, atol=1e-6)
    close(ax.astype(np.float64) @ proj, z[f"out_ax_proj_{tag}"], rtol=1e-5, atol=1e-5)


def check_gram(z, row, col, val, n, x, labels, tag=""):
    c = int(labels.max()) + 1
    oh = np.eye(c, dtype=np.float32)[labels]
    kw = dict(row=row, col=col, val=val, n=n, label_onehot=oh)
    # indicator means: one flipped node moves the score by 1/n, so compare exactly-ish
    close(O.similarity(oh, hard=None, LP=1, **kw), z[f"out_soft_las{tag}"], rtol=0, atol=1e-6)
    close(O.similarity(oh, hard=1, LP=1, **kw), z[f"out_hard_las{tag}"], rtol=0, atol=1e-6)
    close(O.similarity(oh, hard=None, LP=0, **kw), z[f"out_soft_las_lp0{tag}"], rtol=0, atol=1e-6)
    close(O.similarity(oh, hard=1, LP=0, **kw), z[f"out_hard_las_lp0{tag}"], rtol=0, atol=1e-6)
    close(O.similarity(oh, hard=None, LP=1, ifsum=0, **kw), z[f"out_soft_las_mean{tag}"], rtol=0, atol=1e-6)
    m = z[f"in_idx_train{tag}"]
    close(O.similarity(oh, hard=None, LP=1, idx_train=m, **kw), z[f"out_soft_las_idx{tag}"], rtol=0, atol=1e-6)
    close(O.similarity(oh, hard=1, LP=1, idx_train=m, **kw), z[f"out_hard_las_idx{tag}"], rtol=0, atol=1e-6)
    sample = z[f"in_gntk_sample{tag}"]
    for nl in (0, 1):
        kg, kx = O.gntk_kernels(x, row, col, val, n, sample, nl)
        scale = max(1.0, float(np.abs(z[f"out_gntk_KG_l{nl}{tag}"]).max()))
        close(kg, z[f"out_gntk_KG_l{nl}{tag}"], rtol=1e-4, atol=1e-5 * scale)
        scale = max(1.0, float(np.abs(z[f"out_gntk_KX_l{nl}{tag}"]).max()))
        close(kx, z[f"out_gntk_KX_l{nl}{tag}"], rtol=1e-4, atol=1e-5 * scale)


def check_kr(z, row, col, val, n, x, labels, tag=""):
    for clf in ("kernel_reg0", "kernel_reg1", "gnb"):
        seed = int(z[f"in_kr_seed{tag}"])
        random.seed(seed), np.random.seed(seed), torch.manual_seed(seed)
        p = O.kr_metric(x, row, col, val, n, labels, int(z[f"in_kr_sample_max{tag}"]),
                        base_classifier=clf, epochs=int(z[f"in_kr_epochs{tag}"]))
        close(p, z[f"out_kr_p_{clf}{tag}"], rtol=1e-3, atol=1e-9)


# ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["cora"] + G.names("ds_"))
def test_reference_datasets(name):
    """Cora + the other datasets the reference ships (citeseer, texas, cornell, wisconsin, film), loaded by the
    reference's own loaders and pushed through the homophily_tests.py small-dataset flow."""
    z = G.load(name)
    n = int(z["in_n"])
    labels = z["in_labels"]
    ei = z["in_edge_index"].astype(np.int64)
    x_raw = G.cora_dense_features(z)
    x = O.normalize_tensor(x_raw).numpy()
    close(x.astype(np.float64).sum(1), z["out_features_rownorm_rowsum"])
    for sym in (0, 1):
        row, col, val = G.dense_normalized_with_self_loops(z, sym)
        close(val, z[f"out_adj_values__sym{sym}"])
        check_structure(z, row, col, labels, n, f"__sym{sym}")
        close(O.generalized_edge_homophily(row, col, x, n), z[f"out_gen_edge_homo__sym{sym}"], rtol=1e-4)
        check_spmm(z, row, col, val, n, x, f"norm__sym{sym}")
    ones = np.ones(ei.shape[1], np.float32)
    check_gram(z, ei[0], ei[1], ones, n, x_raw, labels)
    check_kr(z, ei[0], ei[1], ones, n, x_raw, labels)
    # scipy normalisers of the LINKX flow
    r, c, v = O.sys_normalized_adjacency(ei[0], ei[1], ones, n)
    assert np.array_equal(np.vstack([r, c]), z["out_sys_norm_index"])
    close(v, z["out_sys_norm_values"], rtol=1e-6)
    check_spmm(z, r, c, v, n, x, "sys")
    r, c, v = O.row_normalized_adjacency(ei[0], ei[1], ones, n)
    close(v, z["out_row_norm_values"], rtol=1e-6)
    check_spmm(z, r, c, v, n, x, "rw")


@pytest.mark.parametrize("name", G.names("syn_"))
def test_synthetic(name):
    z = G.load(name)
    n = int(z["in_n"])
    labels = z["in_labels"]
    ei = z["in_edge_index"].astype(np.int64)
    x = z["in_features"]
    ones = np.ones(ei.shape[1], np.float32)
    for sym, fn in ((1, O.sys_normalized_adjacency), (0, O.row_normalized_adjacency)):
        row, col, val = fn(ei[0], ei[1], ones, n)
        sfx = f"__sym{sym}"
        close(val, z["out_adj_values" + sfx], rtol=1e-6)
        check_structure(z, row, col, labels, n, sfx)
        seed = int(z["in_gen_seed" + sfx])
        random.seed(seed), np.random.seed(seed), torch.manual_seed(seed)
        close(O.generalized_edge_homophily(row, col, x, n), z["out_gen_edge_homo" + sfx], rtol=1e-4, atol=1e-6)
        check_spmm(z, row, col, val, n, x, "norm" + sfx)
        # tag order in the fixture is "<key>_norm__symS"
        zz = {k.replace("_norm" + sfx, "") + "@": z[k] for k in z.files if k.endswith("_norm" + sfx)}

        class _Z(dict):
            files = list(zz)
        check_gram(_Z(zz), row, col, val, n, x, labels, tag="@")
    check_gram(z, ei[0], ei[1], ones, n, x, labels)
    check_kr(z, ei[0], ei[1], ones, n, x, labels)


@pytest.mark.parametrize("name", G.names("ec_"))
def test_edge_cases(name):
    z = G.load(name)
    n = int(z["in_n"])
    labels = z["in_labels"]
    ei = z["in_edge_index"].astype(np.int64)
    val = z["in_edge_values"]
    x = z["in_features"]
    check_structure(z, ei[0], ei[1], labels, n)
    close(O.generalized_edge_homophily(ei[0], ei[1], x, n), z["out_gen_edge_homo"], rtol=1e-4, atol=1e-6)
    seed, smax, it = (int(v) for v in z["in_gen_sampled_args"])
    random.seed(seed), np.random.seed(seed), torch.manual_seed(seed)
    close(O.generalized_edge_homophily(ei[0], ei[1], x, n, sample_max=smax, iteration=it),
          z["out_gen_edge_homo_sampled"], rtol=1e-4, atol=1e-6)
    check_spmm(z, ei[0], ei[1], val, n, x, "w")
    if "out_soft_las" in z.files:
        check_gram(z, ei[0], ei[1], val, n, x, labels)


@pytest.mark.parametrize("name", G.names("plot_"))
def test_plot_variants(name):
    """utils/homophily_plot.py (dense-adjacency variants, synthetic_plot.py flow)."""
    z = G.load(name)
    n = int(z["in_n"])
    labels = z["in_labels"]
    a = G.plot_flow_adjacency(z)
    close(a.double().sum(1), z["out_adj_rowsum"])
    close(torch.diag(a), z["out_adj_diag"], rtol=1e-6)
    sp_ = a.to_sparse().coalesce()
    row, col, val = sp_.indices()[0].numpy(), sp_.indices()[1].numpy(), sp_.values().numpy()
    c = int(labels.max()) + 1
    oh = np.eye(c, dtype=np.float32)[labels]
    x = z["out_features"]
    close(O.plot_edge_homophily(row, col, val, oh), z["out_edge_homo"])
    close(O.plot_node_homophily(row, col, labels, n), z["out_node_homo"])
    close(O.plot_class_homophily(row, col, val, labels, n), z["out_class_homo"])
    nzr, nzc = a.nonzero().T.numpy()
    close(O.plot_compat_matrix(nzr, nzc, labels), z["out_compat"])
    close(O.plot_similarity(oh, row, col, val, n, oh), z["out_soft_las"], rtol=0, atol=1e-6)
    close(O.plot_similarity(oh, row, col, val, n, oh, hard=1), z["out_hard_las"], rtol=0, atol=1e-6)
    close(O.plot_similarity(oh, row, col, val, n, oh, idx_train=z["in_idx_train"]), z["out_soft_las_idx"], rtol=0, atol=1e-6)
    xn = x / np.linalg.norm(x, axis=1, keepdims=True)
    close(O.plot_similarity(xn, row, col, val, n, oh, NTK=True), z["out_soft_las_ntk"], rtol=0, atol=1e-6)
    p, p_bar, pc = O.class_distribution(row, col, labels, n)
    close(p, z["out_p"]); close(p_bar, z["out_p_bar"]); close(pc, z["out_pc"])
    s2 = np.float32(np.sum(p_bar.astype(np.float32) ** 2, dtype=np.float32))
    close((O.plot_edge_homophily(row, col, val, oh) - s2) / (1 - s2), z["out_adj_homo"], rtol=1e-4)
    close(O.label_informativeness(row, col, labels, n), z["out_label_info"], rtol=1e-4, atol=1e-5)
    close(O.generalized_edge_homophily(row, col, x, n), z["out_gen_edge_homo"], rtol=1e-4)
    close(O.normalize_tensor(z["in_features_raw"]).numpy(), x, rtol=1e-6)   # preprocess_features == row normalisation


def test_coalesce_and_counts_small():
    # duplicates are summed, order is row-major (torch .coalesce())
    row = np.array([2, 0, 2, 1, 0]); col = np.array([1, 2, 1, 1, 0]); val = np.array([1, 2, 3, 4, 5], np.float32)
    r, c, v = O.coalesce(row, col, val, 3)
    assert r.tolist() == [0, 0, 1, 2] and c.tolist() == [0, 2, 1, 1] and v.tolist() == [5, 2, 4, 4]
    s = O.structure_counts(r, c, np.array([0, 1, 1]), 3)
    assert s["deg_all"].tolist() == [2, 1, 1] and s["deg_nsl"].tolist() == [1, 0, 1]
    assert s["match_all"] == 3 and s["match_nsl"].tolist() == [0, 0, 1]
    assert s["hist"].tolist() == [[0, 1], [0, 1]]
    assert O.csr_from_coo(r, 3).tolist() == [0, 2, 3, 4]


def test_empty_graph():
    e = np.zeros(0, np.int64)
    s = O.structure_counts(e, e, np.array([0, 1, 0]), 3)
    assert s["nnz"] == 0 and s["hist"].sum() == 0
    assert np.isnan(O.edge_homophily(e, e, np.array([0, 1, 0])))
    y = O.spmm(e, e, np.zeros(0, np.float32), 3, np.ones((3, 4), np.float32))
    assert y.shape == (3, 4) and not y.any()


# ---------------------------------------------------------------------------
# util_funcs.py normalisers next to the path: normalize :29, preprocess_features :39, normalize_adj :429,
# dataset_edge_balance :439 (fixture tests/golden/util_norm.npz, make_golden.case_util_norm)
# ---------------------------------------------------------------------------
def _util_norm_inputs(z):
    import scipy.sparse as sp
    n, d = int(z["in_n"]), int(z["in_feat_dim"])
    a = sp.coo_matrix((z["in_val"], (z["in_row"], z["in_col"])), shape=(n, n)).tocsr()
    f = sp.coo_matrix((z["in_feat_val"], (z["in_feat_row"], z["in_feat_col"])), shape=(n, d)).tocsr()
    return n, a, f


def test_util_normalisers():
    import scipy.sparse as sp
    z = G.load("util_norm")
    n, a, f = _util_norm_inputs(z)
    for got, tag in ((O.normalize(a), "normalize"), (O.preprocess_features(f), "preprocess")):
        m = sp.csr_matrix(got)
        m.sort_indices()
        assert np.array_equal(m.indptr, z[f"out_{tag}_indptr"]) and np.array_equal(m.indices, z[f"out_{tag}_indices"])
        close(m.data, z[f"out_{tag}_data"], rtol=1e-12, atol=0)
    dense = torch.from_numpy(a.toarray()).float()
    close(O.normalize(dense + torch.eye(n)), z["out_normalize_dense"], rtol=1e-7, atol=0)
    close(O.preprocess_features(torch.from_numpy(f.toarray()).float()), z["out_preprocess_dense"], rtol=1e-7, atol=0)
    coo = a.tocoo()
    r, c, v = O.normalize_adj(coo.row, coo.col, coo.data, n)
    ref = sp.csr_matrix((z["out_normalize_adj_data"], z["out_normalize_adj_indices"], z["out_normalize_adj_indptr"]),
                        shape=(n, n)).tocoo()
    assert np.array_equal(r, ref.row) and np.array_equal(c, ref.col)
    close(v, ref.data, rtol=1e-12, atol=0)
    nodes, bal = O.dataset_edge_balance(coo.row, coo.col, coo.data, z["in_labels"], n)
    close(nodes, z["out_balance_nodes"], rtol=0, atol=0)
    close(bal, z["out_balance"], rtol=1e-12, atol=0)
    _, bal_b = O.dataset_edge_balance(coo.row, coo.col, np.ones_like(coo.data), z["in_labels"], n)
    close(bal_b, z["out_balance_binary"], rtol=0, atol=0)
