"""TEST INFRASTRUCTURE (not product code): import the UNMODIFIED reference modules of the hot path.

    hm, uf, hp = load("/root/reference")            # build container: the reference where it lies
    hm, uf, hp = load(default_root())               # oracle/_ref on the GPU box (see oracle/build_ref.py)

`utils/homophily_metrics.py`, `utils/util_funcs.py` and `utils/homophily_plot.py` are imported as they are.  Their
hard dependencies that are absent from this image and are NOT on the graph-statistics path (torch_scatter, dgl,
torch_geometric, ogb, google_drive_downloader) are replaced by stub modules so that the import succeeds;
`torch_scatter.scatter_add` is the only stub that is ever called and forwards to `Tensor.scatter_add_`.
The reference picks `cuda:0` at import time whenever torch sees a GPU (hm.py:17-21, uf.py:21-26); this loader is for
the CPU arm (`bench.py --impl reference`, golden fixtures), so `torch.cuda.is_available` reports False while the
modules are imported and they bind `device = cpu`.

Only tests/, tests/golden/make_golden.py, __graft_entry__.smoke() and bench.py's CPU arm may import this module.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
BUILT = os.path.join(HERE, "_ref")          # build-time copy made by oracle/build_ref.py (git-ignored)
SOURCE = "/root/reference"                  # exists in the build container only


def default_root():
    """The reference tree to import: the real one when present, else the copy that travelled with the repo."""
    if os.path.isdir(os.path.join(SOURCE, "utils")):
        return SOURCE
    if os.path.isdir(os.path.join(BUILT, "utils")):
        return BUILT
    return None


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _scatter_add(src, index, dim=-1, out=None, dim_size=None):
    assert out is not None
    return out.scatter_add_(dim, index, src)


def install_stubs():
    for name in ("torch_scatter", "dgl", "torch_geometric", "google_drive_downloader", "ogb"):
        try:
            importlib.import_module(name)
            continue
        except Exception:
            pass
        if name == "torch_scatter":
            _stub(name, scatter_add=_scatter_add)
        elif name == "torch_geometric":
            tg = _stub(name)
            tg.utils = _stub("torch_geometric.utils", to_undirected=None)
        elif name == "google_drive_downloader":
            _stub(name, GoogleDriveDownloader=object)
        elif name == "ogb":
            _stub(name)
            _stub("ogb.nodeproppred", NodePropPredDataset=object)
        else:
            _stub(name)


def load(root=None):
    """-> (homophily_metrics, util_funcs, homophily_plot) of the reference at `root`, bound to the CPU."""
    import torch

    root = root or default_root()
    if root is None:
        raise ImportError("no reference tree: neither /root/reference nor oracle/_ref (python oracle/build_ref.py)")
    install_stubs()
    for name in [m for m in sys.modules if m == "utils" or m.startswith("utils.")]:
        del sys.modules[name]
    sys.path.insert(0, root)
    real = torch.cuda.is_available
    torch.cuda.is_available = lambda: False
    try:
        hm = importlib.import_module("utils.homophily_metrics")
        uf = importlib.import_module("utils.util_funcs")
        hp = importlib.import_module("utils.homophily_plot")
    finally:
        torch.cuda.is_available = real
        sys.path.remove(root)
    return hm, uf, hp
