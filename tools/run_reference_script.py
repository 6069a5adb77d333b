#!/usr/bin/env python
"""Run the reference's own `homophily_tests.py`, UNCHANGED, against the wdgh_b200 mirrors (the drop-in claim of
BASELINE.json north_star: "the Python call signatures stay intact, so homophily_tests.py ... run against it unchanged").

    python tools/run_reference_script.py --impl wdgh      --runs cora:node_homo:0 texas:agg_homo_soft:1 ...
    python tools/run_reference_script.py --impl reference --runs ...          # the reference's own modules, CPU

One run = `<dataset>:<homophily_metric>:<symmetric>`, i.e. the script's `--dataset_name / --homophily_metric /
--symmetric` arguments.  The script file is executed as it is with `runpy` (module name `__main__`, its own argparse);
the only thing that differs between the two arms is what `import utils.homophily_metrics` / `import utils.util_funcs`
resolve to:

  * `reference`: the reference's modules;
  * `wdgh`: `utils.homophily_metrics` IS `wdgh_b200.homophily_metrics`; `utils.util_funcs` is the reference's module --
    its dataset loaders (`full_load_data_large`: file IO + networkx, out of scope of the hot path) stay -- with every
    function `wdgh_b200.util_funcs` mirrors replaced by the mirror.  This is the binding INTEGRATION.md describes.

In both arms the SCRIPT and the reference modules it imports (callers whose source file lies in the reference tree, nobody
else) see `torch.cuda.is_available() == False`, i.e. the script
follows its CPU path -- BASELINE.json configs[0], "homophily_tests.py on CPU (reference path)" -- and hands host tensors
to whatever `utils.*` is bound; the mirrors move them to the B200 themselves (the library is initialised before the
patch).  The script's own CUDA branch cannot be used for a comparison: it is broken upstream (edge_homo indexes a CPU
`torch.eye` with CUDA labels, homophily_tests.py:112; uf.py:18-20 imports torch_geometric names when CUDA is visible).

The reference tree is /root/reference when present, else the copy under oracle/_ref (oracle/build_ref.py); the script
runs with the tree as working directory because its loaders open `data/...` relatively.  Third-party imports that are
absent from this image and not on the path are stubbed (oracle/ref_shim.py); `to_scipy_sparse_matrix`, which the
script's large-dataset branch calls, gets a two-line implementation here.

Prints one JSON line: {"impl": ..., "results": {"<run>": value}}.  Test infrastructure, not product code.
"""
import argparse
import contextlib
import io
import json
import os
import re
import runpy
import shutil
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "when-do-gnns-help_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def _to_scipy_sparse_matrix(edge_index, edge_attr=None, num_nodes=None):
    import numpy as np
    import scipy.sparse as sp
    row, col = edge_index.cpu().numpy()
    n = int(max(row.max(), col.max())) + 1 if num_nodes is None else num_nodes
    val = np.ones(row.shape[0]) if edge_attr is None else edge_attr.cpu().numpy()
    return sp.coo_matrix((val, (row, col)), shape=(n, n))


def bind(impl, tree):
    """Make `utils.*` importable from `tree` according to `impl`; returns nothing, mutates sys.modules."""
    import torch
    from oracle import ref_shim
    ref_shim.install_stubs()
    conv = types.ModuleType("torch_geometric.utils.convert")
    conv.to_scipy_sparse_matrix = _to_scipy_sparse_matrix
    sys.modules["torch_geometric.utils.convert"] = conv
    sys.modules["torch_geometric.utils"].convert = conv
    for name in [m for m in sys.modules if m == "utils" or m.startswith("utils.")]:
        del sys.modules[name]
    sys.path.insert(0, tree)
    if impl == "wdgh":
        import wdgh_b200
        wdgh_b200._lib.require_device()        # needs the real torch.cuda.is_available(); cached afterwards
    # from here on the script and the reference modules (every caller whose source file lies in `tree`) take their CPU
    # path; torch's own internals and the mirrors keep seeing the real answer
    real = torch.cuda.is_available

    def is_available():
        return False if sys._getframe(1).f_code.co_filename.startswith(tree) else real()
    torch.cuda.is_available = is_available
    import utils.homophily_metrics  # noqa: F401
    import utils.util_funcs as ref_uf
    if impl == "reference":
        return
    uf = types.ModuleType("utils.util_funcs")  # the loaders stay, every mirrored function is replaced
    uf.__dict__.update({k: v for k, v in ref_uf.__dict__.items() if not k.startswith("__")})
    mirrored = [k for k, v in wdgh_b200.util_funcs.__dict__.items()
                if callable(v) and not k.startswith("_") and k in ref_uf.__dict__ and
                getattr(v, "__module__", "") == wdgh_b200.util_funcs.__name__]
    for k in mirrored:
        setattr(uf, k, getattr(wdgh_b200.util_funcs, k))
    sys.modules["utils.util_funcs"] = uf
    sys.modules["utils"].util_funcs = uf
    sys.modules["utils.homophily_metrics"] = wdgh_b200.homophily_metrics
    sys.modules["utils"].homophily_metrics = wdgh_b200.homophily_metrics
    bind.mirrored = sorted(mirrored)


def run_once(tree, dataset, metric, symmetric, seed):
    import random

    import numpy as np
    import torch
    random.seed(seed), np.random.seed(seed), torch.manual_seed(seed)
    argv = sys.argv
    sys.argv = ["homophily_tests.py", "--dataset_name", dataset, "--homophily_metric", metric, "--symmetric", symmetric]
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            runpy.run_path(os.path.join(tree, "homophily_tests.py"), run_name="__main__")
    finally:
        sys.argv = argv
    m = re.search(r"The Homophily level of given dataset \S+ is (.+) using metric (\S+)", buf.getvalue())
    if not m or m.group(2) != metric:
        raise RuntimeError("unexpected script output: " + buf.getvalue()[-400:])
    text = m.group(1).strip()
    t = re.fullmatch(r"tensor\(([^,)]+).*\)", text)       # torch scalars print as tensor(0.81, device='cuda:0')
    return float(t.group(1) if t else text)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["wdgh", "reference"], required=True)
    ap.add_argument("--runs", nargs="+", required=True, help="dataset:metric:symmetric")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import ref_shim
    src = ref_shim.default_root()
    if src is None or not os.path.exists(os.path.join(src, "homophily_tests.py")):
        print(json.dumps({"impl": a.impl, "unavailable": "no reference tree with homophily_tests.py "
                                                         "(python oracle/build_ref.py in the build container)"}))
        return 0
    # the script writes data/acmgcn_features/ next to itself: run it from a scratch copy of the (small) tree so that
    # neither /root/reference (read-only) nor oracle/_ref is written to
    scratch = tempfile.mkdtemp(prefix="wdgh_ref_tree_")
    tree = os.path.join(scratch, "tree")
    keep = ("homophily_tests.py", "utils", "data", "new_data")
    os.makedirs(tree)
    for name in keep:
        s = os.path.join(src, name)
        if os.path.isdir(s):
            shutil.copytree(s, os.path.join(tree, name), ignore=shutil.ignore_patterns(
                "facebook100", "twitch", "splits", "ind.pubmed*", "film", "squirrel", "__pycache__"))
        elif os.path.exists(s):
            shutil.copyfile(s, os.path.join(tree, name))
    cwd = os.getcwd()
    os.chdir(tree)
    out = {}
    try:
        bind(a.impl, tree)
        for run in a.runs:
            dataset, metric, symmetric = run.split(":")
            out[run] = run_once(tree, dataset, metric, symmetric, a.seed)
    finally:
        os.chdir(cwd)
        shutil.rmtree(scratch, ignore_errors=True)
    line = {"impl": a.impl, "script": "homophily_tests.py (unchanged)", "seed": a.seed, "results": out}
    if a.impl == "wdgh":
        import wdgh_b200
        line["util_funcs_replaced_by_mirror"] = bind.mirrored
        line["gpu_launches"] = int(wdgh_b200.launch_count())
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
