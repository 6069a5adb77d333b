#!/usr/bin/env python
"""TEST INFRASTRUCTURE: recipe for oracle/_ref -- the reference's own CPU implementation of the hot path, so that
`bench.py --impl reference` and the `cpu_baseline` leg can time THE REFERENCE (kind = "reference") on the GPU box,
where /root/reference does not exist.

    python oracle/build_ref.py        # build container only; __graft_entry__.build() runs it when /root/reference exists

The reference is pure Python with no setup.py / pyproject.toml (`pip install /root/reference` has nothing to
install), so "building" it means taking the modules of the path, byte for byte, from where they lie:

    /root/reference/utils/{homophily_metrics,util_funcs,homophily_plot,datasets}.py  ->  oracle/_ref/utils/

For the drop-in demonstration (tools/run_reference_script.py: the reference's own `homophily_tests.py`, UNCHANGED, run
against the wdgh_b200 mirrors) it also takes the script itself and the small dataset files its loader reads:

    /root/reference/homophily_tests.py, data/ind.{cora,citeseer}.*, new_data/{texas,cornell,wisconsin}/out1_*.txt

It also packs the 580 `data_synthesis/{800,4000}/<h>/adj_<h>_<s>.pt` graphs the reference ships (2000 nodes, 5 classes;
the inputs of synthetic_plot.py:60-110) into ONE compressed archive, oracle/_ref/data_synthesis.npz, so that the sweep
runner (tools/sweep_synthesis.py) can replay the whole sweep on the GPU box.

oracle/_ref/ is listed in .gitignore (the reference sources and data never enter the history) but not in
.gpurunignore, so it travels to the GPU box like the built .so.  Nothing is patched: the stubs for the absent
third-party imports live in oracle/ref_shim.py, which is ours.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ["utils/homophily_metrics.py", "utils/util_funcs.py", "utils/homophily_plot.py", "utils/datasets.py",
         "homophily_tests.py"]
DATA_GLOBS = ["data/ind.cora.*", "data/ind.citeseer.*", "new_data/texas/out1_*.txt", "new_data/cornell/out1_*.txt",
              "new_data/wisconsin/out1_*.txt"]


def pack_data_synthesis():
    """adj_<h>_<s>.pt / label_<h>_<s>.pt -> {"<size>/<h>/<s>/edges": int16 [2, nnz], ".../labels": int8 [n]}."""
    import glob
    import warnings

    import numpy as np
    import torch

    dst = os.path.join(DST, "data_synthesis.npz")
    files = sorted(glob.glob(os.path.join(SRC, "data_synthesis", "*", "*", "adj_*.pt")))
    if os.path.exists(dst) and all(os.path.getmtime(dst) >= os.path.getmtime(f) for f in files):
        return dst
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for f in files:
            size, h = f.split(os.sep)[-3:-1]
            s = os.path.basename(f)[:-3].split("_")[-1]
            adj = torch.load(f, weights_only=False).coalesce()
            lab = torch.load(f.replace("adj_", "label_"), weights_only=False).to_dense()
            assert adj.shape[0] < 32768 and bool((adj.values() == 1).all())
            out[f"{size}/{h}/{s}/edges"] = adj.indices().numpy().astype(np.int16)
            out[f"{size}/{h}/{s}/labels"] = lab.argmax(1).numpy().astype(np.int8)
    np.savez_compressed(dst, **out)
    return dst


def build():
    if not os.path.isdir(os.path.join(SRC, "utils")):
        return None     # GPU box / no reference: use what travelled with the repo
    os.makedirs(os.path.join(DST, "utils"), exist_ok=True)
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    import glob
    for pat in DATA_GLOBS:      # inputs of the reference's dataset loaders (uf.py:60-100, 288-330), as shipped
        for src in sorted(glob.glob(os.path.join(SRC, pat))):
            dst = os.path.join(DST, os.path.relpath(src, SRC))
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
                shutil.copyfile(src, dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1)
    pack_data_synthesis()
    return DST


if __name__ == "__main__":
    print(build())
