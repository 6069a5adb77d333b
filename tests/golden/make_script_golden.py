#!/usr/bin/env python
"""Golden outputs of the reference's own `homophily_tests.py` (run unchanged, reference modules, CPU) for the drop-in
test: tests/golden/script_homophily_tests.json.

    python tests/golden/make_script_golden.py        # build container (needs /root/reference or oracle/_ref)

Every run is `<dataset>:<homophily_metric>:<symmetric>` = the script's command line.  `class_homo` is not listed: the
reference's script hands a sparse TENSOR to `our_measure`, which expects an edge index, and crashes upstream
(hm.py:36 `NotImplementedError: aten::ne.Tensor ... SparseCPU`); the KR metrics are covered prediction by prediction
elsewhere (tests/test_gpu_parity.py::kr_contract).
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DATASETS = ["cora", "citeseer", "texas", "cornell", "wisconsin"]
METRICS = ["node_homo", "edge_homo", "node_hom_generalized", "agg_homo_soft", "agg_homo_hard", "adj_homo", "label_info"]


def runs():
    out = []
    for i, d in enumerate(DATASETS):
        for j, m in enumerate(METRICS):
            out.append(f"{d}:{m}:{(i + j) % 2}")
    return out


if __name__ == "__main__":
    cmd = [sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py"), "--impl", "reference", "--runs"] + runs()
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith('{"impl"')]
    assert res.returncode == 0 and lines, res.stdout[-2000:] + res.stderr[-2000:]
    rep = json.loads(lines[-1])
    assert "results" in rep and len(rep["results"]) == len(runs()), rep
    with open(os.path.join(HERE, "script_homophily_tests.json"), "w") as f:
        json.dump({"script": "homophily_tests.py, unchanged, reference modules on CPU", "seed": rep["seed"],
                   "results": rep["results"]}, f, indent=1)
    print(json.dumps(rep["results"], indent=1))
