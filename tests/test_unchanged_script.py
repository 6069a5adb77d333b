"""Drop-in check: the reference's own `homophily_tests.py`, UNCHANGED, run against the wdgh_b200 mirrors
(tools/run_reference_script.py) prints the values it prints with the reference's own modules.

Golden values: tests/golden/script_homophily_tests.json, made by tests/golden/make_script_golden.py from the unmodified
reference on CPU.  The script and the small dataset files its loader reads travel in oracle/_ref (oracle/build_ref.py,
git-ignored); without them the tests skip.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "script_homophily_tests.json")))["results"]
NODES = {"cora": 2708, "citeseer": 3327, "texas": 183, "cornell": 183, "wisconsin": 251}


def run_script(impl, runs, env=None):
    cmd = [sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py"), "--impl", impl, "--runs"] + list(runs)
    out = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=1500, env=env)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{"impl"')]
    assert out.returncode == 0 and lines, out.stdout[-1500:] + out.stderr[-3000:]
    rep = json.loads(lines[-1])
    if "unavailable" in rep:
        pytest.skip(rep["unavailable"])
    return rep


def test_reference_arm_reproduces_the_golden_values():
    """The runner itself: reference modules on CPU give the committed values bit for bit (no GPU involved)."""
    runs = ["cora:node_homo:0", "texas:agg_homo_soft:1", "wisconsin:adj_homo:1", "cornell:node_hom_generalized:1"]
    rep = run_script("reference", runs, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    for r in runs:
        assert rep["results"][r] == GOLD[r], (r, rep["results"][r], GOLD[r])


@pytest.mark.gpu
@pytest.mark.timeout(1800)
def test_unchanged_homophily_tests_script_on_the_cuda_path():
    """All 35 (dataset, metric, normalisation) command lines of the golden file through the mirrors on the B200."""
    rep = run_script("wdgh", list(GOLD))
    assert rep["gpu_launches"] > 0
    # the functions of utils/util_funcs.py the script imports are the mirrors, the loader is the reference's
    assert {"normalize_tensor", "row_normalized_adjacency", "sys_normalized_adjacency",
            "sparse_mx_to_torch_sparse_tensor"} <= set(rep["util_funcs_replaced_by_mirror"])
    bad = []
    for run, want in GOLD.items():
        dataset, metric, _ = run.split(":")
        got = rep["results"][run]
        if metric in ("agg_homo_soft", "agg_homo_hard"):   # 2 * mean(indicator) - 1: one node on a float tie may flip
            tol = 2 * 1.5 / NODES[dataset]
        elif metric == "label_info":
            tol = 1e-5 + 1e-4 * abs(want)
        elif metric in ("node_homo", "edge_homo"):
            tol = 1e-6 * max(abs(want), 1e-3)
        else:
            tol = 1e-4 * max(abs(want), 1e-3)
        if not abs(got - want) <= tol:
            bad.append((run, got, want, tol))
    assert not bad, bad
