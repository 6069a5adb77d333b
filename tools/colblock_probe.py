"""Probe (not product code): cost of aggregating a COLUMN BLOCK of the feature matrix in place -- x[:, k0:k0+w] of a
row-major [n][d] matrix, ldx = ldy = d -- next to the full-width launch.  Feeds the column-block form of the e2e
pipeline (H2D of block k+1 and D2H of block k-1 run under the aggregation of block k).

    python tools/colblock_probe.py --nodes 50000000
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "when-do-gnns-help_b200"))

import torch  # noqa: E402

import bench  # noqa: E402
import wdgh_b200 as W  # noqa: E402
from wdgh_b200 import graph as G  # noqa: E402
from wdgh_b200._lib import check, lib, ptr, stream_ptr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=50_000_000)
    ap.add_argument("--dim", type=int, default=128)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    n, d = a.nodes, a.dim
    rowptr, col, x, labels = bench.gen_rows(0, n, n, 20.0, 10, 0.3, d, dev)
    g = G.CSRGraph(rowptr, col, None, n)
    plan, plan_host = g.plan
    dinv, _, code = g.degree_scale(W.NORM_SYM, True)
    y = torch.empty_like(x)
    y_ref = torch.empty_like(x)
    G.spmm(g, x, W.NORM_SYM, True, out=y_ref, dinv=dinv, deg_code=code)
    torch.cuda.synchronize()

    def run(w, k0):
        xv, yv = x[:, k0:k0 + w], y[:, k0:k0 + w]
        partial = G._partial_scratch(g, w)
        check(lib.wdgh_spmm_csr(ptr(g.rowptr), ptr(g.col), None, g.n, xv.data_ptr(), w, d, yv.data_ptr(), d,
                                W.NORM_SYM, 1, ptr(dinv), ptr(code), ptr(plan), plan_host, ptr(partial), 0,
                                stream_ptr()), "wdgh_spmm_csr")

    for w in (128, 64, 32, 16):
        if w > d:
            continue
        y.zero_()
        for k0 in range(0, d, w):
            run(w, k0)
        torch.cuda.synchronize()
        same = bool(torch.equal(y, y_ref))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 3
        for _ in range(reps):
            for k0 in range(0, d, w):
                run(w, k0)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"column blocks of {w:3d}: {d // w} launches, {ms:8.2f} ms per full pass ({ms / (d // w):7.2f} ms per block), "
              f"bit-identical to the one-launch result: {same}", flush=True)


if __name__ == "__main__":
    main()
