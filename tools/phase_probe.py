"""Probe (not product code): device-side cost of the ranged aggregation phases the host pipeline is made of -- R row
chunks (column ranges of every row) at width 128 and on a 64-column strided view -- with everything resident, no PCIe.

    python tools/phase_probe.py --nodes 50000000
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=50_000_000)
    ap.add_argument("--widths", type=int, nargs="+", default=[128, 64])
    ap.add_argument("--chunks", type=int, nargs="+", default=[1, 2, 4, 8, 16])
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--ctas", type=int, nargs="+", default=[0], help="CTAs per SM of the ranged launches (0 = default 32)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    n, d = a.nodes, 128
    rowptr, col, x, labels = bench.gen_rows(0, n, n, 20.0, 10, 0.3, d, dev)
    g = G.CSRGraph(rowptr, col, None, n)
    _ = g.plan
    dinv, _, code = g.degree_scale(W.NORM_SYM, True)
    skip = G.heavy_flags(g) if g.n_chunks else None
    y = torch.empty_like(x)
    for w in a.widths:
        xv, yv = x[:, :w], y[:, :w]
        for R, ctas in [(R, c) for R in a.chunks for c in a.ctas]:
            blk = (n + R - 1) // R
            seg = G.column_segments(g, [min(r * blk, n) for r in range(R)] + [n])
            times = []
            for rep in range(a.reps):
                evs = [torch.cuda.Event(enable_timing=True) for _ in range(R + 1)]
                evs[0].record()
                for r in range(R):
                    last = r == R - 1
                    G.spmm_ranged(g, seg[r], seg[r + 1], xv, yv, W.NORM_SYM, True, dinv, code, skip, accumulate=r > 0,
                                  finalize=last, run_split_rows=last, ctas_per_sm=ctas)
                    evs[r + 1].record()
                torch.cuda.synchronize()
                times = [evs[r].elapsed_time(evs[r + 1]) for r in range(R)]
            print(f"width {w:3d}  R={R:2d} ctas/SM={ctas or 32:2d}: total {sum(times):8.2f} ms   per phase " + " ".join(f"{t:6.1f}" for t in times), flush=True)


if __name__ == "__main__":
    main()
