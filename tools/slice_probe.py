"""Single-GPU probe of the short-row case of the 2-D partition: one row slice (rows of one rank, columns of one
column group, ~5 stored entries per row) aggregated as raw partial sums, next to the full 1-D shard of the same rows.

    python tools/slice_probe.py [--world 8] [--nodes 50000000]

Prints one JSON line per case: kernel time and algorithmic GB/s."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "when-do-gnns-help_b200"))

import torch  # noqa: E402

import bench  # noqa: E402
import wdgh_b200 as W  # noqa: E402
from wdgh_b200 import graph as G  # noqa: E402
from wdgh_b200.sharded import Grid2D  # noqa: E402


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--nodes", type=int, default=50_000_000)
    ap.add_argument("--dim", type=int, default=128)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    n, d, world = args.nodes, args.dim, args.world
    grid = Grid2D(n, world, 2)
    blk = grid.part.block
    r0, r1 = grid.part.bounds(0)
    rowptr, col, _, _ = bench.gen_rows(r0, r1, n, 20, 10, 0.3, d, dev, want_x=False)
    x = torch.empty((world * blk, d), dtype=torch.float32, device=dev).normal_()
    g1 = G.CSRGraph(rowptr, col, None, r1 - r0, row_offset=r0, n_global=n)
    dinv, _, code = g1.degree_scale(W.NORM_SYM, True)
    dinv_f = torch.ones(world * blk, dtype=torch.float32, device=dev)
    code_f = torch.zeros(world * blk, dtype=torch.uint8, device=dev)
    dinv_f[r0:r1], code_f[r0:r1] = dinv, code
    y = torch.empty((blk, d), dtype=torch.float32, device=dev)
    out = {}
    ms = timed(lambda: G.spmm(g1, x, W.NORM_SYM, True, out=y[:g1.n], dinv=dinv_f, deg_code=code_f))
    by = g1.nnz * (8 + 4 * d) + g1.n * (12 + 8 * d)
    out["shard_1d"] = {"rows": g1.n, "entries": g1.nnz, "ms": round(ms, 3), "GBps": round(by / ms / 1e6, 1)}
    frp, fcol = grid.filter_slice(rowptr, col, 0)
    gs = G.CSRGraph(frp, fcol, None, r1 - r0, row_offset=r0, n_global=n)
    skip = G.heavy_flags(gs) if gs.n_chunks else None
    ms = timed(lambda: G.spmm_ranged(gs, gs.rowptr[:-1], gs.rowptr[1:], x, y, W.NORM_SYM, True, dinv_f, code_f, skip,
                                     False, False, True))
    by = gs.nnz * (8 + 4 * d) + gs.n * (8 + 4 * d)
    out["slice_2d"] = {"rows": gs.n, "entries": gs.nnz, "ms": round(ms, 3), "GBps": round(by / ms / 1e6, 1)}
    # the 1-D shard as raw partial sums (no self loop / scale): isolates the epilogue
    g1r = G.CSRGraph(rowptr, col, None, r1 - r0, row_offset=r0, n_global=n)
    skip1 = G.heavy_flags(g1r) if g1r.n_chunks else None
    ms = timed(lambda: G.spmm_ranged(g1r, g1r.rowptr[:-1], g1r.rowptr[1:], x, y, W.NORM_SYM, True, dinv_f, code_f, skip1,
                                     False, False, True))
    by = g1r.nnz * (8 + 4 * d) + g1r.n * (8 + 4 * d)
    out["shard_1d_raw"] = {"ms": round(ms, 3), "GBps": round(by / ms / 1e6, 1)}
    # the slice's entry count with every row exactly K entries long, uniform columns of the same two shards
    for K in (5, 20):
        rows = r1 - r0
        rp = torch.arange(rows + 1, device=dev, dtype=torch.int64) * K
        cc = (torch.randint(0, 2, (rows * K,), device=dev) * (grid.pc * blk) +
              torch.randint(0, blk, (rows * K,), device=dev)).to(torch.int32)
        gu = G.CSRGraph(rp, cc, None, rows, row_offset=r0, n_global=n)
        ms = timed(lambda: G.spmm_ranged(gu, gu.rowptr[:-1], gu.rowptr[1:], x, y, W.NORM_SYM, True, dinv_f, code_f, None,
                                         False, False, True))
        by = gu.nnz * (8 + 4 * d) + gu.n * (8 + 4 * d)
        out[f"uniform_k{K}"] = {"ms": round(ms, 3), "GBps": round(by / ms / 1e6, 1)}
        del gu, rp, cc
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
