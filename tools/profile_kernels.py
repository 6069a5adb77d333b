"""Launch the kernels that `profiles/` documents a fixed number of times, for ncu captures (see profiles/summary_r02.md):

    ncu --set full --clock-control none --import-source on -k regex:"structure_stream|gram_tcgen05" -c 4 \
        -o gpurun_out/prof_r02 python tools/profile_kernels.py --nodes 16000000

Each kernel family runs twice: label pass (labels_to_u8 + structure_stream_kernel + split rows), aggregation
(spmm_rowgroup_kernel + split rows), Gram (gram_split_kernel + gram_tcgen05_kernel at m = 8192, d = 1024).
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
    ap.add_argument("--nodes", type=int, default=16_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--classes", type=int, default=10)
    ap.add_argument("--skip", default="", help="comma list of families to skip: labels,spmm,gram")
    a = ap.parse_args()
    skip = set(a.skip.split(","))
    dev = torch.device("cuda:0")
    rowptr, col, x, labels = bench.gen_rows(0, a.nodes, a.nodes, 20.0, a.classes, 0.3, a.dim, dev)
    g = G.CSRGraph(rowptr, col, None, a.nodes)
    _ = g.plan
    torch.cuda.synchronize()
    if "labels" not in skip:
        scratch = None
        for _ in range(2):
            scratch = G.structure_counts_raw(g, labels, a.classes, scratch)
        torch.cuda.synchronize()
    if "spmm" not in skip:
        y = torch.empty_like(x)
        for _ in range(2):
            G.spmm(g, x, W.NORM_SYM, True, out=y)
        torch.cuda.synchronize()
    if "gram" not in skip:
        z = torch.randn(8192, 1024, device=dev)
        for faithful in (False, True):
            G.gram(z, use_tensor_cores=True, faithful=faithful)
        torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
