#!/usr/bin/env python
"""Multi-GPU cross-check of the partitioned pipelines ON THE SAME GRAPH, in one launch (any even world size >= 2):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 \
        tools/check_partitions.py [--nodes 4000000 1000003]

Every rank builds its row shard of the bench generator's graph and runs, in this order,

    1d-phased  peer-mapped shards pulled by the copy engines, one `y +=` phase per arriving shard
    2d         2 row groups x N/2 column groups: partner shard pulled, foreign row slices stored into their owners'
               memory by the aggregation kernel, own slice finished by its last phase (reduce + self loop + scale)
    2d-overlap (N >= 4) the same with the own slice's partner columns aggregated NEXT TO the foreign slices on a
               second stream and a pure streaming reduction as last launch (opt-in: measured slower at N = 8)
    1d-plain   NCCL all-gather of the features, then ONE aggregation launch (the N = 1 kernel on a row shard)

and compares each result (Y rows of the rank, all-reduced counters, node sum) with 1d-plain, which runs LAST so that
no earlier pipeline can inherit a correctly filled buffer from it through the caching allocator; the peer-mapped
feature buffers start out as NaN (sharded._symmetric_features), so a block that is read without having been filled
shows.  Rank 0 prints one JSON line and exits non-zero on a mismatch.  Y tolerance 5e-6 of max |Y| (float32 sums in a
different association); counters must be identical.  (tests/test_gpu_multi.py runs this under pytest when >= 2 GPUs
are visible.)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "when-do-gnns-help_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def check(args, n, world, rank, device, bench, W, G, CudaShardedStats, Cuda2DShardedStats, Grid2D, RowPartition):
    C, d = args.classes, args.dim
    args.nodes = n                     # bench.make_slice_graphs reads the node count from args
    part = RowPartition(n, world)
    r0, r1 = part.bounds(rank)
    rowptr, col, x_local, labels_local = bench.gen_rows(r0, r1, n, args.avg_degree, C, args.homophily, d, device)
    g = G.CSRGraph(rowptr, col, None, r1 - r0, row_offset=r0, n_global=n)

    def run(pipe):
        for _ in range(2):                      # the second step also exercises buffer reuse across steps
            y, counters, node_sum = pipe.step(W.NORM_SYM, True)
        torch.cuda.synchronize()
        return y.clone(), counters.clone(), node_sum.clone()

    results = {}
    results["1d-phased"] = run(CudaShardedStats(g, part, rank, x_local, labels_local, C, phased=True))
    torch.cuda.empty_cache()
    if world % 2 == 0:
        grid2 = Grid2D(n, world, 2)
        slices = bench.make_slice_graphs(G, grid2, rank, rowptr, col, args, device)
        results["2d"] = run(Cuda2DShardedStats(grid2, rank, slices, g, x_local, labels_local, C, overlap=False))
        torch.cuda.empty_cache()
        if world >= 4:   # own slice aggregated next to the foreign slices on a second stream + streaming reduction
            results["2d-overlap"] = run(Cuda2DShardedStats(grid2, rank, slices, g, x_local, labels_local, C, overlap=True))
        del slices
        torch.cuda.empty_cache()
    results["1d-plain"] = run(CudaShardedStats(g, part, rank, x_local, labels_local, C, phased=False))

    y_ref, cnt_ref, ns_ref = results["1d-plain"]
    scale = torch.tensor([float(y_ref.abs().max())], dtype=torch.float64, device=device)
    dist.all_reduce(scale, op=dist.ReduceOp.MAX)
    report, ok = {}, True
    for name, (y, cnt, ns) in results.items():
        err = torch.tensor([float((y - y_ref).abs().max())], dtype=torch.float64, device=device)
        bad = torch.tensor([int(not torch.equal(cnt, cnt_ref))], dtype=torch.int64, device=device)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        dist.all_reduce(bad)
        rel = float(err.item() / scale.item())
        ns_rel = abs(float(ns[0]) - float(ns_ref[0])) / max(abs(float(ns_ref[0])), 1e-30)
        report[name] = {"y_max_err_rel": rel, "counters_equal": int(bad.item()) == 0, "node_sum_rel": ns_rel}
        ok &= rel <= 5e-6 and int(bad.item()) == 0 and ns_rel <= 1e-12   # NaN fails the first comparison
    return {"nodes": n, "results": report}, ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, nargs="+", default=[4_000_000, 1_000_003],
                    help="graph sizes to check, one after the other (a size that is not a multiple of the world size "
                         "exercises the padded last shard)")
    ap.add_argument("--avg-degree", type=float, default=20.0)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--classes", type=int, default=10)
    ap.add_argument("--homophily", type=float, default=0.3)
    args = ap.parse_args()

    import bench
    import wdgh_b200 as W
    from wdgh_b200 import graph as G
    from wdgh_b200.sharded import Cuda2DShardedStats, CudaShardedStats, Grid2D, RowPartition

    world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    W._lib.require_device()
    dist.init_process_group("nccl", device_id=device)
    reports, all_ok = [], True
    for n in list(args.nodes):
        rep, ok = check(args, n, world, rank, device, bench, W, G, CudaShardedStats, Cuda2DShardedStats, Grid2D,
                        RowPartition)
        reports.append(rep)
        all_ok &= ok
        torch.cuda.empty_cache()
    ok = all_ok
    if rank == 0:
        print(json.dumps({"check": "partitions", "n_gpus": world, "dim": args.dim, "ok": bool(ok), "graphs": reports}),
              flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
