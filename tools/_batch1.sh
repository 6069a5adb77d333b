# scratch helper for gpurun (single GPU): PCIe probe, label-kernel occupancy A/B, column-block probe, GPU tests, linkx, sweep
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pcie_probe tools/pcie_probe.cu && timeout 300 /tmp/pcie_probe > gpurun_out/r3_pcie_probe.log 2>&1; echo "pcie rc=$?"; cat gpurun_out/r3_pcie_probe.log
for mb in 3 4 5; do
WDGH_LABEL_MINB=$mb timeout 300 python - <<PY 2>&1 | tail -2
import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "when-do-gnns-help_b200")
import torch, bench
import wdgh_b200 as W
from wdgh_b200 import graph as G
dev = torch.device("cuda:0")
n = 50_000_000
rowptr, col, x, labels = bench.gen_rows(0, n, n, 20.0, 10, 0.3, 128, dev, want_x=False)
g = G.CSRGraph(rowptr, col, None, n); _ = g.plan
r = bench.time_label_pass(G, g, labels, 10, int(col.shape[0]), n, 6545.0, "x")
print("label pass MINB=$mb", round(r["kernel_ms"], 3), "ms")
PY
done
timeout 600 python tools/colblock_probe.py --nodes 50000000 > gpurun_out/r3_colblock.log 2>&1; echo "colblock rc=$?"; cat gpurun_out/r3_colblock.log | tail -8
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3_gputest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r3_gputest.log
timeout 600 python bench.py --workload linkx --steps 2 --warmup 1 > gpurun_out/r3_linkx.json 2> gpurun_out/r3_linkx.err; echo "linkx rc=$?"; cut -c1-1500 gpurun_out/r3_linkx.json
timeout 600 python tools/sweep_synthesis.py --limit 60 --out gpurun_out/r3_sweep60.json > gpurun_out/r3_sweep.log 2>&1; echo "sweep rc=$?"; tail -2 gpurun_out/r3_sweep.log | cut -c1-1500
