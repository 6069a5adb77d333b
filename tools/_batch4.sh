# scratch helper for gpurun (single GPU): drop-in script test, pipeline forms, e2e A/B of the row-chunk phases
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "unchanged or pipeline_host" --durations=5 > gpurun_out/r5_tests.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r5_tests.log
timeout 300 python tools/run_reference_script.py --impl wdgh --runs cora:node_homo:0 cora:class_homo:0 texas:edge_homo:1 2>&1 | tail -1 | cut -c1-900
for cr in 1 8 4; do
WDGH_E2E_COLROWS=$cr timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r5_bench_colrows$cr.json 2> gpurun_out/r5_bench_colrows$cr.err; echo "bench colrows=$cr rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r5_bench_colrows$cr.json") if l.startswith("{")][-1])
print("colrows $cr: N1 ms", round(d["ms_per_step"],2), "e2e ms", round(d["e2e"]["ms_per_step"],1), "GE/s", round(d["e2e"]["value"],4), "metrics_only ms", round(d["e2e"]["metrics_only"]["ms_per_step"],1), "chk", d["e2e"]["y_host_checksum"], "traffic_frac", d["roofline"].get("traffic_frac"))
PY
done
