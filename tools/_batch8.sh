cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "spmm or pipeline_host or phased or block or width or propert" > gpurun_out/r8_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r8_tests.log | cut -c1-300
for cr in 1 4 8; do
WDGH_E2E_COLROWS=$cr timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r8_bench_colrows$cr.json 2> gpurun_out/r8_bench_colrows$cr.err; echo "bench colrows=$cr rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r8_bench_colrows$cr.json") if l.startswith("{")][-1])
print("colrows $cr: N1 ms", round(d["ms_per_step"],2), "e2e ms", round(d["e2e"]["ms_per_step"],1), "GE/s", round(d["e2e"]["value"],4), "metrics_only ms", round(d["e2e"]["metrics_only"]["ms_per_step"],1), "chk", d["e2e"]["y_host_checksum"])
PY
done
timeout 300 python tools/colblock_probe.py 2>&1 | tail -4
