# scratch helper for gpurun (single GPU): targeted tests, full N=1 bench line, ncu launch list + full captures, linkx
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "pipeline_host or data_synthesis_sweep or powerlaw or structure or edge_cases or smoke" --durations=10 > gpurun_out/r4_tests.log 2>&1; echo "pytest rc=$?"; tail -18 gpurun_out/r4_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r4_bench_n1.json 2> gpurun_out/r4_bench_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r4_bench_n1.json") if l.startswith("{")][-1])
print("N1 ms", d["ms_per_step"], "GE/s", d["value"], "frac", d["roofline"]["frac"], "labels", d["roofline_labels"]["kernel_ms"], "gram", d["roofline_gram"]["frac"])
print("e2e", d["e2e"]); print("cpu", d["cpu_baseline"]); print("clocks", d["clocks"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmm_|structure_|degree_scale|plan_|labels_to|gram_" -c 120 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_r02_bench.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/launches_r02.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spmm_rowgroup|structure_stream|gram_tcgen05" -c 6 -f -o gpurun_out/prof_r02c python tools/profile_kernels.py --nodes 16000000 > gpurun_out/prof_r02c.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/prof_r02c.log
timeout 400 python bench.py --workload linkx --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r4_linkx.json 2> gpurun_out/r4_linkx.err; echo "linkx rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r4_linkx.json") if l.startswith("{")][-1])
print("linkx s", d["value"], "kr_check", d["kr_check"])
PY
