# usage (on an N-GPU box): bash tools/multi_gpu_check.sh N [check]  -- final-code bench (default settings, traced stages) at world size N, optional partition check
N=$1
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
if [ "$2" = "check" ]; then
timeout 900 $TR --nproc-per-node $N --master-port 29541 tools/check_partitions.py --nodes 1000003 2000000 > gpurun_out/multi_check_n$N.log 2>&1; echo "check n$N rc=$?"; grep '"check"' gpurun_out/multi_check_n$N.log | tail -1 | cut -c1-900
fi
WDGH_STAGE_TIMES=1 timeout 900 $TR --nproc-per-node $N --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/multi_bench_n$N.json 2> gpurun_out/multi_bench_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/multi_bench_n$N.json") if l.startswith("{")][-1])
print("N", d["n_gpus"], "ms", round(d["ms_per_step"],2), "GE/s", round(d["value"],2), "verify", d["verify"]["ok"], d["verify"]["max_abs_err_over_max_abs"], "frac", round(d["roofline"]["frac"],3), d["config"]["partition"][:60])
print("stages", d["stage_ms"])
PY
