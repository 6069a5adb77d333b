# usage: bash tools/_multi.sh N   -- partition checks + benches at world size N (scratch helper for gpurun)
N=$1
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node $N --master-port 29541 tools/check_partitions.py --nodes 1000003 2000000 > gpurun_out/r2_check_n$N.log 2>&1; echo "check n$N rc=$?"; grep '"check"' gpurun_out/r2_check_n$N.log | tail -1 | cut -c1-1500
for ov in on off; do
  WDGH_STAGE_TIMES=1 timeout 900 $TR --nproc-per-node $N --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --overlap $ov > gpurun_out/r2_bench_n${N}_$ov.json 2> gpurun_out/r2_bench_n${N}_$ov.err; echo "bench n$N overlap=$ov rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench_n${N}_$ov.json") if l.startswith("{")][-1])
    print("N", d["n_gpus"], "ms", round(d["ms_per_step"],2), "GE/s", round(d["value"],2), "verify", d["verify"]["ok"], d["verify"]["max_abs_err_over_max_abs"], "nvlink", d["nvlink"])
    print("stages", d["stage_ms"])
except Exception as e:
    print("no line", e)
PY
  grep -i "error\|Traceback" gpurun_out/r2_bench_n${N}_$ov.err | head -5
done
