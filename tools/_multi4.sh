set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,name --format=csv | head -6
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29531 tools/check_partitions.py --nodes 1000003 2000000 > gpurun_out/r2_check_n2.log 2>&1; echo "check n2 rc=$?"; grep '"check"' gpurun_out/r2_check_n2.log | tail -1; tail -5 gpurun_out/r2_check_n2.log | cut -c1-400
timeout 600 $TR --nproc-per-node 4 --master-port 29532 tools/check_partitions.py --nodes 1000003 2000000 > gpurun_out/r2_check_n4.log 2>&1; echo "check n4 rc=$?"; grep '"check"' gpurun_out/r2_check_n4.log | tail -1; tail -5 gpurun_out/r2_check_n4.log | cut -c1-400
WDGH_STAGE_TIMES=1 timeout 900 $TR --nproc-per-node 4 --master-port 29533 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "bench n4 rc=$?"; cat gpurun_out/r2_bench_n4.json | cut -c1-3000; grep stages gpurun_out/r2_bench_n4.err | head -4; tail -3 gpurun_out/r2_bench_n4.err | cut -c1-400
WDGH_STAGE_TIMES=1 timeout 900 $TR --nproc-per-node 2 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"; cat gpurun_out/r2_bench_n2.json | cut -c1-3000; grep stages gpurun_out/r2_bench_n2.err | head -2; tail -3 gpurun_out/r2_bench_n2.err | cut -c1-400
python -m pytest tests/test_gpu_parity.py -q -x -k "plot_variants" 2>&1 | tail -40 > gpurun_out/r2_plot_test.log; grep -E "Error|assert|passed|failed" gpurun_out/r2_plot_test.log | head -20
