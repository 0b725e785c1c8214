#!/bin/bash
# final state of the round: GPU suite + the default bench command + c2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2w_pytest.log
tail -4 gpurun_out/r2w_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2w_bench_c3.log 2>&1
tail -1 gpurun_out/r2w_bench_c3.log | cut -c1-260
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_bench_c2.log 2>&1
tail -1 gpurun_out/r2w_bench_c2.log | cut -c1-260
