#!/bin/bash
# round-2 run A: full GPU test suite, default bench (c3), c2 bench, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench_c3.log 2>&1
echo "bench rc=$?" >> gpurun_out/r2a_bench_c3.log
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2a_launches.csv \
  python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-consensus --no-cd > gpurun_out/r2a_ncu_bench.log 2>&1
tail -3 gpurun_out/r2a_pytest.log
tail -c 600 gpurun_out/r2a_bench_c3.log
