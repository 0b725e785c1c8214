#!/bin/bash
# where a small shard's time goes + launch lists (shard of 8, and the default bench command)
mkdir -p gpurun_out
timeout 600 python tools/probe_shard.py 8 3 > gpurun_out/r2n_shard.log 2>&1; tail -2 gpurun_out/r2n_shard.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2n_launches_shard8.csv python tools/probe_shard.py 8 1 > gpurun_out/r2n_ncu_shard.log 2>&1; tail -2 gpurun_out/r2n_ncu_shard.log | cut -c1-300
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r2n_launches_bench_c3.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-consensus --no-cd > gpurun_out/r2n_ncu_bench.log 2>&1; tail -1 gpurun_out/r2n_ncu_bench.log | cut -c1-200
ls -la gpurun_out/r2n_*
