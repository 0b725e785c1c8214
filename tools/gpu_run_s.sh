#!/bin/bash
# final-state: ncu --set full of the leaner update kernel (c2, same command as run G), then the default bench lines
mkdir -p gpurun_out
B="python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-consensus --no-cd"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:update_kernel -s 40 -c 2 -o gpurun_out/r2s_ncu_update $B > gpurun_out/r2s_ncu_update.log 2>&1
ls -la gpurun_out/r2s_ncu_update.ncu-rep
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2s_bench_c3.log 2>&1
tail -1 gpurun_out/r2s_bench_c3.log | cut -c1-300
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/r2s_bench_c2.log 2>&1
tail -1 gpurun_out/r2s_bench_c2.log | cut -c1-300
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2s_bench_ref.log 2>&1
tail -1 gpurun_out/r2s_bench_ref.log | cut -c1-400
