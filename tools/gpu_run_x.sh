#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2x_smoke.log 2>&1; tail -3 gpurun_out/r2x_smoke.log
timeout 120 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-cd --no-consensus > gpurun_out/r2x_bench_c2.log 2>&1; tail -1 gpurun_out/r2x_bench_c2.log | cut -c1-200
