#!/bin/bash
mkdir -p gpurun_out
for f in 1 0; do
CNMF_FUSE_W=$f timeout 300 python bench.py --workload c2 --steps 5 --warmup 2 --no-cpu-baseline --no-consensus --no-cd > gpurun_out/r2e_bench_c2_fuse$f.log 2>&1
CNMF_FUSE_W=$f timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-consensus --no-cd > gpurun_out/r2e_bench_c3_fuse$f.log 2>&1
done
for f in c2_fuse1 c2_fuse0 c3_fuse1 c3_fuse0; do echo $f; grep -o '"value": [0-9.]*' gpurun_out/r2e_bench_$f.log | head -1; grep -o '"n_iter": {[^}]*}' gpurun_out/r2e_bench_$f.log; grep -o '[0-9]* launches, [0-9.]* ms of [0-9.]* ms timed' gpurun_out/r2e_bench_$f.log; tail -2 gpurun_out/r2e_bench_$f.log | grep -v "^{" | cut -c1-300; done
PROBE_REPS=2 CNMF_FUSE_W=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/probe_fused.py > gpurun_out/r2e_memcheck.log 2>&1
tail -4 gpurun_out/r2e_memcheck.log | cut -c1-300
