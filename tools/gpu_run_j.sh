#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/probe_stalls.py > gpurun_out/r2j_stalls_blas4.log 2>&1
tail -8 gpurun_out/r2j_stalls_blas4.log | cut -c1-420
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-cd > gpurun_out/r2j_bench_c3b.log 2>&1
grep '^{' gpurun_out/r2j_bench_c3b.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], json.dumps(d['with_consensus'])[:600])"
