#!/bin/bash
# round-2 run C: fused W-half epilogue bring-up
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "factorize or pipeline or worker or k_selection" > gpurun_out/r2c_pytest_fused.log 2>&1
echo "rc=$?" >> gpurun_out/r2c_pytest_fused.log
tail -25 gpurun_out/r2c_pytest_fused.log | cut -c1-250
CNMF_FUSE_W=1 timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-consensus --no-cd > gpurun_out/r2c_bench_fused.log 2>&1
CNMF_FUSE_W=0 timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-consensus --no-cd > gpurun_out/r2c_bench_unfused.log 2>&1
CNMF_FUSE_W=1 timeout 300 python bench.py --workload c2 --steps 5 --warmup 2 --no-cpu-baseline --no-consensus --no-cd > gpurun_out/r2c_bench_c2_fused.log 2>&1
CNMF_FUSE_W=0 timeout 300 python bench.py --workload c2 --steps 5 --warmup 2 --no-cpu-baseline --no-consensus --no-cd > gpurun_out/r2c_bench_c2_unfused.log 2>&1
for f in fused unfused c2_fused c2_unfused; do echo $f; grep -o '"value": [0-9.]*' gpurun_out/r2c_bench_$f.log | head -1; tail -3 gpurun_out/r2c_bench_$f.log | grep -v "^{" | tail -2; done
