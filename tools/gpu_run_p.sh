#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2p_pytest.log
tail -5 gpurun_out/r2p_pytest.log | cut -c1-300
timeout 600 python tools/probe_consensus.py > gpurun_out/r2p_consensus.log 2>&1
tail -4 gpurun_out/r2p_consensus.log | cut -c1-900
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2p_bench_c3.log 2>&1
tail -1 gpurun_out/r2p_bench_c3.log | cut -c1-1800
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/r2p_bench_c2.log 2>&1
tail -1 gpurun_out/r2p_bench_c2.log | cut -c1-600
