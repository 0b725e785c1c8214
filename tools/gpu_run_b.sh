#!/bin/bash
# round-2 run B: CTA-pair GEMM bring-up (tests, A/B timing), consensus phase probe
mkdir -p gpurun_out
export CUDA_LAUNCH_BLOCKING=0
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gemm" > gpurun_out/r2b_pytest_gemm.log 2>&1
echo "rc=$?" >> gpurun_out/r2b_pytest_gemm.log
tail -3 gpurun_out/r2b_pytest_gemm.log
CNMF_GEMM_PAIR=1 timeout 300 python tools/probe_gemm.py > gpurun_out/r2b_gemm_pair.log 2>&1
CNMF_GEMM_PAIR=0 timeout 300 python tools/probe_gemm.py > gpurun_out/r2b_gemm_1cta.log 2>&1
tail -1 gpurun_out/r2b_gemm_pair.log; tail -1 gpurun_out/r2b_gemm_1cta.log
CNMF_GEMM_PAIR=1 timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-consensus --no-cd > gpurun_out/r2b_bench_pair.log 2>&1
CNMF_GEMM_PAIR=0 timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-consensus --no-cd > gpurun_out/r2b_bench_1cta.log 2>&1
tail -c 1500 gpurun_out/r2b_bench_pair.log | head -c 700
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1
tail -3 gpurun_out/r2b_pytest.log
timeout 600 python tools/probe_consensus.py > gpurun_out/r2b_consensus.log 2>&1
tail -5 gpurun_out/r2b_consensus.log
