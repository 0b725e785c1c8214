#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2g_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2g_pytest.log
tail -14 gpurun_out/r2g_pytest.log | cut -c1-200
timeout 600 python tools/probe_consensus.py > gpurun_out/r2g_consensus.log 2>&1
tail -4 gpurun_out/r2g_consensus.log | cut -c1-700
B="python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-consensus --no-cd"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3_kernel -s 40 -c 2 -o gpurun_out/r2g_ncu_gemm $B > gpurun_out/r2g_ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:update_kernel -s 40 -c 2 -o gpurun_out/r2g_ncu_update $B > gpurun_out/r2g_ncu_update.log 2>&1
CNMF_FUSE_W=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_fused_w -s 20 -c 1 -o gpurun_out/r2g_ncu_fused $B > gpurun_out/r2g_ncu_fused.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
