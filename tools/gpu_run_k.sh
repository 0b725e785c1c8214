#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tests/gpu_probe.py big > gpurun_out/r2k_big.log 2>&1
tail -3 gpurun_out/r2k_big.log | cut -c1-400
timeout 300 python tests/gpu_probe.py cd > gpurun_out/r2k_cd.log 2>&1
tail -3 gpurun_out/r2k_cd.log | cut -c1-300
timeout 300 python tests/gpu_probe.py kl > gpurun_out/r2k_kl.log 2>&1
tail -3 gpurun_out/r2k_kl.log | cut -c1-300
timeout 600 python tools/probe_consensus.py > gpurun_out/r2k_consensus.log 2>&1
tail -4 gpurun_out/r2k_consensus.log | cut -c1-600
