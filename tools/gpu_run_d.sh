#!/bin/bash
mkdir -p gpurun_out
CNMF_FUSE_W=0 timeout 300 python tools/probe_fused.py > gpurun_out/r2d_probe_unfused.log 2>&1
CNMF_FUSE_W=1 timeout 300 python tools/probe_fused.py > gpurun_out/r2d_probe_fused.log 2>&1
tail -2 gpurun_out/r2d_probe_unfused.log | cut -c1-1200; tail -2 gpurun_out/r2d_probe_fused.log | cut -c1-1200
PROBE_REPS=2 CNMF_FUSE_W=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/probe_fused.py > gpurun_out/r2d_memcheck.log 2>&1
grep -v "^$" gpurun_out/r2d_memcheck.log | head -60 | cut -c1-220
