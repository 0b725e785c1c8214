#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2y_pytest.log
tail -25 gpurun_out/r2y_pytest.log | cut -c1-400
