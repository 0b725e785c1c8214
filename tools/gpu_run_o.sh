#!/bin/bash
mkdir -p gpurun_out
SHARD_RANKS=0,1,2,3,4,5,6,7 PROFILE=1 timeout 600 python tools/probe_shard.py 8 3 > gpurun_out/r2o_shards_prof.log 2>&1; cat gpurun_out/r2o_shards_prof.log | cut -c1-330
SHARD_RANKS=0,1,2,3,4,5,6,7 PROFILE=0 timeout 600 python tools/probe_shard.py 8 3 > gpurun_out/r2o_shards_noprof.log 2>&1; cat gpurun_out/r2o_shards_noprof.log | cut -c1-200
