#!/bin/bash
# 2-GPU box: the whole GPU suite (2-rank torchrun test included) + the sharded bench at N = 2 and N = 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_pytest_2gpu.log 2>&1
echo "rc=$?" >> gpurun_out/r2u_pytest_2gpu.log
tail -4 gpurun_out/r2u_pytest_2gpu.log | cut -c1-300
for n in 2 1; do
if [ "$n" = 1 ]; then L="python bench.py"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py"; fi
timeout 900 $L --gpus $n --steps 3 --warmup 3 --no-cpu-baseline --no-cd > gpurun_out/r2u_bench_n$n.log 2>&1
grep '^{' gpurun_out/r2u_bench_n$n.log | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('N=%d value %.1f e2e %.1f ms/step %.1f with_consensus %s' % (d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], json.dumps(d['with_consensus']['ms_per_step']) if d.get('with_consensus') else None))
"
done
