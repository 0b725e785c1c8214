#!/bin/bash
# strong scaling on one box: the c3 table at N = (all GPUs) and N = 1
NG=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
for n in $NG 1; do
if [ "$n" = 1 ]; then L="python bench.py"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py"; fi
timeout 900 $L --gpus $n --steps 3 --warmup 3 --no-cpu-baseline --no-cd > gpurun_out/r2v_bench_n$n.log 2>&1
grep '^{' gpurun_out/r2v_bench_n$n.log | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('N=%d value %.1f e2e %.1f ms/step %.1f with_consensus %s' % (d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], json.dumps(d['with_consensus']['ms_per_step']) if d.get('with_consensus') else None))
"
done
