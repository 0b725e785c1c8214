#!/bin/bash
# multi-GPU run: N = number of visible GPUs
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
if [ "$N" -ge 2 ] && [ -z "$SKIP_PYTEST" ]; then
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k two_gpu > gpurun_out/${TAG:-r2i}_pytest_2gpu.log 2>&1
tail -3 gpurun_out/${TAG:-r2i}_pytest_2gpu.log | cut -c1-300
fi
for n in ${BENCH_NS:-2 4 8}; do
if [ "$N" -ge "$n" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 2 $BENCH_ARGS > gpurun_out/${TAG:-r2i}_bench_n$n.log 2>&1
grep '^{' gpurun_out/${TAG:-r2i}_bench_n$n.log | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('N=%d value %.1f e2e %.1f ms/step %.1f with_consensus %s' % (d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], json.dumps(d['with_consensus']['ms_per_step']) if d.get('with_consensus') else None))
"
fi
done
