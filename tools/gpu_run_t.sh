#!/bin/bash
# validation of the leaner update kernel (component loop, packed emission, symmetric Gram): tests + quick benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2t_pytest.log
tail -5 gpurun_out/r2t_pytest.log | cut -c1-300
for w in c3 c2; do
timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-cd --no-consensus > gpurun_out/r2t_bench_$w.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/r2t_bench_$w.log") if x.startswith('{"metric')]
if l:
    d=json.loads(l[-1]); print("$w", round(d["value"],1), "ms", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],1), "gemm frac", round(d["roofline"]["frac"],3), "upd frac", round(d["roofline_update"]["frac"],3), d["roofline_update"]["note"][-150:-70], d["n_iter"], d["clocks"])
else: print("$w FAILED"); print(open("gpurun_out/r2t_bench_$w.log").read()[-1500:])
PY
done
