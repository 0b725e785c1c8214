#!/bin/bash
# A/B of the MU update-kernel variants (CNMF_UPD_VARIANT) on c2 and c3
mkdir -p gpurun_out
for v in 0 1 2 3; do
  for w in c2 c3; do
    CNMF_UPD_VARIANT=$v timeout 600 python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline --no-consensus --no-cd > gpurun_out/r2m_v${v}_$w.log 2>&1
    python - <<PY
import json
l=[x for x in open("gpurun_out/r2m_v${v}_$w.log") if x.startswith('{"metric')]
if l:
    d=json.loads(l[-1]); print("variant $v $w", round(d["value"],1), "ms", round(d["ms_per_step"],1), "upd", d.get("roofline_update",{}).get("note","")[ -140:-60], "n_iter", d["n_iter"])
else: print("variant $v $w FAILED")
PY
  done
done
CNMF_UPD_VARIANT=2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "factorize or golden or baseline" > gpurun_out/r2m_pytest_v2.log 2>&1; tail -3 gpurun_out/r2m_pytest_v2.log
