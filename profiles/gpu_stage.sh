#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "staged" > gpurun_out/stage_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/stage_tests.log
for cfg in "0 0" "4 1" "5 1"; do
  set -- $cfg
  FDTD2D_STAGE=$1 FDTD2D_AUTO_K12=$2 timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/bench_stage$1.json 2> gpurun_out/bench_stage$1.err; echo "stage=$1 rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_stage$1.json') if l.startswith('{')][-1])
print('stage=$1 value',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'k',d['config']['k_temporal'],'launches',d['gpu_launches'],d['clocks'])
PY
done
