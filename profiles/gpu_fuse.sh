#!/bin/bash
# fused double pass: correctness tests first, then cfg3 with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_structure.py -m gpu -q --tb=short -k "fused or default_path or structure or canvas or slabs_draw" > gpurun_out/fuse_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/fuse_tests.log
for f in 1 0; do
  FDTD2D_FUSE=$f timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/bench_fuse$f.json 2> gpurun_out/bench_fuse$f.err; echo "fuse=$f rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_fuse$f.json') if l.startswith('{')][-1])
print('fuse=$f value',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'launches',d['gpu_launches'],'passes',d['tile_kernel_launches'],d['clocks'])
PY
done
