#!/bin/bash
# usage: scale_run.sh WORKLOAD "N list" [extra bench args]; writes gpurun_out/scale_<workload>_<N>.json
WL=$1; NS=$2; shift 2
for N in $NS; do
  OUT=gpurun_out/scale_${WL}_${N}.json
  if [ "$N" = 1 ]; then
    python bench.py --workload $WL --gpus 1 "$@" > $OUT 2> gpurun_out/scale_${WL}_${N}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
      bench.py --workload $WL --gpus $N "$@" > $OUT 2> gpurun_out/scale_${WL}_${N}.err
  fi
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$OUT") if l.startswith("{")][-1]
    print("$WL N=$N value=%.1f Gcell/s ms_per_step=%.2f e2e=%s" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value")))
except Exception as e:
    print("$WL N=$N FAILED", e); print(open("gpurun_out/scale_${WL}_${N}.err").read()[-1500:])
PY
done
