#!/bin/bash
# N-GPU bench (gpurun --gpus N): the default command the driver runs (peer links), then the NCCL exchange for comparison
# (second argument "p2p" skips the NCCL run)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps ${STEPS:-5} --warmup ${WARMUP:-3} > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err; echo "p2p rc=$?"; tail -3 gpurun_out/scale_$N.err
[ "${2:-nccl}" = nccl ] && timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 5 --warmup 3 --exchange nccl --no-extras --no-e2e > gpurun_out/scale_${N}_nccl.json 2> gpurun_out/scale_${N}_nccl.err; echo "nccl rc=$?"
python - <<PY
import json
for f in ('scale_$N.json','scale_${N}_nccl.json'):
    try:
        d=json.loads([l for l in open('gpurun_out/'+f) if l.startswith('{')][-1])
    except Exception as e:
        print(f,'no line',e); continue
    print(f,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],2),'e2e',d['e2e'] and round(d['e2e']['value'],1),d.get('slab_parity') and d['slab_parity']['result'],'strong',d.get('strong_scaling') and (round(d['strong_scaling']['value'],1),round(d['strong_scaling']['ms_per_step'],3)))
PY
