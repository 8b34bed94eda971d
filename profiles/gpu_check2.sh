#!/bin/bash
# Multi-GPU check (gpurun --gpus N): slab tests across the GPUs, the multi-process checks, a short bench at N ranks.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "== slab tests"; timeout 900 python -m pytest tests/test_gpu_slabs.py -m gpu -q --tb=short > gpurun_out/slabs_$N.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/slabs_$N.log
echo "== bench N=$N"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err; echo "rc=$?"; tail -c 2500 gpurun_out/bench_$N.json; tail -5 gpurun_out/bench_$N.err
echo "== bench N=$N nccl"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 --exchange nccl --no-extras --no-cpu > gpurun_out/bench_${N}_nccl.json 2> gpurun_out/bench_${N}_nccl.err; echo "rc=$?"; head -c 600 gpurun_out/bench_${N}_nccl.json; tail -3 gpurun_out/bench_${N}_nccl.err
