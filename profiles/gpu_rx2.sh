#!/bin/bash
# resident tests, cfg5 timing (old kernel = cfg 0, packed kernel = cfg 5), one ncu --set full capture of the packed kernel
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_resident.py -m gpu -q --maxfail=12 --tb=line 2>&1 | tail -18
CFGS=0,5 TRIMS=-1 timeout 100 python profiles/resident_diag.py
FDTD2D_RESIDENT_CFG=5 B=132 N=200 timeout 100 python profiles/resident_prof.py
FDTD2D_RESIDENT_CFG=5 B=132 N=200 timeout 300 ncu --set full --clock-control none --import-source on -k regex:grid_resident_x2 -c 1 -f -o gpurun_out/r2_resident_x2 python profiles/resident_prof.py > gpurun_out/r2_resident_x2.ncu.log 2>&1; echo "ncu rc=$?"
