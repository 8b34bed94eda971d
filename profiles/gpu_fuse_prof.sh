#!/bin/bash
mkdir -p gpurun_out
python profiles/fuse_prof.py > gpurun_out/fuse_plain.log 2>&1 || exit 1
tail -1 gpurun_out/fuse_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_fuse_launches.csv python profiles/fuse_prof.py > /dev/null 2>&1
grep -c strip_wave gpurun_out/r2_fuse_launches.csv
ncu --set full --clock-control none --import-source on -k regex:strip_wave_x2 -s 2 -c 2 -o gpurun_out/r2_fuse_full python profiles/fuse_prof.py > gpurun_out/ncu_fuse.log 2>&1; echo "ncu rc=$?"
