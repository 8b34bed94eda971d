#!/bin/bash
# ncu --set full of the stepping kernel for the workloads behind profiles/traffic.json (one launch each)
mkdir -p gpurun_out
run() { # name, kernel regex, env...
  name=$1; rx=$2; shift 2
  env "$@" python profiles/pass_prof.py > gpurun_out/$name.plain.log 2>&1 || { echo "$name plain run failed"; tail -3 gpurun_out/$name.plain.log; return; }
  env "$@" ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -o gpurun_out/$name python profiles/pass_prof.py > gpurun_out/$name.ncu.log 2>&1; echo "$name ncu rc=$?"
}
run r2_cfg2_wave strip_wave_x2 R=4096
run r2_fp64_wave8 strip_wave_kernel R=8192 DTYPE=f64
run r2_cfg3_wave strip_wave_x2 R=16384
env R=4096 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_cfg2_launches.csv python profiles/pass_prof.py > /dev/null 2>&1
env R=8192 DTYPE=f64 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_fp64_launches.csv python profiles/pass_prof.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 400 --log-file gpurun_out/r2_default_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > /dev/null 2>&1
ls -la gpurun_out | tail -12
