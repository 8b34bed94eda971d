#!/bin/bash
# usage: e2e_repeat.sh N [env assignments...] : run the cfg2 bench N times, report crashes
N=$1; shift
ok=0; bad=0
for i in $(seq $N); do
  if env "$@" python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu > /tmp/e2e_rep.out 2> /tmp/e2e_rep.err; then ok=$((ok+1)); else bad=$((bad+1)); grep -m1 "Fdtd2dError\|Error" /tmp/e2e_rep.err | cut -c1-200; fi
done
echo "$* : ok=$ok bad=$bad"
