#!/bin/bash
# A/B of an env knob on one box: tools/ab_env.sh VAR A B [rounds]
VAR=$1; A=$2; B=$3; ROUNDS=${4:-2}
for r in $(seq 1 $ROUNDS); do
  for v in $A $B; do
    env $VAR=$v python bench.py --steps 20 --warmup 3 --quick 2>/dev/null | python -c "import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$VAR=$v', 'ms_per_step %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'])"
  done
done
