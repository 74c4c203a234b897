#!/bin/bash
# bench.py --quick under several values of one env knob, alternating: tools/ab_env3.sh VAR rounds v1 v2 v3 ...
VAR=$1; ROUNDS=$2; shift 2
for r in $(seq 1 $ROUNDS); do
  for v in "$@"; do
    env $VAR=$v python bench.py --steps 20 --warmup 3 --quick 2>/dev/null | python -c "import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$VAR=$v', 'ms_per_step %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'])"
  done
done
