#!/bin/bash
# A/B of two builds of the library on ONE GPU box: alternating bench.py runs (ms per step of the graph-replayed train step).
#   tools/ab_step.sh fpl-plus_b200/libfplplus_b200_prev.so [rounds]
# The other build is made from another commit:  git archive <commit> | tar -x -C /tmp/prev && python /tmp/prev/fpl-plus_b200/build.py
OTHER=$1; ROUNDS=${2:-2}
for r in $(seq 1 $ROUNDS); do
  for which in new old; do
    if [ $which = old ]; then export FPL_LIB_AB=$OTHER; else unset FPL_LIB_AB; fi
    python bench.py --steps 20 --warmup 3 --quick 2>/dev/null | python -c "import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$which', 'ms_per_step %.3f' % d['ms_per_step'])"
  done
done
