#!/bin/bash
# A/B of two source trees on ONE GPU box: alternating bench.py runs (ms per step of the graph-replayed train step).
#   tools/ab_tree.sh tools/_build/prev_tree [rounds]
# The other tree is a built checkout of another commit:  git archive <commit> | tar -x -C <dir> && python <dir>/fpl-plus_b200/build.py
OTHER=$1; ROUNDS=${2:-2}; HERE=$(pwd)
for r in $(seq 1 $ROUNDS); do
  for which in new old; do
    if [ $which = old ]; then cd $OTHER; else cd $HERE; fi
    python bench.py --steps 20 --warmup 3 --quick 2>/dev/null | python -c "import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$which', 'ms_per_step %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'])"
    cd $HERE
  done
done
