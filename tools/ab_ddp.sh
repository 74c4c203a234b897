#!/bin/bash
# A/B of the gradient all-reduce modes on N GPUs of one box: tools/ab_ddp.sh N [rounds]
N=${1:-2}; ROUNDS=${2:-2}
for r in $(seq 1 $ROUNDS); do
  for v in deferred overlapped; do
    FPL_GRAD_ALLREDUCE=$v python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
      bench.py --gpus $N --steps 20 --warmup 3 --quick 2>/dev/null | python -c "import sys, json; d = json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1]); print('$v', 'n_gpus', d['n_gpus'], 'ms_per_step %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], d['config'].get('grad_allreduce'))"
  done
done
