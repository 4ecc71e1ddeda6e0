#!/bin/bash
# usage: scratch/ab.sh "<env ids>" <lib1> <lib2> ...
ids="$1"; shift
for lib in "$@"; do
  for id in $ids; do
    MGB_LIB=$lib timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs --env-id $id 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib'.split('/')[-1], d['config']['workload'][:32], 'value=%.3e frac=%.3f step_mode=%.3e' % (d['value'], d['roofline']['frac'], d['step_mode_env_steps_per_s_per_gpu']))"
  done
done
