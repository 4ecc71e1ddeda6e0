#!/bin/bash
# A/B of library variants on the GPU box.  Build a variant next to the product library with
#   python -m gym_minigrid_b200.build -DMGB_<SWITCH>=<v> --out=build/ab/<name>.so
# (build/ is git-ignored but travels with gpurun; gpurun_out/ does not), then
#   gpurun -- 'bash profiles/tools/ab.sh "<env ids>" $PWD/gym_minigrid_b200/libmgb200.so $PWD/build/ab/<name>.so ...'
# MGB_LIB selects the library for one process; run the parity tests on a variant the same way.
ids="$1"; shift
for lib in "$@"; do
  for id in $ids; do
    MGB_LIB=$lib timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs --env-id $id 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib'.split('/')[-1], d['config']['workload'][:32], 'value=%.3e frac=%.3f step_mode=%.3e' % (d['value'], d['roofline']['frac'], d['step_mode_env_steps_per_s_per_gpu']))"
  done
done
