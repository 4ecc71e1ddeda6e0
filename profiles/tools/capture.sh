#!/bin/bash
# One `ncu --set full` capture of a persistent rollout launch (T = 32, 2^20 envs) of one env id, on the GPU box:
#   gpurun -- 'bash profiles/tools/capture.sh <env-id> <tag> [lib.so]'
# The bench command is run WITHOUT ncu first (must exit 0), then under ncu; the report lands in gpurun_out/<tag>.ncu-rep.
# Launch 0 of k_rollout is the reset, launches 1.. are rollouts: the third rollout is captured.
id="$1"; tag="$2"; lib="${3:-}"; skip="${4:-3}"      # skip=8: a single-step launch (reset, 4 rollouts, then the step-mode leg)
[ -n "$lib" ] && export MGB_LIB="$lib"
python -c "from gym_minigrid_b200 import _lib; print(_lib.load().mgb_version().decode())" > gpurun_out/${tag}_stamp.txt    # source revision + -D switches of the library
cmd="python bench.py --env-id $id --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs"
$cmd > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_rollout --launch-skip $skip --launch-count 1 -f -o gpurun_out/$tag $cmd > gpurun_out/${tag}_ncu.log 2>&1
ls -la gpurun_out/$tag.ncu-rep
