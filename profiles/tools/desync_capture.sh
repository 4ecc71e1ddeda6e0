#!/bin/bash
# `ncu --set full` capture of ONE rollout launch with spread-out episode ends (profiles/tools/desync.py state), summarised
# on the box:  gpurun -- 'bash profiles/tools/desync_capture.sh <env-id> <tag> <mangled-kernel-substring> [launch]'
id="$1"; tag="$2"; kern="$3"; skip="${4:-24}"
python -c "from gym_minigrid_b200 import _lib; print(_lib.load().mgb_version().decode())" > gpurun_out/${tag}_stamp.txt
timeout 300 python profiles/tools/desync.py $id 30 > gpurun_out/${tag}_plain.txt 2>&1 || { echo "plain run failed"; exit 1; }
# desync.py 30: 1 reset + 33 synchronised launches, then 3 + 30 with spread-out ends
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rollout --launch-skip $((34 + skip)) --launch-count 1 -f -o gpurun_out/$tag \
    python profiles/tools/desync.py $id 30 > gpurun_out/${tag}_ncu.log 2>&1
python profiles/tools/summarize.py gpurun_out/$tag.ncu-rep $tag k_rollout$kern 1048576 32 --out gpurun_out > /dev/null 2> gpurun_out/${tag}_summarize.err
python profiles/tools/line_stalls.py gpurun_out/$tag.ncu-rep gym_minigrid_b200/libmgb200.so k_rollout$kern 1048576 > gpurun_out/${tag}_line_stalls.txt 2>&1
rm -f gpurun_out/$tag.ncu-rep
ls -la gpurun_out | grep $tag
