#!/bin/bash
# Round evidence in one GPU call:  gpurun --timeout 1500 -- 'bash profiles/tools/evidence.sh r2'
# 1. the bench line (both arms)  2. the ncu launch list of the same bench command  3. one `ncu --set full` capture per
# BASELINE config (rollout launch) + one single-step launch  4. sustained runs (>= 4 s per config, clocks sampled).
R="${1:-r2}"
set -x
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${R}_bench_reference_arm.json 2> gpurun_out/${R}_bench_reference_arm.err
python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_1gpu.json 2> gpurun_out/${R}_bench_1gpu.err
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs > gpurun_out/${R}_launchlist_plain.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench_empty8x8.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs > gpurun_out/${R}_launchlist_ncu.log 2>&1
# reports are ~14 MB each and gpurun brings back at most 64 MiB: they are summarised HERE (ncu -i works on the box) and deleted
for triple in "MiniGrid-Empty-8x8-v0:empty8x8:ILi0ELb1ELi7" "MiniGrid-DoorKey-16x16-v0:doorkey16x16:ILi1ELb0ELi7" "MiniGrid-FourRooms-v0:fourrooms:ILi2ELb0ELi7" \
              "MiniGrid-Dynamic-Obstacles-16x16-v0:dynobs16x16:ILi3ELb1ELi7" "MiniGrid-KeyCorridorS6R3-v0:keycorridors6r3:ILi4ELb0ELi7"; do
  IFS=: read id short kern <<< "$triple"
  tag="${R}_k_rollout_${short}"
  bash profiles/tools/capture.sh "$id" "$tag"
  extra=""; [ "$short" = "empty8x8" ] && extra="--traffic $id"
  python profiles/tools/summarize.py gpurun_out/$tag.ncu-rep $tag k_rollout$kern 1048576 32 --out gpurun_out $extra > /dev/null 2> gpurun_out/${tag}_summarize.err
  rm -f gpurun_out/$tag.ncu-rep
done
tag=${R}_k_step_empty8x8
bash profiles/tools/capture.sh MiniGrid-Empty-8x8-v0 $tag "" 8
python profiles/tools/summarize.py gpurun_out/$tag.ncu-rep $tag k_rolloutILi0ELb1ELi7 1048576 1 --out gpurun_out > /dev/null 2> gpurun_out/${tag}_summarize.err
rm -f gpurun_out/$tag.ncu-rep
python profiles/tools/sustained.py --seconds 4 > gpurun_out/${R}_sustained.txt 2> gpurun_out/${R}_sustained.err
tail -3 gpurun_out/${R}_sustained.txt
# 5. episode ends spread over time (every env its own remaining time) next to the lock-step case: the five BASELINE configs
#    and the generator families that keep spare layouts; then the spare-layout counters of one KeyCorridor run
#    (build/ab/dbg.so = the same source with -DMGB_DEBUG_SPARES=1, build/ab/nospares.so = -DMGB_SPARES=0)
(python -c "from gym_minigrid_b200 import _lib; print('# library:', _lib.load().mgb_version().decode())"
 for e in MiniGrid-Empty-8x8-v0 MiniGrid-DoorKey-16x16-v0 MiniGrid-FourRooms-v0 MiniGrid-Dynamic-Obstacles-16x16-v0 MiniGrid-KeyCorridorS6R3-v0 \
          MiniGrid-KeyCorridorS3R3-v0 MiniGrid-SimpleCrossingS11N5-v0 MiniGrid-LavaCrossingS9N2-v0 MiniGrid-MultiRoom-N6-v0 MiniGrid-DistShift1-v0; do
   timeout 120 python profiles/tools/desync.py $e 40
 done) > gpurun_out/${R}_desync.txt 2>&1
if [ -f build/ab/nospares.so ]; then
 (for e in MiniGrid-KeyCorridorS6R3-v0 MiniGrid-KeyCorridorS3R3-v0 MiniGrid-SimpleCrossingS11N5-v0 MiniGrid-LavaCrossingS9N2-v0 MiniGrid-MultiRoom-N6-v0; do
   MGB_LIB=$PWD/build/ab/nospares.so timeout 120 python profiles/tools/desync.py $e 40 | sed "s|^|-DMGB_SPARES=0 |"
 done) > gpurun_out/${R}_desync_without_spares.txt 2>&1
fi
[ -f build/ab/dbg.so ] && MGB_LIB=$PWD/build/ab/dbg.so timeout 120 python profiles/tools/spare_probe.py MiniGrid-KeyCorridorS6R3-v0 44 > gpurun_out/${R}_spare_probe_keycorridor.txt 2>&1
tail -4 gpurun_out/${R}_desync.txt
