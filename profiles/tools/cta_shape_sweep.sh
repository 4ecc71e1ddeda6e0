# needs a library built with -DMGB_EXPERIMENT (python -m gym_minigrid_b200.build -DMGB_EXPERIMENT --out=build/ab/exp.so; MGB_LIB=...)
for w in $2; do echo "wpb=$w"; MGB_WARPS_PER_BLOCK=$w bash profiles/tools/ab.sh "$1" ${3:-gym_minigrid_b200/libmgb200.so}; done
