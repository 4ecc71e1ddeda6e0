for w in 8 7 6 5 4; do echo "wpb=$w"; MGB_WARPS_PER_BLOCK=$w bash profiles/tools/ab.sh "$1" scratch/noseg.so; done
