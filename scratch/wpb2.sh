for w in $2; do echo "wpb=$w"; MGB_WARPS_PER_BLOCK=$w bash profiles/tools/ab.sh "$1" scratch/noseg.so; done
