for lib in "$@"; do echo "== $lib"; MGB_LIB=$lib PYTHONPATH=. timeout 300 python scratch/dyn_probe.py 2>&1 | tail -2; done
