#!/usr/bin/env python
"""Copy what `evidence.sh <round>` left in gpurun_out/ into profiles/ (tracked): bench lines, stamped ncu summaries
(line_stalls trimmed to the lines holding >= 0.25 % of the stall samples or >= 2 warp-instructions per warp-step),
traffic.json, launch list, sustained / spread-out-end tables, and the SASS opcode histograms of the library in the tree.

    python profiles/tools/collect_evidence.py r2
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PROFILES = os.path.dirname(HERE)
ROOT = os.path.dirname(PROFILES)
OUT = os.path.join(ROOT, "gpurun_out")
R = sys.argv[1] if len(sys.argv) > 1 else "r2"

KERNELS = {"empty8x8": "k_rolloutILi0ELb1ELi7", "doorkey16x16": "k_rolloutILi1ELb0ELi7", "fourrooms": "k_rolloutILi2ELb0ELi7",
           "dynobs16x16": "k_rolloutILi3ELb1ELi7", "keycorridors6r3": "k_rolloutILi4ELb0ELi7"}


def trim(src, dst):
    keep = []
    for i, ln in enumerate(open(src)):
        if i < 3 or "|" not in ln:
            keep.append(ln)
            continue
        f = ln.split("|")[0].split()
        try:
            instr, samp = float(f[1]), float(f[3].rstrip("%"))
        except (IndexError, ValueError):
            keep.append(ln)
            continue
        if samp >= 0.25 or instr >= 2.0:
            keep.append(ln)
    open(dst, "w").writelines(keep)


def main():
    # gpurun_out/ is scratch and keeps files of earlier calls: take what the last evidence run wrote (its first output is
    # the reference-arm bench line)
    t0 = os.path.getmtime(os.path.join(OUT, R + "_bench_reference_arm.json")) - 120
    for f in sorted(glob.glob(os.path.join(OUT, R + "_*"))):
        b = os.path.basename(f)
        if os.path.getmtime(f) < t0:
            continue
        if b.endswith(("_plain.json", "_plain.err", "_ncu.log", "_summarize.err", "_stamp.txt", ".err", "_evidence.log", ".ncu-rep")):
            continue
        if b.endswith("_line_stalls.txt"):
            trim(f, os.path.join(PROFILES, b))
        elif b.endswith((".json", ".txt", ".csv", ".jsonl")) and os.path.getsize(f) < 200_000:
            text = open(f).read()
            if b.endswith(".txt"):
                text = "".join(ln for ln in text.splitlines(True) if not ln.startswith("+ "))      # `set -x` traces
            open(os.path.join(PROFILES, b), "w").write(text)
    t = os.path.join(OUT, "traffic.json")
    if os.path.exists(t):
        shutil.copy(t, os.path.join(PROFILES, "traffic.json"))
    lib = os.path.join(ROOT, "gym_minigrid_b200", "libmgb200.so")
    import ctypes
    L = ctypes.CDLL(lib)
    L.mgb_version.restype = ctypes.c_char_p
    stamp = L.mgb_version().decode()
    for short, k in KERNELS.items():
        out = subprocess.run([sys.executable, os.path.join(HERE, "sass_hist.py"), lib, "--kernel", k], stdout=subprocess.PIPE, text=True).stdout
        open(os.path.join(PROFILES, "%s_sass_opcodes_%s.txt" % (R, short)), "w").write(
            "# library: %s\n# cuobjdump -sass opcode histogram, no cut-off (UBLKCP = cp.async.bulk, SYNCS = mbarrier, UTMACMDFLUSH = bulk commit, LDGSTS = cp.async)\n%s" % (stamp, out))
    print("collected into profiles/ for", stamp)


if __name__ == "__main__":
    main()
