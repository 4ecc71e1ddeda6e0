#!/usr/bin/env python
"""Aggregate an ncu SASS source page by the device function each instruction was inlined from.

    python profiles/tools/by_function.py <report.ncu-rep> <lib.so> <mangled-kernel-substring> <warp_steps>
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_by_line as S  # noqa: E402


def main():
    rep, so, ksub, ws = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    inst = []
    for r in rows[2:]:
        try:
            inst.append((r[iS].strip(), int(r[iE]), int(r[iSm])))
        except (ValueError, IndexError):
            pass
    lt = S.line_table(so, ksub)
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    src = open(os.path.join(root, "gym_minigrid_b200", "csrc", "mgb_kernels.cuh")).read().split("\n")
    starts = []
    for i, l in enumerate(src, 1):
        m = re.match(r"(?:template.*>\s*)?__(?:device|global)__.*?(\w+)\(", l)
        if m and m.group(1) != "__launch_bounds__":
            starts.append((i, m.group(1)))
        elif "k_rollout(" in l:
            starts.append((i, "k_rollout"))

    def fn(line):
        name = "?"
        for i, n in starts:
            if i <= line:
                name = n
            else:
                break
        return name
    agg, sm = collections.Counter(), collections.Counter()
    for (s, ex, sa), (_, where, _) in zip(inst, lt):
        k = fn(where[1]) if where and where[0] == "mgb_kernels.cuh" else (where[0] if where else "?")
        agg[k] += ex
        sm[k] += sa
    tot, ts = sum(agg.values()), sum(sm.values())
    print("total %.1f warp-instructions per warp-step" % (tot / ws))
    for k, v in agg.most_common(24):
        print("%-28s %8.1f %5.1f%%  stall-samples %5.1f%%" % (k, v / ws, 100 * v / tot, 100 * sm[k] / max(ts, 1)))


if __name__ == "__main__":
    main()
