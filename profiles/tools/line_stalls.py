#!/usr/bin/env python
"""Per CUDA source line of one kernel: executed warp-instructions per warp-step, average active threads, and the
warp-stall samples by reason (ncu source page joined with nvdisasm line info).

    python profiles/tools/line_stalls.py <report.ncu-rep> <lib.so> <mangled-kernel-substring> <warp_steps> [lo hi]
"""
import collections
import csv
import io
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_by_line as S  # noqa: E402

REASONS = ["stall_wait", "stall_short_sb", "stall_not_selected", "stall_math", "stall_dispatch", "stall_branch_resolving",
           "stall_long_sb", "stall_no_inst", "stall_mio", "stall_selected"]


def main():
    rep, so, ksub, ws = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    lo, hi = (int(sys.argv[5]), int(sys.argv[6])) if len(sys.argv) > 6 else (0, 1 << 30)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ix = {n: hdr.index(n) for n in ["Instructions Executed", "Thread Instructions Executed", "# Samples"] + REASONS}
    lt = S.line_table(so, ksub)
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    if len(body) != len(lt):
        print("# warning: %d profiled instructions vs %d disassembled" % (len(body), len(lt)))
    agg = collections.defaultdict(lambda: collections.Counter())
    tot = collections.Counter()
    for r, (_, line, _ins) in zip(body, lt):
        for n, i in ix.items():
            try:
                v = int(r[i])
            except ValueError:
                v = 0
            agg[line][n] += v
            tot[n] += v
    src = {}
    print("total: %.1f warp-instr/warp-step, %d samples; by reason: %s" % (
        tot["Instructions Executed"] / ws, tot["# Samples"], " ".join("%s=%.1f%%" % (k[6:], 100.0 * tot[k] / max(1, tot["# Samples"])) for k in REASONS)))
    print("%-26s %8s %6s %7s | %s" % ("line", "instr", "thr", "samp%", " ".join(k[6:12] for k in REASONS)))
    for line in sorted(agg, key=lambda l: (l is None, l)):
        if line is None or not (lo <= line[1] <= hi):
            continue
        a = agg[line]
        if line[0] not in src:
            p = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", line[0])
            src[line[0]] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src[line[0]][line[1] - 1].strip()[:70] if 0 < line[1] <= len(src[line[0]]) else ""
        ie = a["Instructions Executed"]
        print("%-26s %8.1f %6.1f %6.2f%% | %s | %s" % (
            "%s:%d" % line, ie / ws, a["Thread Instructions Executed"] / max(1, ie), 100.0 * a["# Samples"] / max(1, tot["# Samples"]),
            " ".join("%6.2f" % (100.0 * a[k] / max(1, tot["# Samples"])) for k in REASONS), text))


if __name__ == "__main__":
    main()
