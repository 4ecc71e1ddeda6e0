#!/usr/bin/env python
"""Join an ncu SASS-level source page (per-instruction executed counts / stall samples) with
nvdisasm line info of the same cubin, and aggregate by CUDA source line and by opcode.

    python profiles/tools/sass_by_line.py <report.ncu-rep> <lib.so> <kernel-substring> <warp_steps>

warp_steps = (#envs/32) * T of the profiled launch, to express counts per warp-step
(= thread-instructions per env-step).  Read-only analysis; nothing here runs on the GPU.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def line_table(so, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], stdout=subprocess.PIPE, text=True).stdout
    lines, cur, infn = [], None, False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            infn = kernel_sub in m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return lines


def main():
    rep, so, ksub, wsteps = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
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
    lt = line_table(so, ksub)
    n = min(len(inst), len(lt))
    if len(inst) != len(lt):
        print("# warning: %d profiled instructions vs %d disassembled" % (len(inst), len(lt)))
    by_line = collections.Counter()
    samp_line = collections.Counter()
    by_op = collections.Counter()
    tot = ts = 0
    for (src, ex, sm), (_, where, _) in zip(inst[:n], lt[:n]):
        by_line[where] += ex
        samp_line[where] += sm
        op = re.sub(r"^@!?U?P\w+\s+", "", src).split()[0].split(".")[0]
        by_op[op] += ex
        tot += ex
        ts += sm
    print("total warp-instructions %d = %.1f per warp-step (thread-instructions per env-step)" % (tot, tot / wsteps))
    print("\n== by opcode ==")
    for k, v in by_op.most_common(24):
        print("%-12s %7.1f  %5.1f%%" % (k, v / wsteps, 100.0 * v / tot))
    print("\n== by source line ==")
    for k, v in by_line.most_common(45):
        print("%-28s %7.1f  %5.1f%%  stall-samples %5.1f%%" % ("%s:%s" % k if k else "?", v / wsteps, 100.0 * v / tot, 100.0 * samp_line[k] / max(ts, 1)))


if __name__ == "__main__":
    main()
