#!/usr/bin/env python
"""Opcode histogram (no cut-off) and opcode sequence of one kernel of a built library, from `cuobjdump -sass`.

    python profiles/tools/sass_hist.py [lib.so] --kernel 'k_rolloutILi0ELb1ELi7' [--seq out.txt]

Used (1) to show the TMA-unit bulk copies / mbarriers in the product kernels (UBLKCP, SYNCS, UTMACMDFLUSH are what
cp.async.bulk.* and mbarrier.* compile to on sm_100a) and (2) to check that a source clean-up left the instruction
stream of the shipped kernels unchanged (compare two --seq files)."""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def kernels(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
    cur, out = None, collections.OrderedDict()
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m and cur:
            ins = m.group(1).strip()
            ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
            out[cur].append(ins)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("lib", nargs="?", default=os.path.join(ROOT, "gym_minigrid_b200", "libmgb200.so"))
    ap.add_argument("--kernel", default="k_rolloutILi0ELb1ELi7")
    ap.add_argument("--seq", default=None, help="write the opcode sequence (operands stripped) to this file")
    a = ap.parse_args()
    ks = kernels(a.lib)
    for name, ins in ks.items():
        if a.kernel not in name:
            continue
        ops = [i.split()[0] for i in ins]
        h = collections.Counter(o.split(".")[0] for o in ops)
        print("== %s: %d static instructions" % (name, len(ops)))
        for o, n in sorted(h.items(), key=lambda kv: -kv[1]):
            print("  %-14s %6d" % (o, n))
        if a.seq:
            with open(a.seq, "w") as f:
                f.write("\n".join(ops) + "\n")


if __name__ == "__main__":
    main()
