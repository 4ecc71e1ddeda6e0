#!/usr/bin/env python
"""Turn an .ncu-rep (one k_rollout launch, --set full) into the small text/CSV summaries that are
committed under profiles/:  <tag>_metrics.txt (selected raw metrics + stall reasons),
<tag>_by_line.txt (instruction mix by opcode and by CUDA source line), and, for the bench
workload, traffic.json (dram bytes per launch, read by bench.py for roofline.traffic).

    python profiles/tools/summarize.py <report.ncu-rep> <tag> <kernel-substring> <num_envs> <T> [--traffic ENV_ID]
"""
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PROFILES = os.path.dirname(HERE)
ROOT = os.path.dirname(PROFILES)

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    rep, tag, ksub, n_envs, T = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    so = sys.argv[sys.argv.index("--lib") + 1] if "--lib" in sys.argv else os.path.join(ROOT, "gym_minigrid_b200", "libmgb200.so")
    global PROFILES
    if "--out" in sys.argv:                      # e.g. gpurun_out/ when run on the GPU box (reports are too large to bring back)
        PROFILES = sys.argv[sys.argv.index("--out") + 1]
    # the library the capture was taken with names its own source revision and -D switches (mgb_version()); capture.sh stores
    # it next to the report.  A by-line join is only valid against that very library.
    stamp_file = os.path.splitext(rep)[0] + "_stamp.txt"
    import ctypes
    lib = ctypes.CDLL(so)
    lib.mgb_version.restype = ctypes.c_char_p
    lib_stamp = lib.mgb_version().decode()
    cap_stamp = open(stamp_file).read().strip() if os.path.exists(stamp_file) else None
    if cap_stamp is not None and cap_stamp != lib_stamp:
        sys.exit("summarize.py: %s was captured with [%s] but %s is [%s]: rebuild that revision (--lib) for a valid join" % (
            os.path.basename(rep), cap_stamp, so, lib_stamp))
    stamp = "# library: %s%s" % (lib_stamp, "" if cap_stamp else "  (no capture stamp: join unverified)")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
    lines = ["# %s  (kernel %s, %d envs x T=%d, from %s)" % (tag, d.get("Kernel Name", "?"), n_envs, T, os.path.basename(rep)), stamp]
    for k in KEYS:
        if k in d:
            lines.append("%-72s %s %s" % (k, d[k], u[k]))
    inst = num(d.get("smsp__inst_executed.sum", "0")) or 0
    wsteps = n_envs / 32.0 * T
    lines.append("thread-instructions per env-step (= warp-instructions per warp-step): %.1f" % (inst / wsteps))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = (num(d["dram__bytes_read.sum"]) or 0) * scale.get(u["dram__bytes_read.sum"], 1)
    wr = (num(d["dram__bytes_write.sum"]) or 0) * scale.get(u["dram__bytes_write.sum"], 1)
    algo = 158 * n_envs * T
    lines.append("dram bytes per launch: read %.4g + write %.4g = %.4g ; algorithmic bytes %.4g ; ratio %.3f" % (rd, wr, rd + wr, algo, (rd + wr) / algo))
    lines.append("")
    lines.append("warp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active.ratio > 0.1):")
    for k in hdr:
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
            v = num(d[k])
            if v and v > 0.1:
                lines.append("  %-40s %.3f" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
    open(os.path.join(PROFILES, tag + "_metrics.txt"), "w").write("\n".join(lines) + "\n")
    bl = subprocess.run([sys.executable, os.path.join(HERE, "sass_by_line.py"), rep, so, ksub, str(wsteps)],
                        stdout=subprocess.PIPE, text=True).stdout
    open(os.path.join(PROFILES, tag + "_by_line.txt"), "w").write(stamp + "\n" + bl)
    bf = subprocess.run([sys.executable, os.path.join(HERE, "by_function.py"), rep, so, ksub, str(wsteps)],
                        stdout=subprocess.PIPE, text=True).stdout
    open(os.path.join(PROFILES, tag + "_by_function.txt"), "w").write(stamp + "\n" + bf)
    ls = subprocess.run([sys.executable, os.path.join(HERE, "line_stalls.py"), rep, so, ksub, str(wsteps)],
                        stdout=subprocess.PIPE, text=True).stdout
    open(os.path.join(PROFILES, tag + "_line_stalls.txt"), "w").write(stamp + "\n" + ls)
    if "--traffic" in sys.argv:
        env_id = sys.argv[sys.argv.index("--traffic") + 1]
        json.dump({"env_id": env_id, "num_envs": n_envs, "rollout_T": T, "dram_bytes_per_launch": rd + wr,
                   "dram_bytes_read": rd, "dram_bytes_write": wr, "source": os.path.basename(rep), "library": lib_stamp},
                  open(os.path.join(PROFILES, "traffic.json"), "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
