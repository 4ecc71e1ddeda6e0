#!/usr/bin/env python
"""bench.py -- env-steps/sec (observations materialised in HBM) of the batched MiniGrid hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--env-id ID] [--num-envs N_PER_GPU] [--rollout-T T]

One bench "step" = one pass of the hot path over one batch: ONE persistent mgb_rollout launch that
advances every env of the batch by T env-steps and writes all T*N observations (+reward/done/dir)
to HBM.  value = (envs over all ranks) * T * K / max-over-ranks device time.

Workload (BASELINE.json configs[0], fits one GPU): MiniGrid-Empty-8x8-v0, uniform random actions
pre-generated on the device (torch.randint, seed 1234), 2^20 envs per GPU, T = 32, auto-reset on.
Outputs per launch (~5 GB) are far larger than L2 (126 MB), so no L2 flush is needed between steps.

Extra keys: roofline (HBM bound, 158 algorithmic B / env-step), cpu_baseline (C oracle port on the
host cores, bounded sample), e2e (mgb_step_host: pinned host actions in, host obs/reward/done/dir
out, copies inside the timed region), clocks (nvidia-smi sampled during the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_STEP = 158        # 147 obs + 8 reward + 1 done + 1 direction (writes) + 1 action (read); SURVEY §8d
METRIC = "env-steps/sec (with obs)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--env-id", default="MiniGrid-Empty-8x8-v0")
    ap.add_argument("--num-envs", type=int, default=1 << 20, help="envs per GPU")
    ap.add_argument("--rollout-T", type=int, default=32, help="env-steps per launch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-python-reference", action="store_true", help="skip timing the unmodified Python reference on the host cores")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short per-GPU runs of the other BASELINE configs")
    ap.add_argument("--cpu-sample-envs", type=int, default=0, help="CPU-port sample (default: 16384 envs in the b200 arm, --num-envs in the reference arm)")
    ap.add_argument("--cpu-sample-T", type=int, default=0, help="default: 64 in the b200 arm, --rollout-T in the reference arm")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            try:
                pw.append(float(f[3]))
            except ValueError:
                pass
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "sm_mhz_min": min(sm) if sm else None,
                "power_w_max": max(pw) if pw else None}


def oracle_cfg(env_id):
    import gym_minigrid_b200 as mgb
    return {k: v for k, v in mgb.spec(env_id)["config"].items() if k not in ("mission", "reward_range")}


def bench_config(env_id, n_envs, T):
    """`config` of the JSON line -- the same dict for the b200 arm and the reference arm"""
    return {"workload": env_id + " random-action rollout, auto-reset on, obs+reward+done+dir written every env-step",
            "envs_per_gpu": n_envs, "env_steps_per_launch": T, "actions": "uniform random, pre-generated, seed 1234",
            "l2": "outputs per launch (%.1f GB) exceed L2; no flush needed" % (n_envs * T * 157 / 1e9),
            "parallelism": "env shards by global env id, no collective"}


class CpuLeg:
    """The CPU oracle (the C port of the reference algorithm, oracle/minigrid_oracle.c) on the host cores: one call of
    run() = one pass of the hot path over n_envs x T env-steps, observations materialised in (re-used) host buffers."""

    def __init__(self, env_id, n_envs, T, threads=0):
        import numpy as np
        from oracle.oracle import OracleVec
        cfg = oracle_cfg(env_id)
        self.cores = threads or os.cpu_count() or 1
        self.n, self.T = n_envs, T
        self.orc = OracleVec(cfg, n_envs, seed=0, threads=self.cores)
        self.orc.reset()
        rs = np.random.RandomState(1234)
        self.acts = rs.randint(0, cfg["n_actions"], size=(T, n_envs)).astype(np.uint8)
        self.out = self.orc.rollout(self.acts, autoreset=True)             # warm-up; the buffers are touched now

    def run(self):
        t0 = time.perf_counter()
        self.orc.rollout(self.acts, autoreset=True, out=self.out)          # obs [T,n,147] materialised in host memory
        return time.perf_counter() - t0


def cpu_leg(env_id, n_envs, T, reps, threads=0):
    leg = CpuLeg(env_id, n_envs, T, threads)
    times = [leg.run() for _ in range(reps)]
    return n_envs * T / min(times), times, leg.cores


def _ref_worker(env_id, idx, warm_s, timed_s, q):
    """One forked worker of the Python-reference harness: the unmodified reference env, uniform
    random actions, reset() on done, observation dict built every step."""
    try:
        import numpy as np
        from oracle import ref_shim as R
        env = R.make(env_id)
        env.seed(idx)
        env.reset()
        n_act = env.action_space.n
        rs = np.random.RandomState(1000 + idx)

        def run(sec):
            n, t_end = 0, time.perf_counter() + sec
            while time.perf_counter() < t_end:
                for a in rs.randint(0, n_act, size=32):
                    _, _, d, _ = env.step(int(a))
                    if d:
                        env.reset()
                n += 32
            return n
        run(warm_s)
        t0 = time.perf_counter()
        n = run(timed_s)
        q.put((n, time.perf_counter() - t0))
    except Exception as e:                       # report, never hang the parent
        q.put(("error", repr(e)))


def _ref_rate(env_id, procs, warm_s, timed_s):
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    ws = [ctx.Process(target=_ref_worker, args=(env_id, i, warm_s, timed_s, q)) for i in range(procs)]
    for w in ws:
        w.start()
    res = [q.get(timeout=60 + 4 * (warm_s + timed_s)) for _ in ws]
    for w in ws:
        w.join(timeout=10)
    bad = [r for r in res if r[0] == "error"]
    if bad:
        raise RuntimeError(bad[0][1])
    return sum(n / dt for n, dt in res)


def _benchmark_py(env_id):
    """The three numbers of the reference's own benchmark.py (benchmark.py:22-53), bounded iteration counts."""
    from oracle import ref_shim as R
    W = sys.modules["gym_minigrid.wrappers"]
    env = R.make(env_id)
    n_reset, n_frames = 100, 300
    t0 = time.perf_counter()
    for _ in range(n_reset):
        env.reset()
    reset_ms = 1e3 * (time.perf_counter() - t0) / n_reset
    t0 = time.perf_counter()
    for _ in range(n_frames):
        env.render('rgb_array')
    render_fps = n_frames / (time.perf_counter() - t0)
    env = W.ImgObsWrapper(W.RGBImgPartialObsWrapper(R.make(env_id)))
    env.reset()
    t0 = time.perf_counter()
    for _ in range(n_frames):
        env.step(0)
    view_fps = n_frames / (time.perf_counter() - t0)
    return {"reset_ms": reset_ms, "render_fps": render_fps, "agent_view_fps": view_fps, "resets": n_reset, "frames": n_frames}


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def python_reference_leg(env_id, others=()):
    """The UNMODIFIED Python reference timed on this box's host cores: P =
    os.cpu_count() forked workers, 1 s warm-up + 5 s timed, sum of steps/s; single-process rate; benchmark.py's numbers.
    Must run before CUDA is initialised in this process (fork).  Returns a dict, or {"value": None, "reason": ...}."""
    try:
        from oracle import ref_shim as R
        root = R.reference_root()
        if root is None:
            msg = "reference tree not found: looked at $MGB_REFERENCE, /root/reference, baseline/_ref"
            print("bench.py: cpu_baseline_python unavailable -- " + msg, file=sys.stderr)
            return {"value": None, "reason": msg}
        R.load_reference()
        P = os.cpu_count() or 1
        single = _ref_rate(env_id, 1, 0.3, 1.0)
        total = _ref_rate(env_id, P, 1.0, 5.0)
        out = {"value": total, "unit": "env-steps/s", "cores": P, "cpu_model": cpu_model(), "single_process": single,
               "kind": "reference", "reference_root": root,
               "sample": "%d forked workers x (1 s warm-up + 5 s timed), unmodified reference under the oracle/ref_shim.py gym "
                         "stub, uniform random actions, reset() on done, obs dict built every step" % P,
               "other_configs": {}, "benchmark_py": None}
        for oid in others:
            out["other_configs"][oid] = _ref_rate(oid, P, 0.5, 2.0)
        try:
            out["benchmark_py"] = _benchmark_py(env_id)
        except Exception as e:
            out["benchmark_py"] = {"error": repr(e)}
        return out
    except Exception as e:
        print("bench.py: cpu_baseline_python failed -- %r" % (e,), file=sys.stderr)
        return {"value": None, "reason": repr(e)}


def run_reference(args, rank, world):
    """--impl reference: the reference is pure Python and does not travel to the GPU box; per the
    task contract the arm times the CPU oracle port with all host threads on the same config."""
    if rank != 0:
        return
    pyref = None if args.no_python_reference else python_reference_leg(args.env_id)
    # by default the arm runs the b200 arm's own config (same envs per "GPU", same env-steps per bench step)
    n = args.cpu_sample_envs if args.cpu_sample_envs else args.num_envs
    T = args.cpu_sample_T if args.cpu_sample_T else args.rollout_T
    leg = CpuLeg(args.env_id, n, T)
    cores = leg.cores
    for _ in range(args.warmup):
        leg.run()
    per = [leg.run() for _ in range(args.steps)]
    value = n * T * len(per) / sum(per)
    sample = "%d envs x %d env-steps per bench step, C oracle port, %d threads" % (n, T, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(per) / len(per), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": bench_config(args.env_id, n, T),       # the b200 arm's config: a bench step of this arm is one pass over the same batch
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "cpu_baseline_python": pyref,
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


OTHER_CONFIGS_OF = "MiniGrid-Empty-8x8-v0"          # the config the metric is quoted on (BASELINE.json north_star)
OTHER_CONFIGS = ["MiniGrid-DoorKey-16x16-v0", "MiniGrid-FourRooms-v0", "MiniGrid-Dynamic-Obstacles-16x16-v0",
                 "MiniGrid-KeyCorridorS6R3-v0"]


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # the unmodified Python reference on this box's host cores, rank 0 at N = 1 only; forks, so it runs before CUDA exists
    pyref = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.no_python_reference:
        pyref = python_reference_leg(args.env_id, OTHER_CONFIGS if args.env_id == OTHER_CONFIGS_OF and not args.no_other_configs else ())

    import torch
    import torch.distributed as dist
    import gym_minigrid_b200 as mgb

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    N, T, K, W = args.num_envs, args.rollout_T, args.steps, max(args.warmup, 3)
    cfg = mgb.spec(args.env_id)["config"]
    env = mgb.make(args.env_id, num_envs=N, device=dev, seed=0, env_id_base=rank * N)   # shard by global env id
    env.reset()
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    n_pool = 4
    acts = [torch.randint(0, cfg["n_actions"], (T, N), dtype=torch.uint8, device=dev, generator=g) for _ in range(n_pool)]
    out = (torch.empty((T, N, 7, 7, 3), dtype=torch.uint8, device=dev),
           torch.empty((T, N), dtype=torch.float64, device=dev),
           torch.empty((T, N), dtype=torch.uint8, device=dev),
           torch.empty((T, N), dtype=torch.uint8, device=dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(W):
        env.rollout(acts[i % n_pool], out=out)
    barrier()
    launches0 = env.kernel_launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e_start.record()
    for i in range(K):
        evs[i][0].record()
        env.rollout(acts[i % n_pool], out=out)
        evs[i][1].record()
    e_end.record()
    barrier()
    total_ms = e_start.elapsed_time(e_end)
    launches = env.kernel_launches - launches0
    kern_ms = [a.elapsed_time(b) for a, b in evs]
    clocks = sampler.stop() if rank == 0 else None
    env.check_errors()
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * N * T * K / (total_ms_max * 1e-3)

    # ---- roofline of the dominant (only) kernel: algorithmic bytes / CUDA-event launch duration ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    avg_kern_s = statistics.mean(kern_ms) * 1e-3
    achieved = ALGO_BYTES_PER_STEP * N * T / avg_kern_s / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")        # dram bytes per launch from the committed ncu capture
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if tj.get("env_id") == args.env_id and tj.get("num_envs") == N and tj.get("rollout_T") == T:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "mgb::k_rollout", "algorithmic_bytes_per_launch": ALGO_BYTES_PER_STEP * N * T,
                "avg_launch_ms": statistics.mean(kern_ms), "peak_source": peak_src}

    # ---- e2e: host buffers through mgb_step_host (H2D actions + kernel + D2H results per step) ----
    e2e = None
    if not args.no_e2e:
        Ke = max(3, min(K, 10))
        # host buffers are pinned next to this rank's GPU (one process per GPU: no D2H stream crosses the socket link)
        from gym_minigrid_b200.sharding import bind_to_gpu_numa_node
        affinity0 = os.sched_getaffinity(0)
        numa = bind_to_gpu_numa_node(local)
        hacts = [torch.randint(0, cfg["n_actions"], (N,), dtype=torch.uint8).pin_memory() for _ in range(2)]
        for i in range(2):
            env.step_host(hacts[i % 2])
        barrier()
        t0 = time.perf_counter()
        for i in range(Ke):
            env.step_host(hacts[i % 2])            # synchronous: outputs are in pinned host memory on return
        torch.cuda.synchronize(dev)
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": world * N * Ke / float(te.item()), "unit": "env-steps/s", "h2d_bytes_per_step": N,
               "d2h_bytes_per_step": N * (147 + 8 + 1 + 1), "steps": Ke, "api": "VecMiniGridEnv.step_host -> mgb_step_host",
               "numa_node": numa}
        # The ceiling of this number: the same bytes per step as BARE pinned copies (no kernel), all ranks at the same time.
        # e2e is PCIe/host-bound; `frac_of_copy_ceiling` says how much of what the box can copy the pipeline delivers.
        d2h_bytes = N * (147 + 8 + 1 + 1)
        dbuf = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
        hbuf = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
        dact = torch.empty(N, dtype=torch.uint8, device=dev)
        for i in range(2 + Ke):
            if i == 2:
                barrier()
                t0 = time.perf_counter()
            dact.copy_(hacts[i % 2], non_blocking=True)
            hbuf.copy_(dbuf, non_blocking=True)
            torch.cuda.synchronize(dev)
        tc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        ceil_steps = world * N * Ke / float(tc.item())
        e2e.update({"copy_ceiling_env_steps_per_s": ceil_steps, "copy_ceiling_gbs": ceil_steps * 158 / 1e9,
                    "achieved_gbs": e2e["value"] * 158 / 1e9, "frac_of_copy_ceiling": e2e["value"] / ceil_steps})
        try:
            e2e["numa_nodes_online"] = open("/sys/devices/system/node/online").read().strip()
        except OSError:
            pass
        del dbuf, hbuf, dact
        os.sched_setaffinity(0, affinity0)

    # ---- secondary: single-step launches, i.e. the gym API env.step() (state round-trips HBM every step) ----
    def step_mode_of(e, ecfg, a_TN):
        Ks = 16
        o1 = (out[0][0], out[1][0], out[2][0], out[3][0])
        for t_ in range(3):
            e.step(a_TN[t_], out=o1)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for t_ in range(Ks):
            e.step(a_TN[t_ % T], out=o1)
        s1.record()
        barrier()
        rate = N * Ks / (s0.elapsed_time(s1) * 1e-3)
        # bytes a single-step launch must move per env-step: the 158 output/action bytes plus the env's state block read
        # (S words) and its non-grid words written back.  Template-grid envs (Empty, Dynamic-Obstacles:
        # grid = static template + ball list) never move their grid rows: S-GW words in, S-GW words out.
        hp = (ecfg["height"] + 3) // 4 * 4
        gw = ecfg["width"] * hp // 4
        s_words = gw + 4 + (4 if ecfg.get("n_obstacles", 0) > 0 else 0)
        implied = ecfg["gen"] in (0, 3)
        sbytes = 158 + 4 * (s_words - gw if implied else s_words) + 4 * (s_words - gw)
        return {"env_steps_per_s_per_gpu": rate, "state_inclusive_bytes_per_env_step": sbytes, "state_inclusive_gbs": rate * sbytes / 1e9,
                "frac_of_hbm_peak": (rate * sbytes / 1e9) / peak}

    step_mode_info = step_mode_of(env, cfg, acts[0])
    step_mode = step_mode_info["env_steps_per_s_per_gpu"]

    # ---- secondary: the other BASELINE.json configs on the same kernel family, per GPU (short: 3 warm-up + 5 launches each) ----
    others = None
    if not args.no_other_configs and args.env_id == OTHER_CONFIGS_OF:
        others = {}
        del env
        for oid in OTHER_CONFIGS:
            ocfg = mgb.spec(oid)["config"]
            oenv = mgb.make(oid, num_envs=N, device=dev, seed=0, env_id_base=rank * N)
            oenv.reset()
            oa = torch.randint(0, ocfg["n_actions"], (T, N), dtype=torch.uint8, device=dev, generator=g)
            for _ in range(3):
                oenv.rollout(oa, out=out)
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(5):
                oenv.rollout(oa, out=out)
            s1.record()
            barrier()
            ov = N * T * 5 / (s0.elapsed_time(s1) * 1e-3)
            oenv.check_errors()
            others[oid] = {"env_steps_per_s_per_gpu": ov, "roofline_frac": ov * ALGO_BYTES_PER_STEP / 1e9 / peak,
                           "step_mode": step_mode_of(oenv, ocfg, oa)}
            oenv.check_errors()
            del oenv

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cn, cT = args.cpu_sample_envs or (1 << 14), args.cpu_sample_T or 64
        v, times, cores = cpu_leg(args.env_id, cn, cT, 3)
        cpu = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
               "sample": "%d envs x %d env-steps, best of 3, C oracle port (oracle/minigrid_oracle.c), obs materialised"
                         % (cn, cT)}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": bench_config(args.env_id, N, T),
            "roofline": roofline, "cpu_baseline": cpu, "cpu_baseline_python": pyref, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "step_mode_env_steps_per_s_per_gpu": step_mode, "step_mode": step_mode_info, "other_configs_per_gpu": others,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
