#!/usr/bin/env python
"""Sustained throughput: back-to-back persistent rollouts for >= SECONDS per config (default 4 s), clocks sampled every
100 ms while the launches run (>= 20 samples), rate printed per config.  Complements bench.py's short timed region.

    python profiles/tools/sustained.py [--seconds 4] [--env-id ID ...]
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CONFIGS = ["MiniGrid-Empty-8x8-v0", "MiniGrid-DoorKey-16x16-v0", "MiniGrid-FourRooms-v0", "MiniGrid-Dynamic-Obstacles-16x16-v0",
           "MiniGrid-KeyCorridorS6R3-v0"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--env-id", nargs="*", default=CONFIGS)
    ap.add_argument("--num-envs", type=int, default=1 << 20)
    ap.add_argument("--rollout-T", type=int, default=32)
    a = ap.parse_args()
    import torch
    import gym_minigrid_b200 as mgb
    from gym_minigrid_b200 import _lib
    from bench import ClockSampler
    dev = torch.device("cuda", 0)
    N, T = a.num_envs, a.rollout_T
    out = (torch.empty((T, N, 7, 7, 3), dtype=torch.uint8, device=dev), torch.empty((T, N), dtype=torch.float64, device=dev),
           torch.empty((T, N), dtype=torch.uint8, device=dev), torch.empty((T, N), dtype=torch.uint8, device=dev))
    print("# library: %s" % _lib.load().mgb_version().decode())
    for env_id in a.env_id:
        cfg = mgb.spec(env_id)["config"]
        env = mgb.make(env_id, num_envs=N, device=dev, seed=0)
        env.reset()
        g = torch.Generator(device=dev).manual_seed(1234)
        acts = [torch.randint(0, cfg["n_actions"], (T, N), dtype=torch.uint8, device=dev, generator=g) for _ in range(4)]
        for i in range(3):
            env.rollout(acts[i], out=out)
        torch.cuda.synchronize(dev)
        sampler = ClockSampler(0)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches, t0 = 0, time.perf_counter()
        e0.record()
        while time.perf_counter() - t0 < a.seconds:
            for i in range(8):                                  # keep the queue full between clock checks
                env.rollout(acts[i % 4], out=out)
            launches += 8
            torch.cuda.synchronize(dev)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop()
        env.check_errors()
        rate = launches * N * T / (ms * 1e-3)
        print(json.dumps({"env_id": env_id, "env_steps_per_s": rate, "seconds": ms * 1e-3, "launches": launches,
                          "algorithmic_gbs": rate * 158 / 1e9, "clocks": clocks}))
        del env


if __name__ == "__main__":
    main()
