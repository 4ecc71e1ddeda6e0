#!/usr/bin/env python
"""SURVEY §8(d) sweep: env-steps/s for N in {2^14, 2^17, 2^20, 2^22} envs per GPU, rollout mode (one
persistent launch of T steps) and step mode (T single-step launches), per config.  CUDA events, best of 5.

    python profiles/tools/sweep.py [env ids ...]  > profiles/r1_sweep.txt      (on a GPU box)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gym_minigrid_b200 as mgb  # noqa: E402

IDS = sys.argv[1:] or ["MiniGrid-Empty-8x8-v0", "MiniGrid-DoorKey-16x16-v0", "MiniGrid-FourRooms-v0",
                        "MiniGrid-Dynamic-Obstacles-16x16-v0", "MiniGrid-KeyCorridorS6R3-v0"]


def timed(fn, reps=5):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e-3


print("%-40s %9s %4s %14s %8s %14s" % ("config", "N", "T", "rollout st/s", "GB/s", "step-mode st/s"))
for env_id in IDS:
    for logn in (14, 17, 20, 22):
        N = 1 << logn
        T = 32 if logn <= 20 else 8                        # keep the [T,N,147] output buffer <= 5.3 GB
        env = mgb.make(env_id, num_envs=N, seed=0)
        env.reset()
        g = torch.Generator(device="cuda").manual_seed(1234)
        a = torch.randint(0, env.action_space.n, (T, N), dtype=torch.uint8, device="cuda", generator=g)
        out = (torch.empty((T, N, 7, 7, 3), dtype=torch.uint8, device="cuda"), torch.empty((T, N), dtype=torch.float64, device="cuda"),
               torch.empty((T, N), dtype=torch.uint8, device="cuda"), torch.empty((T, N), dtype=torch.uint8, device="cuda"))
        for _ in range(3):
            env.rollout(a, out=out)
        torch.cuda.synchronize()
        tr = timed(lambda: env.rollout(a, out=out))
        o1 = (out[0][0], out[1][0], out[2][0], out[3][0])

        def steps():
            for t in range(T):
                env.step(a[t], out=o1)
        steps()
        ts = timed(steps)
        print("%-40s %9d %4d %14.3e %8.0f %14.3e" % (env_id, N, T, N * T / tr, 158.0 * N * T / tr / 1e9, N * T / ts), flush=True)
        env.close()
        del out, a, env
        torch.cuda.empty_cache()
