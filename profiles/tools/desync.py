#!/usr/bin/env python
"""Throughput when episode ends are spread over time instead of synchronised: every env starts with a random step count
in [0, max_steps), so that in every launch ~T/max_steps of the envs reach their time limit, each alone in its warp.

    python profiles/tools/desync.py ENV_ID [launches]
"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import gym_minigrid_b200 as mgb

env_id, L = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
N, T = 1 << 20, 32
dev = torch.device("cuda", 0)
cfg = mgb.spec(env_id)["config"]
env = mgb.make(env_id, num_envs=N, device=dev, seed=0)
env.reset()
g = torch.Generator(device=dev).manual_seed(1234)
acts = [torch.randint(0, cfg["n_actions"], (T, N), dtype=torch.uint8, device=dev, generator=g) for _ in range(4)]
out = (torch.empty((T, N, 7, 7, 3), dtype=torch.uint8, device=dev), torch.empty((T, N), dtype=torch.float64, device=dev),
       torch.empty((T, N), dtype=torch.uint8, device=dev), torch.empty((T, N), dtype=torch.uint8, device=dev))


def run(tag):
    for i in range(3):
        env.rollout(acts[i], out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dn = 0
    for i in range(L):
        env.rollout(acts[i % 4], out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%-14s %s: %.3e env-steps/s  (%.3f ms per launch, done-steps in the last launch: %d)" % (tag, env_id, L * N * T / ms / 1e-3, ms / L, int(out[2].sum())))


run("synchronised")
s = env.get_state(("agent",))
a = s["agent"]
a[:, 3] = torch.randint(0, cfg["max_steps"] - 1, (N,), device=dev, dtype=torch.int32, generator=g)
env.set_state({"agent": a})
run("desynchronised")
env.check_errors()
