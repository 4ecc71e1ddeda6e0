#!/usr/bin/env python
"""Duration of every rollout launch (T = 32) from reset onwards: shows how throughput moves as episodes age and reach
their time limits (the short bench of an id whose max_steps exceeds the benchmarked steps never sees a reset).

    python profiles/tools/per_launch.py ENV_ID [launches]
"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import gym_minigrid_b200 as mgb

env_id, L = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 100
N, T = 1 << 20, 32
dev = torch.device("cuda", 0)
cfg = mgb.spec(env_id)["config"]
env = mgb.make(env_id, num_envs=N, device=dev, seed=0)
env.reset()
g = torch.Generator(device=dev).manual_seed(1234)
acts = [torch.randint(0, cfg["n_actions"], (T, N), dtype=torch.uint8, device=dev, generator=g) for _ in range(4)]
out = (torch.empty((T, N, 7, 7, 3), dtype=torch.uint8, device=dev), torch.empty((T, N), dtype=torch.float64, device=dev),
       torch.empty((T, N), dtype=torch.uint8, device=dev), torch.empty((T, N), dtype=torch.uint8, device=dev))
evs = [torch.cuda.Event(enable_timing=True) for _ in range(L + 1)]
dones = []
evs[0].record()
for i in range(L):
    env.rollout(acts[i % 4], out=out)
    evs[i + 1].record()
    dones.append(out[2].sum())
torch.cuda.synchronize()
print("# %s  max_steps=%d  N=2^20 T=32; launch: first step, ms, env-steps/s, done-steps in the launch" % (env_id, cfg["max_steps"]))
for i in range(L):
    ms = evs[i].elapsed_time(evs[i + 1])
    print("%4d %6d %8.3f %.3e %9d" % (i, i * T, ms, N * T / ms / 1e-3, int(dones[i])))
env.check_errors()
