#!/usr/bin/env python
"""The reference's own benchmark.py (benchmark.py:22-53: reset time, render FPS, agent-view FPS =
step(0) under RGBImgPartialObsWrapper + ImgObsWrapper) on the batched GPU env.  CUDA events; frames = envs x steps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gym_minigrid_b200 as mgb  # noqa: E402
from gym_minigrid_b200 import wrappers as W  # noqa: E402

env_id = sys.argv[1] if len(sys.argv) > 1 else "MiniGrid-Empty-8x8-v0"      # benchmark.py default --env-name is LavaGapS7; Empty-8x8 is BASELINE configs[0]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18


def ev():
    return torch.cuda.Event(enable_timing=True)


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / reps


base = mgb.make(env_id, num_envs=N, seed=1337)
t_reset = timed(lambda: base.reset(), 20)
full = W.RGBImgObsWrapper(base)
obs = base.reset()
t_render = timed(lambda: full.observation(obs), 20)
env = W.ImgObsWrapper(W.RGBImgPartialObsWrapper(base))
zero = torch.zeros(N, dtype=torch.uint8, device="cuda")
env.reset()
t_view = timed(lambda: env.step(zero), 50)
print("env %s, %d envs" % (env_id, N))
print("Env reset time: %.4f ms per batch = %.2f ns per env   (reference: 0.2 ms per env)" % (t_reset * 1e3, t_reset / N * 1e9))
print("Rendering FPS : %.3e frames/s (full-grid RGB, %d B/frame, %.0f GB/s)   (reference: 886)" % (N / t_render, base.width * base.height * 192, N * base.width * base.height * 192 / t_render / 1e9))
print("Agent view FPS: %.3e frames/s (step + 56x56x3 RGB view, %.0f GB/s written)   (reference: 1233)" % (N / t_view, N * (9408 + 157) / t_view / 1e9))
