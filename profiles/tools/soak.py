#!/usr/bin/env python
"""Long-horizon soak: GPU rollout vs the C oracle, bit for bit, N envs x STEPS env-steps per config (auto-reset on),
in chunks so that the outputs fit in memory.  Test infrastructure (uses oracle/); run on a GPU box:

    PYTHONPATH=. python profiles/tools/soak.py [N] [STEPS] [--scatter] [env ids ...]

--scatter: every env starts at its own step count, so that episode ends hit single lanes of a warp (spare layouts, in-register
resets and the lone-reset counters are exercised at scale instead of lock-step regeneration).
"""
import sys
import time

import numpy as np
import torch

import gym_minigrid_b200 as mgb
from oracle.oracle import OracleVec

SCATTER = "--scatter" in sys.argv
if SCATTER:
    sys.argv.remove("--scatter")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 15
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
CHUNK = 50
IDS = ["MiniGrid-Empty-8x8-v0", "MiniGrid-DoorKey-16x16-v0", "MiniGrid-FourRooms-v0", "MiniGrid-Dynamic-Obstacles-16x16-v0",
       "MiniGrid-KeyCorridorS6R3-v0", "MiniGrid-Dynamic-Obstacles-Random-6x6-v0", "MiniGrid-DoorKey-5x5-v0"]
for env_id in (sys.argv[3:] or IDS):
    cfg = {k: v for k, v in mgb.spec(env_id)["config"].items() if k not in ("mission", "reward_range")}
    env = mgb.make(env_id, num_envs=N, seed=2026, env_id_base=7_000_000_000)
    orc = OracleVec(cfg, N, seed=2026, env0=7_000_000_000)
    g0, o0 = env.reset(), orc.reset()
    assert np.array_equal(g0["image"].cpu().numpy(), o0[0])
    rs = np.random.RandomState(1)
    if SCATTER:
        so = orc.get_state()
        so["agent"][:, 3] = rs.randint(0, cfg["max_steps"] - 1, size=N)
        orc.set_state(so)
        env.set_state({"agent": torch.as_tensor(so["agent"])})
    t0, dones = time.time(), 0
    for c in range(STEPS // CHUNK):
        a = rs.randint(0, cfg["n_actions"], size=(CHUNK, N)).astype(np.uint8)
        o, r, dn, dr = env.rollout(torch.as_tensor(a))
        oo, orr, odn, odr = orc.rollout(a)
        ok = (np.array_equal(o.cpu().numpy(), oo) and np.array_equal(r.cpu().numpy().view(np.uint64), orr.view(np.uint64))
              and np.array_equal(dn.cpu().numpy().astype(np.uint8), odn) and np.array_equal(dr.cpu().numpy(), odr))
        if not ok:
            print("MISMATCH", env_id, "chunk", c)
            sys.exit(1)
        dones += int(odn.sum())
    env.check_errors()
    s = env.get_state()
    so = orc.get_state()
    for k in ("grid", "agent", "carrying"):
        assert np.array_equal(s[k].cpu().numpy(), so[k]), k
    print("%-44s %d envs x %d steps = %.2e env-steps bit-exact, %d episode ends  [%.0f s]" % (env_id, N, STEPS, N * STEPS, dones, time.time() - t0), flush=True)
print("soak ok")
