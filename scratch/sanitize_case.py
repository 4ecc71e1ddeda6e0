import sys, numpy as np, torch
sys.path.insert(0, '.')
import gym_minigrid_b200 as mgb
from gym_minigrid_b200 import wrappers as W
for env_id in ["MiniGrid-Empty-8x8-v0", "MiniGrid-DoorKey-16x16-v0", "MiniGrid-FourRooms-v0", "MiniGrid-Dynamic-Obstacles-16x16-v0", "MiniGrid-KeyCorridorS6R3-v0", "MiniGrid-Dynamic-Obstacles-Random-5x5-v0"]:
    N, T = 32 * 5 + 9, 24
    env = mgb.make(env_id, num_envs=N, seed=3)
    env.reset()
    a = torch.randint(0, env.action_space.n, (T, N), dtype=torch.uint8, device="cuda")
    o, r, d, dr = env.rollout(a)
    for t in range(4):
        env.step(a[t])
    s = env.get_state(); env.set_state(s)
    env.full_obs()
    W.FlatObsWrapper(env).observation(env.reset())
    W.OneHotPartialObsWrapper(env).observation(env.reset())
    env.step_host(a[0].cpu().pin_memory())
    env.check_errors()
    torch.cuda.synchronize()
    print("ok", env_id, int(d.sum()))
