import sys, torch, gym_minigrid_b200 as mgb
env_id = sys.argv[1] if len(sys.argv) > 1 else "MiniGrid-Dynamic-Obstacles-16x16-v0"
N, T = 1 << 20, 32
for name, hi in (("random 0..2", 3), ("turn only 0..1", 2)):
    env = mgb.make(env_id, num_envs=N, seed=0)
    env.reset()
    acts = torch.randint(0, hi, (T, N), dtype=torch.uint8, device="cuda")
    out = None
    obs, r, d, di = env.rollout(acts)
    out = (obs, r.clone(), d.view(torch.uint8), di)
    for _ in range(3): env.rollout(acts, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): env.rollout(acts, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("%s  %-16s %.3e env-steps/s  done rate %.4f" % (env_id, name, N * T / ms * 1e3, float(out[2].float().mean())))
    del env
