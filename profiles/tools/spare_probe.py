import sys, os, ctypes as C
sys.path.insert(0, os.getcwd())
import torch, gym_minigrid_b200 as mgb
from gym_minigrid_b200 import _lib
L = _lib.load()
def counters(clear=True):
    a = (C.c_ulonglong * 8)()
    assert L.mgb_debug_spares(a, int(clear)) == 0
    return list(a)[:6]
env_id = sys.argv[1]
N, T = 1 << 20, 32
dev = torch.device("cuda", 0)
cfg = mgb.spec(env_id)["config"]
env = mgb.make(env_id, num_envs=N, device=dev, seed=0)
env.reset()
print("after reset", counters())
g = torch.Generator(device=dev).manual_seed(1234)
s = env.get_state(("agent",)); a = s["agent"]
a[:, 3] = torch.randint(0, cfg["max_steps"] - 1, (N,), device=dev, dtype=torch.int32, generator=g)
env.set_state({"agent": a})
for i in range(int(sys.argv[2]) if len(sys.argv) > 2 else 12):
    acts = torch.randint(0, cfg["n_actions"], (T, N), dtype=torch.uint8, device=dev, generator=g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    o = env.rollout(acts)
    e1.record()
    torch.cuda.synchronize()
    print(i, "%.3f ms" % e0.elapsed_time(e1), "dones", int(o[2].sum()), "calls,consumed,stale,passes,live,ahead =", counters())
