import sys, numpy as np, torch, ctypes as C
sys.path.insert(0, '.'); sys.path.insert(0,'tests')
from helpers import load
import gym_minigrid_b200 as mgb
d = load('tests/golden/scenes_occluded_8x8.npz')
n = d['actions'].shape[0]
def chk(fields):
    env = mgb.make(d['env_id'], num_envs=n, autoreset=False)
    env.set_state({k: d[k+'0'] for k in fields})
    s = env.get_state()
    print(fields, 'grid mismatches', int((s['grid'].cpu().numpy() != d['grid0']).sum()), 'agent', int((s['agent'].cpu().numpy() != d['agent0']).sum()))
chk(['grid']); chk(['grid','aux']); chk(['grid','agent']); chk(['grid','carrying']); chk(['grid','aux','agent','carrying'])
env = mgb.make(d['env_id'], num_envs=n, autoreset=False)
g = torch.as_tensor(d['grid0']).cuda(); a = torch.as_tensor(d['agent0']).cuda(); torch.cuda.synchronize()
env.set_state(dict(grid=g, agent=a))
s = env.get_state(); print('device tensors: grid mismatches', int((s['grid'].cpu().numpy() != d['grid0']).sum()))
