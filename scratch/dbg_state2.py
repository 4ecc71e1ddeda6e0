import sys, numpy as np, torch, ctypes as C
sys.path.insert(0, '.')
import gym_minigrid_b200 as mgb
from gym_minigrid_b200 import _lib
n=3; W=H=8
env = mgb.make('MiniGrid-Empty-8x8-v0', num_envs=n, autoreset=False)
grid = torch.zeros((n,W,H,3), dtype=torch.uint8, device='cuda'); grid[...,0]=2; grid[...,1]=5
torch.cuda.synchronize()
print('grid sum', int(grid.sum()), hex(grid.data_ptr()))
L=_lib.load()
rc = L.mgb_set_state(env._h, 0, n, C.c_void_p(grid.data_ptr()), None, None, None, None, None, None, None)
torch.cuda.synchronize(); print('rc', rc)
s = env.get_state()
print(s['grid'][0,:,:,0])
f = C.c_uint32(0); L.mgb_error_flags(env._h, None, C.byref(f)); print('flags', f.value)
