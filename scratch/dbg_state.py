import sys, numpy as np, torch
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from helpers import load
import gym_minigrid_b200 as mgb
d = load('tests/golden/scenes_occluded_8x8.npz')
n = d['actions'].shape[0]
env = mgb.make(d['env_id'], num_envs=n, autoreset=False)
env.set_state(dict(grid=d['grid0'], aux=d['aux0'], agent=d['agent0'], carrying=d['carrying0']))
s = env.get_state()
for k, w in (('grid','grid0'),('aux','aux0'),('agent','agent0'),('carrying','carrying0')):
    g = s[k].cpu().numpy(); 
    bad = np.argwhere(g != d[w])
    print(k, 'mismatches', len(bad), bad[:5].tolist())
    if len(bad): 
        i = tuple(bad[0]); print(' got', g[i], 'want', d[w][i])
obs = env.reset(mask=np.zeros(n, np.uint8))
o = obs['image'].cpu().numpy()
bad = np.argwhere(o != d['obs0']); print('obs0 mismatches', len(bad), 'envs', sorted(set(bad[:,0].tolist()))[:40])
i = bad[0][0]
print('agent', d['agent0'][i], 'carry', d['carrying0'][i])
print('got\n', o[i,:,:,0].T, '\nwant\n', d['obs0'][i,:,:,0].T)
print('grid\n', d['grid0'][i,:,:,0].T)
