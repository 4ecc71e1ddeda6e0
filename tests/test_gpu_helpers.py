"""GPU: batched MiniGridEnv geometry helpers (dir_vec, right_vec, front/left/right_pos, get_view_exts, get_view_coords,
in_view, agent_sees; minigrid.py:1092-1225) against the reference's own values for every cell of the grid along
Philox-injected trajectories (tests/golden/helpers_*.npz)."""
import os

import numpy as np
import pytest

from helpers import assert_same, golden_files

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("path", golden_files("helpers_"), ids=os.path.basename)
def test_geometry_helpers_match_reference(path):
    import gym_minigrid_b200 as mgb
    z = np.load(path)
    env_id, seed = str(z["env_id"]), int(z["seed"])
    W, H = z["in_view"].shape[2:4]
    ys = torch.arange(H).cuda()
    for k, idx in enumerate(z["env_indices"]):
        env = mgb.make(env_id, num_envs=1, seed=seed, env_id_base=int(idx))
        env.reset()
        T = z["actions"].shape[1]
        for t in range(T + 1):
            tag = "%s[%d]@%d " % (os.path.basename(path), k, t)
            for name in ("dir_vec", "right_vec", "front_pos", "left_pos", "right_pos"):
                assert_same(tag + name, _np(getattr(env, name))[0], z[name][k, t])
            assert_same(tag + "view_exts", _np(env.get_view_exts())[0], z["view_exts"][k, t])
            for x in range(W):
                vx, vy = env.get_view_coords(x, ys)          # a column of cells at once (broadcast over the batch of 1)
                assert_same(tag + "view_coords x", _np(vx), z["view_coords"][k, t, x, :, 0])
                assert_same(tag + "view_coords y", _np(vy), z["view_coords"][k, t, x, :, 1])
                assert_same(tag + "in_view", _np(env.in_view(x, ys)), z["in_view"][k, t, x])
                for y in range(H):
                    assert bool(env.agent_sees(x, y)[0]) == bool(z["agent_sees"][k, t, x, y]), tag + "agent_sees (%d,%d)" % (x, y)
            if t < T:
                env.step(torch.as_tensor(z["actions"][k, t:t + 1]))
