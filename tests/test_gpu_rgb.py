"""GPU: RGB observation wrappers -- pixel-exact against images rendered by the reference's
own RGBImgPartialObsWrapper / RGBImgObsWrapper (tests/golden/rgb_*.npz).  The tile atlas the kernels gather
from was itself produced by the reference rasteriser (oracle/gen_atlas.py)."""
import os

import numpy as np
import pytest

from helpers import assert_same, golden_files

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("path", golden_files("rgb_"), ids=os.path.basename)
def test_rgb_wrappers_match_reference(path):
    import gym_minigrid_b200 as mgb
    from gym_minigrid_b200 import wrappers as W
    z = np.load(path)
    env_id, seed = str(z["env_id"]), int(z["seed"])
    for k, idx in enumerate(z["env_indices"]):
        env = mgb.make(env_id, num_envs=1, seed=seed, env_id_base=int(idx))
        wp, wf = W.RGBImgPartialObsWrapper(env), W.RGBImgObsWrapper(env)
        assert wp.observation_space["image"].shape == (56, 56, 3)
        assert wf.observation_space["image"].shape == (env.width * 8, env.height * 8, 3)
        obs = env.reset()
        T = z["actions"].shape[1]
        for t in range(T + 1):
            tag = "%s[%d]@%d" % (os.path.basename(path), k, t)
            assert_same(tag + " partial", wp.observation(obs)["image"].cpu().numpy()[0], z["partial"][k, t])
            assert_same(tag + " full", wf.observation(obs)["image"].cpu().numpy()[0], z["full"][k, t])
            # MiniGridEnv.render('rgb_array', highlight=True): the agent's visible cells highlighted
            assert_same(tag + " render(highlight)", env.render('rgb_array', highlight=True, tile_size=8).cpu().numpy()[0], z["full_highlight"][k, t])
            if t < T:
                obs, _, _, _ = env.step(torch.as_tensor(z["actions"][k, t:t + 1]))


def test_rgb_batch_consistency():
    """a batch renders exactly like its members one by one (coalesced indexing), also for view size 5"""
    import gym_minigrid_b200 as mgb
    from gym_minigrid_b200 import wrappers as W
    N = 700
    env = mgb.make("MiniGrid-DoorKey-16x16-v0", num_envs=N, seed=4, agent_view_size=5)
    wp, wf = W.RGBImgPartialObsWrapper(env), W.RGBImgObsWrapper(env)
    obs = env.reset()
    for _ in range(3):
        obs, _, _, _ = env.step(torch.randint(0, 7, (N,), dtype=torch.uint8))
    p, f = wp.observation(obs)["image"], wf.observation(obs)["image"]
    assert tuple(p.shape) == (N, 40, 40, 3) and tuple(f.shape) == (N, 128, 128, 3)
    atlas = wp._atlas.cpu().numpy()
    img = obs["image"].cpu().numpy()
    for n in (0, 1, 31, 32, 333, N - 1):
        want = np.zeros((40, 40, 3), np.uint8)
        for vx in range(5):
            for vy in range(5):
                t, c, s = img[n, vx, vy]
                variant = 6 if (vx, vy) == (2, 4) else (1 if t != 0 else 0)
                want[vy * 8:(vy + 1) * 8, vx * 8:(vx + 1) * 8] = atlas[t * 21 + c * 3 + s, variant]
        assert_same("partial env %d" % n, p[n].cpu().numpy(), want)
