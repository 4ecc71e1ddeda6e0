"""GPU: batched observation wrappers against outputs of the reference's own
wrapper classes (tests/golden/wrappers_*.npz, made by oracle/gen_golden.py from the live reference).
uint8 outputs bit-exact; FlatObsWrapper is float32 holding small integers -> compared exactly (0 ulp)."""
import os

import numpy as np
import pytest

from helpers import assert_same, golden_files

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("path", golden_files("wrappers_"), ids=os.path.basename)
def test_wrappers_match_reference(path):
    import gym_minigrid_b200 as mgb
    from gym_minigrid_b200 import wrappers as W
    z = np.load(path)
    env_id, seed = str(z["env_id"]), int(z["seed"])
    for k, idx in enumerate(z["env_indices"]):
        env = mgb.make(env_id, num_envs=1, seed=seed, env_id_base=int(idx))
        w_full = W.FullyObsWrapper(env)
        w_flat = W.FlatObsWrapper(env)
        w_foh = W.FullyObsOneHotWrapper(W.ImgObsWrapper(W.FullyObsWrapper(env)), flatten=True)
        w_img = W.ImgObsWrapper(env)
        assert w_full.observation_space["image"].shape == (env.width, env.height, 3)
        assert w_flat.observation_space.shape == (1, 147 + 27 * 96)
        assert w_img.observation_space.shape == (7, 7, 3)
        env.reset()
        T = z["actions"].shape[1]
        for t in range(T):
            obs, r, d, _ = env.step(torch.as_tensor(z["actions"][k, t:t + 1]))
            tag = "%s[%d]@%d" % (os.path.basename(path), k, t)
            fo = w_full.observation(obs)["image"]
            assert_same(tag + " FullyObsWrapper", _np(fo)[0], z["full"][k, t])
            fl = w_flat.observation(obs)
            assert fl.dtype == torch.float32
            assert_same(tag + " FlatObsWrapper", _np(fl)[0], z["flat"][k, t])
            assert_same(tag + " FullyObsOneHotWrapper", _np(w_foh.observation(fo))[0], z["full_onehot"][k, t])
            assert obs["mission"][0] == str(z["missions"][k, t])
            assert torch.equal(w_img.observation(obs), obs["image"])


def test_onehot_partial_and_reseed():
    """OneHotPartialObsWrapper: the reference class cannot run (wrappers.py:228 reads .shape of a Dict
    space), so the check is the literal loop body of wrappers.py:230-238 restated in numpy.
    ReseedWrapper (wrappers.py:12-32): same seed list => same layouts after every reset."""
    import gym_minigrid_b200 as mgb
    from gym_minigrid_b200 import wrappers as W
    env = W.OneHotPartialObsWrapper(mgb.make("MiniGrid-KeyCorridorS6R3-v0", num_envs=257, seed=9))
    assert env.observation_space["image"].shape == (7, 7, 21)
    obs = env.reset()
    a = torch.randint(0, 7, (257,), dtype=torch.uint8)
    for _ in range(5):
        obs, _, _, _ = env.step(a)
    img = _np(env.unwrapped.gen_obs()["image"])      # the raw observation of the current state
    want = np.zeros((257, 7, 7, 21), np.uint8)
    n, i, j = np.meshgrid(np.arange(257), np.arange(7), np.arange(7), indexing="ij")
    want[n, i, j, img[..., 0]] = 1
    want[n, i, j, 11 + img[..., 1]] = 1
    want[n, i, j, 11 + 7 + img[..., 2]] = 1
    assert_same("OneHotPartialObsWrapper", _np(obs["image"]), want)

    rs = W.ReseedWrapper(mgb.make("MiniGrid-FourRooms-v0", num_envs=64, seed=1), seeds=[5, 6])
    first = [rs.reset()["image"].clone() for _ in range(4)]
    assert torch.equal(first[0], first[2]) and torch.equal(first[1], first[3])
    assert not torch.equal(first[0], first[1])
    # the layouts equal those of a plain env created with that seed
    plain = mgb.make("MiniGrid-FourRooms-v0", num_envs=64, seed=5)
    assert torch.equal(plain.reset()["image"], first[0])


def test_wrapper_kernels_full_size():
    """size-independent properties at batch scale: every one-hot cell has exactly 3 bits set
    (2 when colours are dropped), flat obs = image ++ constant mission block."""
    import gym_minigrid_b200 as mgb
    from gym_minigrid_b200 import wrappers as W
    N = 1 << 16
    base = mgb.make("MiniGrid-DoorKey-16x16-v0", num_envs=N, seed=3)
    oh = W.OneHotPartialObsWrapper(base)
    obs = oh.reset()
    assert bool((obs["image"].sum(-1) == 3).all())
    foh = W.FullyObsOneHotWrapper(W.ImgObsWrapper(W.FullyObsWrapper(base)), drop_color=True, flatten=False)
    o = foh.observation(base.full_obs())
    assert o.shape == (N, 16, 16, 11 + 4) and bool((o.sum(-1) == 2).all())
    fl = W.FlatObsWrapper(base)
    f = fl.observation(base.reset())
    assert torch.equal(f[:, :147].to(torch.uint8).reshape(N, 7, 7, 3), base.gen_obs()["image"])
    assert bool((f[:, 147:] == f[0, 147:]).all()) and float(f[0, 147:].sum()) == len(base._mission)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["hook_fetch_8x8_n3.npz", "hook_gotodoor_6x6.npz", "hook_putnear_8x8_n3.npz", "hook_gotoobject_8x8_n2.npz"])
def test_flat_obs_encodes_each_levels_own_mission(name):
    """FlatObsWrapper on a level-pool id must one-hot the mission of the level each env is playing
    (fetch.py:63-72 etc. name the target), not the registry placeholder.  level_flat = the reference's own
    FlatObsWrapper output on each level's first observation."""
    import os
    import numpy as np
    import torch
    import gym_minigrid_b200 as mgb
    from gym_minigrid_b200 import wrappers as W
    from helpers import load, assert_same
    d = load(os.path.join(os.path.dirname(__file__), "golden", name))
    levels = dict(grid=d["level_grid"], aux=d["level_aux"], agent=d["level_agent"], missions=[str(m) for m in d["level_mission"]],
                  hook_params=d["level_hook"] if d["cfg"]["hook"] else None)
    N = 200
    env = W.FlatObsWrapper(mgb.make(d["env_id"], num_envs=N, seed=5, levels=levels))
    flat = env.reset()
    for step in range(6):
        lv = env.unwrapped.level_index()
        got = flat.cpu().numpy()
        assert got.shape == (N, 147 + 27 * 96) and got.dtype == np.float32
        assert_same("mission one-hot", got[:, 147:], d["level_flat"][lv][:, 147:])
        if step == 0:        # at reset the image part is the level's first observation too
            assert_same("image part", got[:, :147], d["level_flat"][lv][:, :147])
        assert len(set(lv.tolist())) > 1
        flat, _, _, _ = env.step(torch.randint(0, 7, (N,), dtype=torch.uint8))
    # without missions the wrapper refuses instead of encoding the placeholder
    bare = mgb.make(d["env_id"], num_envs=4, levels={k: v for k, v in levels.items() if k != "missions"})
    with pytest.raises(mgb.MgbError):
        W.FlatObsWrapper(bare)
