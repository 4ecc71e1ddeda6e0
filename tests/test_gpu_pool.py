"""GPU: level-pool mode through the C-ABI against the reference traces (bit-exact) and against the
oracle on a batch; missions follow the level each env is playing."""
import os

import numpy as np
import pytest

from helpers import assert_same, bits, golden_files, load

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _np(t):
    return t.detach().cpu().numpy()


def _levels(d):
    return dict(grid=d["level_grid"], aux=d["level_aux"], agent=d["level_agent"], missions=[str(m) for m in d["level_mission"]])


@pytest.mark.parametrize("path", golden_files("pool_"), ids=os.path.basename)
def test_cuda_pool_matches_reference(path):
    import gym_minigrid_b200 as mgb
    d = load(path)
    for k, idx in enumerate(d["env_indices"]):
        env = mgb.make(d["env_id"], num_envs=1, seed=int(d["seed"]), env_id_base=int(idx), levels=_levels(d))
        for key, val in d["cfg"].items():          # ids with a device generator run in pool mode when levels= is given
            assert int(env._cfg[key]) == val, key
        tag = "%s[%d]" % (os.path.basename(path), k)
        obs = env.reset()
        assert_same(tag + " obs0", _np(obs["image"])[0], d["obs0"][k])
        assert int(obs["direction"][0]) == int(d["dir0"][k])
        assert int(env.level_index()[0]) == int(d["lvl"][k][0])
        assert obs["mission"][0] == str(d["level_mission"][d["lvl"][k][0]])
        o, r, dn, dr = env.rollout(torch.as_tensor(d["actions"][k].reshape(-1, 1)))
        assert_same(tag + " done", _np(dn)[:, 0].astype(np.uint8), d["done"][k])
        assert_same(tag + " obs", _np(o)[:, 0], d["obs"][k])
        assert_same(tag + " dir", _np(dr)[:, 0], d["dir"][k])
        assert_same(tag + " reward bits", bits(_np(r)[:, 0].copy()), bits(d["reward"][k]))
        s = env.get_state()
        assert_same(tag + " grid_end", _np(s["grid"])[0], d["grid_end"][k])
        assert_same(tag + " agent_end", _np(s["agent"])[0], d["agent_end"][k])
        n_ep = int((d["lvl"][k] >= 0).sum())
        assert int(env.level_index()[0]) == int(d["lvl"][k][n_ep - 1])
        env.check_errors()


@pytest.mark.parametrize("name", ["pool_multiroom_n4_s5.npz", "pool_lavagaps7_v1.npz", "pool_simpleroom.npz"])
def test_cuda_pool_matches_oracle_batch(name):
    from oracle.oracle import OracleVec
    import gym_minigrid_b200 as mgb
    d = load(os.path.join(os.path.dirname(__file__), "golden", name))
    N, T, seed, base = 1024 + 13, 200, 5, 77
    actions = np.random.RandomState(2).randint(0, 7, size=(T, N)).astype(np.uint8)
    env = mgb.make(d["env_id"], num_envs=N, seed=seed, env_id_base=base, levels=_levels(d))
    orc = OracleVec(d["cfg"], N, seed=seed, env0=base)
    orc.set_level_pool(d["level_grid"], d["level_aux"], d["level_agent"])
    o0 = env.reset()
    w0, wd0 = orc.reset()
    assert_same("reset obs", _np(o0["image"]), w0)
    o, r, dn, dr = env.rollout(torch.as_tensor(actions))
    wo, wr, wdn, wdr = orc.rollout(actions, autoreset=True)
    assert_same("done", _np(dn).astype(np.uint8), wdn)
    assert_same("obs", _np(o), wo)
    assert_same("dir", _np(dr), wdr)
    assert_same("reward bits", bits(_np(r)), bits(wr))
    s, so = env.get_state(), orc.get_state()
    assert_same("grid", _np(s["grid"]), so["grid"])
    assert_same("rng", _np(s["rng"]).view(np.uint32), so["rng"])
    env.check_errors()


def test_pool_requires_levels():
    import gym_minigrid_b200 as mgb
    with pytest.raises(ValueError, match="level-pool"):
        mgb.make("MiniGrid-SimpleRoom-v0", num_envs=4)
    with pytest.raises(ValueError, match="only for level-pool"):
        mgb.make("MiniGrid-Empty-8x8-v0", num_envs=4, levels=dict(grid=np.zeros((1, 8, 8, 3), np.uint8), agent=np.zeros((1, 3))))
