"""agent_view_size other than 7 (minigrid.py:776,795; ViewSizeWrapper wrappers.py:579-608): traces of the
reference's own ViewSizeWrapper (tests/golden/viewsize*.npz) against the oracle (CPU) and the CUDA path (GPU)."""
import os

import numpy as np
import pytest

from helpers import assert_same, bits, golden_files, load
from oracle.oracle import OracleVec


@pytest.mark.parametrize("path", golden_files("viewsize"), ids=os.path.basename)
def test_oracle_viewsize_matches_reference(path):
    d = load(path)
    V = int(d["view"])
    for k, idx in enumerate(d["env_indices"]):
        v = OracleVec(d["cfg"], 1, seed=int(d["seed"]), env0=int(idx), threads=1)
        o0, _ = v.reset()
        assert o0.shape == (1, V, V, 3)
        assert_same("obs0", o0[0], d["obs0"][k])
        o, r, dn, dr = v.rollout(d["actions"][k].reshape(-1, 1), autoreset=True)
        assert_same("done", dn[:, 0], d["done"][k])
        assert_same("obs", o[:, 0], d["obs"][k])
        assert_same("dir", dr[:, 0], d["dir"][k])
        assert_same("reward bits", bits(r[:, 0]), bits(d["reward"][k]))


@pytest.mark.gpu
@pytest.mark.parametrize("path", golden_files("viewsize"), ids=os.path.basename)
def test_cuda_viewsize_matches_reference(path):
    import torch
    import gym_minigrid_b200 as mgb
    from gym_minigrid_b200 import wrappers as W
    d = load(path)
    V = int(d["view"])
    for k, idx in enumerate(d["env_indices"]):
        if k == 0:      # through make(..., agent_view_size=V)
            env = mgb.make(d["env_id"], num_envs=1, seed=int(d["seed"]), env_id_base=int(idx), agent_view_size=V)
        else:           # through the wrapper, like the reference
            env = W.ViewSizeWrapper(mgb.make(d["env_id"], num_envs=1, seed=int(d["seed"]), env_id_base=int(idx)), V)
            assert env.observation_space["image"].shape == (V, V, 3)
        obs = env.reset()
        assert tuple(obs["image"].shape) == (1, V, V, 3)
        assert_same("obs0", obs["image"].cpu().numpy()[0], d["obs0"][k])
        o, r, dn, dr = env.unwrapped.rollout(torch.as_tensor(d["actions"][k].reshape(-1, 1)))
        assert_same("done", dn.cpu().numpy()[:, 0].astype(np.uint8), d["done"][k])
        assert_same("obs", o.cpu().numpy()[:, 0], d["obs"][k])
        assert_same("dir", dr.cpu().numpy()[:, 0], d["dir"][k])
        assert_same("reward bits", bits(r.cpu().numpy()[:, 0].copy()), bits(d["reward"][k]))
        env.unwrapped.check_errors()


@pytest.mark.gpu
@pytest.mark.parametrize("env_id,V", [("MiniGrid-DoorKey-16x16-v0", 5), ("MiniGrid-Dynamic-Obstacles-16x16-v0", 3),
                                      ("MiniGrid-FourRooms-v0", 11), ("MiniGrid-Empty-8x8-v0", 9)])
def test_cuda_viewsize_matches_oracle_batch(env_id, V):
    """full groups (TMA store path), ragged tail, auto-reset, every view size's staging layout"""
    import torch
    import gym_minigrid_b200 as mgb
    cfg = {k: v for k, v in mgb.spec(env_id)["config"].items() if k not in ("mission", "reward_range")}
    cfg["view_size"] = V
    N, T, seed = 1024 + 5, 100, 21
    a = np.random.RandomState(4).randint(0, cfg["n_actions"], size=(T, N)).astype(np.uint8)
    env = mgb.make(env_id, num_envs=N, seed=seed, agent_view_size=V)
    orc = OracleVec(cfg, N, seed=seed)
    o0 = env.reset()
    w0, _ = orc.reset()
    assert_same("reset", o0["image"].cpu().numpy(), w0)
    o, r, dn, dr = env.rollout(torch.as_tensor(a))
    wo, wr, wdn, wdr = orc.rollout(a, autoreset=True)
    assert_same("obs", o.cpu().numpy(), wo)
    assert_same("done", dn.cpu().numpy().astype(np.uint8), wdn)
    assert_same("reward bits", bits(r.cpu().numpy()), bits(wr))
    ob, rr, dd, _ = env.step_host(torch.as_tensor(a[0]).pin_memory())
    assert tuple(ob["image"].shape) == (N, V, V, 3)
    env.check_errors()
