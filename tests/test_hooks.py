"""Level-pool envs with step() hooks (include/mgb200.h MGB_HOOK_*): Unlock, UnlockPickup,
BlockedUnlockPickup, Fetch, GoToDoor, GoToObject, PutNear, RedBlueDoors, Memory (+ LockedRoom, Playground with
the base step).  tests/golden/hook_*.npz hold, per env id, reference-generated levels with their hook attributes,
random-action traces, and directed scenarios that force every success / failure branch of the hook.
CPU: the oracle; GPU: the CUDA path through the C-ABI.  Bit-exact, rewards as fp64 bit patterns."""
import os

import numpy as np
import pytest

from helpers import assert_same, bits, golden_files, load
from oracle.oracle import OracleVec


def _check_random(d, run):
    for k, idx in enumerate(d["env_indices"]):
        o0, d0, o, r, dn, dr, grid_end, agent_end, last_level = run(k, int(idx))
        tag = "%s[%d]" % (os.path.basename(d["path"]), k)
        assert_same(tag + " obs0", o0, d["obs0"][k])
        assert int(d0) == int(d["dir0"][k])
        assert_same(tag + " done", dn, d["done"][k])
        assert_same(tag + " obs", o, d["obs"][k])
        assert_same(tag + " dir", dr, d["dir"][k])
        assert_same(tag + " reward bits", bits(r), bits(d["reward"][k]))
        assert_same(tag + " grid_end", grid_end, d["grid_end"][k])
        assert_same(tag + " agent_end", agent_end, d["agent_end"][k])
        n_ep = int((d["lvl"][k] >= 0).sum())
        assert int(last_level) == int(d["lvl"][k][n_ep - 1])


def _check_scenarios(d, run_sc):
    n = len(d["sc_len"])
    if n == 0:
        return
    o, r, dn, dr = run_sc()
    for i in range(n):
        L = int(d["sc_len"][i])
        tag = "%s scenario %d (level %d)" % (os.path.basename(d["path"]), i, int(d["sc_level"][i]))
        assert_same(tag + " done", dn[:L, i], d["sc_done"][i][:L])
        assert_same(tag + " reward bits", bits(r[:L, i].copy()), bits(d["sc_reward"][i][:L]))
        assert_same(tag + " obs", o[:L, i], d["sc_obs"][i][:L])
        assert_same(tag + " dir", dr[:L, i], d["sc_dir"][i][:L])


@pytest.mark.parametrize("path", golden_files("hook_"), ids=os.path.basename)
def test_oracle_hooks_match_reference(path):
    d = load(path)
    d["path"] = path
    hp = d["level_hook"] if d["cfg"]["hook"] else None

    def run(k, idx):
        v = OracleVec(d["cfg"], 1, seed=int(d["seed"]), env0=idx, threads=1)
        v.set_level_pool(d["level_grid"], d["level_aux"], d["level_agent"], hp)
        o0, d0 = v.reset()
        o, r, dn, dr = v.rollout(d["actions"][k].reshape(-1, 1), autoreset=True)
        s = v.get_state()
        return o0[0], d0[0], o[:, 0], r[:, 0], dn[:, 0], dr[:, 0], s["grid"][0], s["agent"][0], v.get_levels()[0]

    _check_random(d, run)

    def run_sc():
        n = len(d["sc_len"])
        v = OracleVec(d["cfg"], n, threads=1)
        v.set_level_pool(d["level_grid"], d["level_aux"], d["level_agent"], hp)
        v.set_state(dict(grid=d["sc_grid"], aux=d["sc_aux"], agent=d["sc_agent"], carrying=d["sc_carrying"]))
        v.set_levels(d["sc_level"])
        return v.rollout(d["sc_actions"].T.copy(), autoreset=False)

    _check_scenarios(d, run_sc)


@pytest.mark.gpu
@pytest.mark.parametrize("path", golden_files("hook_"), ids=os.path.basename)
def test_cuda_hooks_match_reference(path):
    import torch
    import gym_minigrid_b200 as mgb
    d = load(path)
    d["path"] = path
    for key, val in d["cfg"].items():
        if key not in ("view_size", "room_size", "num_rows"):     # RoomGrid attributes are irrelevant without a device generator
            assert int(mgb.spec(d["env_id"])["config"].get(key, 0)) == val, key
    levels = dict(grid=d["level_grid"], aux=d["level_aux"], agent=d["level_agent"], missions=[str(m) for m in d["level_mission"]],
                  hook_params=d["level_hook"] if d["cfg"]["hook"] else None)

    def run(k, idx):
        env = mgb.make(d["env_id"], num_envs=1, seed=int(d["seed"]), env_id_base=idx, levels=levels)
        obs = env.reset()
        assert obs["mission"][0] == str(d["level_mission"][d["lvl"][k][0]])
        o0, d0 = obs["image"].cpu().numpy()[0].copy(), int(obs["direction"][0])
        o, r, dn, dr = env.rollout(torch.as_tensor(d["actions"][k].reshape(-1, 1)))
        s = env.get_state()
        env.check_errors()
        return (o0, d0, o.cpu().numpy()[:, 0], r.cpu().numpy()[:, 0].copy(), dn.cpu().numpy()[:, 0].astype(np.uint8), dr.cpu().numpy()[:, 0],
                s["grid"].cpu().numpy()[0], s["agent"].cpu().numpy()[0], env.level_index()[0])

    _check_random(d, run)

    def run_sc():
        n = len(d["sc_len"])
        env = mgb.make(d["env_id"], num_envs=n, autoreset=False, levels=levels)
        env.set_state(dict(grid=d["sc_grid"], aux=d["sc_aux"], agent=d["sc_agent"], carrying=d["sc_carrying"]))
        env.set_levels(d["sc_level"])
        o, r, dn, dr = env.rollout(torch.as_tensor(d["sc_actions"].T.copy()))
        env.check_errors()
        return o.cpu().numpy(), r.cpu().numpy(), dn.cpu().numpy().astype(np.uint8), dr.cpu().numpy()

    _check_scenarios(d, run_sc)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["hook_fetch_8x8_n3.npz", "hook_putnear_8x8_n3.npz", "hook_gotoobject_8x8_n2.npz", "hook_memorys7.npz",
                                  "hook_redbluedoors_6x6.npz", "hook_gotodoor_6x6.npz"])
def test_cuda_hooks_match_oracle_batch(name):
    """thousands of envs x random actions: hooks fire often (fetch/putnear/gotoobject end episodes every few
    steps), auto-reset picks new levels, everything against the oracle"""
    import torch
    import gym_minigrid_b200 as mgb
    d = load(os.path.join(os.path.dirname(__file__), "golden", name))
    hp = d["level_hook"]
    N, T, seed, base = 2048 + 9, 150, 12, 5
    a = np.random.RandomState(6).randint(0, 7, size=(T, N)).astype(np.uint8)
    env = mgb.make(d["env_id"], num_envs=N, seed=seed, env_id_base=base,
                   levels=dict(grid=d["level_grid"], aux=d["level_aux"], agent=d["level_agent"], hook_params=hp))
    orc = OracleVec(d["cfg"], N, seed=seed, env0=base)
    orc.set_level_pool(d["level_grid"], d["level_aux"], d["level_agent"], hp)
    o0 = env.reset()
    w0, _ = orc.reset()
    assert_same("reset", o0["image"].cpu().numpy(), w0)
    o, r, dn, dr = env.rollout(torch.as_tensor(a))
    wo, wr, wdn, wdr = orc.rollout(a, autoreset=True)
    assert_same("done", dn.cpu().numpy().astype(np.uint8), wdn)
    assert_same("reward bits", bits(r.cpu().numpy()), bits(wr))
    assert_same("obs", o.cpu().numpy(), wo)
    assert_same("dir", dr.cpu().numpy(), wdr)
    assert_same("levels", env.level_index(), orc.get_levels().astype(np.int64))
    assert int((wr > 0).sum()) >= 0 and int(wdn.sum()) > 0
    env.check_errors()
