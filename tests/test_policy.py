"""CPU: the counter-based uniform random policy (mgb_rollout_random / orc_vec_rollout_random) is what its header says:
Philox4x32-10 on (t>>2, rollout epoch, env id) under the tweaked key, mulhi32 to [0, n_actions) -- restated here
with the pure-Python Philox of oracle/ref_shim.py -- and a random-policy rollout equals an ordinary rollout fed with the
actions it reports."""
import numpy as np

from helpers import assert_same, bits
from oracle import ref_shim as R
from oracle.oracle import OracleVec, lib

ACTION_KEY = 0x41435431


def _py_action(seed, env_id, epoch, t, n):
    key = (seed & 0xFFFFFFFF, ((seed >> 32) & 0xFFFFFFFF) ^ ACTION_KEY)
    ctr = (t >> 2, epoch & 0xFFFFFFFF, env_id & 0xFFFFFFFF, (env_id >> 32) & 0xFFFFFFFF)
    return (R.philox4x32_10(ctr, key)[t & 3] * n) >> 32


def test_policy_stream_is_philox():
    L = lib()
    rs = np.random.RandomState(0)
    for _ in range(300):
        seed = int(rs.randint(0, 2 ** 62)); env_id = int(rs.randint(0, 2 ** 40)); ep = int(rs.randint(0, 1000))
        step = int(rs.randint(0, 5000)); n = int(rs.randint(1, 10))
        assert L.orc_policy_action(seed, env_id, ep, step, n) == _py_action(seed, env_id, ep, step, n)


def test_random_rollout_equals_replay():
    import gym_minigrid_b200 as mgb
    for env_id in ("MiniGrid-Empty-8x8-v0", "MiniGrid-Dynamic-Obstacles-6x6-v0", "MiniGrid-DoorKey-5x5-v0"):
        cfg = {k: v for k, v in mgb.spec(env_id)["config"].items() if k not in ("mission", "reward_range")}
        a_env, b_env = OracleVec(cfg, 37, seed=11, env0=5), OracleVec(cfg, 37, seed=11, env0=5)
        a_env.reset(); b_env.reset()
        o, r, dn, d, acts = a_env.rollout_random(60)
        assert acts.max() < cfg["n_actions"] and len(np.unique(acts)) == cfg["n_actions"]
        o2, r2, dn2, d2 = b_env.rollout(acts)
        assert_same(env_id + " obs", o, o2); assert_same(env_id + " done", dn, dn2); assert_same(env_id + " dir", d, d2)
        assert_same(env_id + " reward", bits(r), bits(r2))
        want = [_py_action(11, 5 + 3, 0, t, cfg["n_actions"]) for t in range(60)]          # first random rollout: epoch 0
        assert_same("python restatement", acts[:, 3], np.array(want, np.uint8))
        _, _, _, _, acts2 = a_env.rollout_random(8)                                           # second one: epoch 1
        assert_same("epoch 1", acts2[:, 3], np.array([_py_action(11, 8, 1, t, cfg["n_actions"]) for t in range(8)], np.uint8))
