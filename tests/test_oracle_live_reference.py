"""CPU: live fuzz of the C oracle against the UNMODIFIED reference imported from /root/reference
(skipped where the tree is not mounted, e.g. on the GPU box).  Complements the committed fixtures with
fresh seeds: Philox injection for every registered env id, bit-exact obs/dir/reward/done + full grid."""
import numpy as np
import pytest

from helpers import assert_same, bits
from oracle import ref_shim
from oracle.oracle import OracleVec

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


def _cfg(env_id):
    import gym_minigrid_b200 as mgb
    return {k: v for k, v in mgb.spec(env_id)["config"].items() if k not in ("mission", "reward_range")}


def _ids():
    import gym_minigrid_b200 as mgb
    return [i for i in mgb.env_list if mgb.spec(i)["config"]["gen"] != 5]    # level-pool ids: tests/test_pool_*.py


@pytest.mark.parametrize("env_id", _ids())
def test_oracle_vs_live_reference(env_id):
    cfg = _cfg(env_id)
    seed, idx, T = 987654321, 1234567, 260
    env = ref_shim.make(env_id)
    shim = ref_shim.PhiloxShim(seed, idx, 0)
    env.np_random = shim
    obs = env.reset()
    orc = OracleVec(cfg, 1, seed=seed, env0=idx, threads=1)
    o0, d0 = orc.reset()
    assert_same(env_id + " obs0", o0[0], obs["image"])
    assert d0[0] == obs["direction"]
    import zlib
    rs = np.random.RandomState(zlib.crc32(env_id.encode()))
    acts = rs.randint(0, cfg["n_actions"], size=T).astype(np.uint8)
    want_o, want_r, want_d, want_dir = [], [], [], []
    ep = 1
    for t in range(T):
        obs, r, d, _ = env.step(int(acts[t]))
        if d:
            shim.new_episode(ep)
            obs = env.reset()
            ep += 1
        want_o.append(obs["image"]); want_r.append(float(r)); want_d.append(int(d)); want_dir.append(int(obs["direction"]))
    o, r, dn, dr = orc.rollout(acts.reshape(-1, 1), autoreset=True)
    assert_same(env_id + " done", dn[:, 0], np.array(want_d, np.uint8))
    assert_same(env_id + " obs", o[:, 0], np.stack(want_o))
    assert_same(env_id + " dir", dr[:, 0], np.array(want_dir, np.uint8))
    assert_same(env_id + " reward bits", bits(r[:, 0]), bits(np.array(want_r)))
    s = ref_shim.snapshot(env)
    so = orc.get_state()
    assert_same(env_id + " grid", so["grid"][0], s["grid"])
    assert_same(env_id + " agent", so["agent"][0], s["agent"])
