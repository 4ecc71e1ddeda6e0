"""CPU: level-pool mode -- the oracle against traces of the live reference stepping
reference-generated layouts of envs whose step() is the base MiniGridEnv.step (tests/golden/pool_*.npz)."""
import os

import numpy as np
import pytest

from helpers import assert_same, bits, golden_files, load
from oracle.oracle import OracleVec


@pytest.mark.parametrize("path", golden_files("pool_"), ids=os.path.basename)
def test_oracle_pool_matches_reference(path):
    d = load(path)
    K = d["level_grid"].shape[0]
    for k, idx in enumerate(d["env_indices"]):
        v = OracleVec(d["cfg"], 1, seed=int(d["seed"]), env0=int(idx), threads=1)
        v.set_level_pool(d["level_grid"], d["level_aux"], d["level_agent"])
        o0, d0 = v.reset()
        tag = "%s[%d]" % (os.path.basename(path), k)
        assert_same(tag + " obs0", o0[0], d["obs0"][k])
        assert d0[0] == d["dir0"][k]
        o, r, dn, dr = v.rollout(d["actions"][k].reshape(-1, 1), autoreset=True)
        assert_same(tag + " done", dn[:, 0], d["done"][k])
        assert_same(tag + " obs", o[:, 0], d["obs"][k])
        assert_same(tag + " dir", dr[:, 0], d["dir"][k])
        assert_same(tag + " reward bits", bits(r[:, 0]), bits(d["reward"][k]))
        s = v.get_state()
        assert_same(tag + " grid_end", s["grid"][0], d["grid_end"][k])
        assert_same(tag + " agent_end", s["agent"][0], d["agent_end"][k])
        n_ep = int((d["lvl"][k] >= 0).sum())
        assert int(s["rng"][0, 0]) == n_ep and 0 <= d["lvl"][k][:n_ep].max() < K
