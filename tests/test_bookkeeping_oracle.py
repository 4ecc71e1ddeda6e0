"""CPU: the bookkeeping-wrapper fixtures (tests/golden/bookkeeping_*.npz, produced by the reference's own
ActionBonus / StateBonus / DACWrapper / AgentExtraInfoWrapper) are consistent with the C oracle's trajectories: the
wrapper arithmetic restated in numpy on top of oracle steps reproduces the reference's rewards bit for bit.  This pins
the oracle on the terminal-state semantics the GPU wrappers rely on (auto-reset off, masked reset afterwards)."""
import math
import os

import numpy as np
import pytest

from helpers import assert_same, bits, golden_files
from oracle.oracle import OracleVec


def _cfg(env_id):
    import gym_minigrid_b200 as mgb
    return {k: v for k, v in mgb.spec(env_id)["config"].items() if k not in ("mission", "reward_range")}


@pytest.mark.parametrize("path", golden_files("bookkeeping_"), ids=os.path.basename)
def test_bonus_and_dac_restated_on_oracle(path):
    z = np.load(path)
    cfg = _cfg(str(z["env_id"]))
    T = z["actions"].shape[1]
    for k, idx in enumerate(z["env_indices"]):
        for tag in ("ab", "sb"):
            o = OracleVec(cfg, 1, seed=int(z["seed"]), env0=int(idx))
            o.reset()
            counts, rr = {}, []
            for t in range(T):
                a = int(z["actions"][k, t])
                _, r, d, _ = o.step(np.array([a], np.uint8), autoreset=False)
                ag = o.get_state()["agent"][0]
                key = (int(ag[0]), int(ag[1]), int(ag[2]), a) if tag == "ab" else (int(ag[0]), int(ag[1]))
                counts[key] = counts.get(key, 0) + 1
                rr.append(float(r[0]) + 1 / math.sqrt(counts[key]))
                assert bool(d[0]) == bool(z[tag + "_done"][k, t])
                if d[0]:
                    o.reset(np.array([1], np.uint8))
            assert_same(tag + " reward bits", bits(np.array(rr)), bits(z[tag + "_reward"][k]))
        # DACWrapper
        o = OracleVec(cfg, 1, seed=int(z["seed"]), env0=int(idx))
        obs0 = o.reset()
        dir0, env_done = int(obs0[1][0]), False
        for t in range(int(z["dac_len"])):
            img, r, d, di = o.step(z["actions"][k, t:t + 1].astype(np.uint8), autoreset=False)
            was = env_done
            env_done = env_done or bool(d[0])
            want_img = np.ones_like(img[0]) if env_done else img[0]
            assert_same("dac image@%d" % t, want_img.reshape(z["dac_image"][k, t].shape), z["dac_image"][k, t])
            assert (dir0 if env_done else int(di[0])) == int(z["dac_dir"][k, t])
            assert bits(np.array([0.0 if was else float(r[0])]))[0] == bits(z["dac_reward"][k, t:t + 1])[0]
            assert (env_done and t + 1 >= cfg["max_steps"]) == bool(z["dac_done"][k, t])
