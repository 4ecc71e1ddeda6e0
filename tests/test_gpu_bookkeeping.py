"""GPU: batched bookkeeping wrappers against the outputs of the reference's own wrapper classes
(tests/golden/bookkeeping_*.npz, made by oracle/gen_golden.py:bookkeeping_traces from the live reference).
uint8 outputs bit-exact; rewards (fp64: env reward + 1/sqrt(count)) compared by bit pattern."""
import os

import numpy as np
import pytest

from helpers import assert_same, bits, golden_files

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _np(t):
    return t.detach().cpu().numpy()


def _envs(z, **kw):
    import gym_minigrid_b200 as mgb
    env_id, seed = str(z["env_id"]), int(z["seed"])
    for k, idx in enumerate(z["env_indices"]):
        yield k, mgb.make(env_id, num_envs=1, seed=seed, env_id_base=int(idx), **kw)


@pytest.mark.parametrize("path", golden_files("bookkeeping_"), ids=os.path.basename)
def test_visit_bonus(path):
    from gym_minigrid_b200 import wrappers as W
    z = np.load(path)
    for tag, cls in (("ab", W.ActionBonus), ("sb", W.StateBonus)):
        for k, env in _envs(z):
            w = cls(env)
            assert not env.autoreset                       # the wrapper took auto-reset over
            w.reset()
            rr, dd = [], []
            for t in range(z["actions"].shape[1]):
                obs, r, d, _ = w.step(torch.as_tensor(z["actions"][k, t:t + 1]))
                rr.append(float(r[0])); dd.append(bool(d[0]))
                if tag == "ab":
                    assert_same("%s ActionBonus obs[%d]@%d" % (os.path.basename(path), k, t), _np(obs["image"])[0], z["ab_image"][k, t])
            assert_same(tag + " done", np.array(dd), z[tag + "_done"][k])
            assert_same(tag + " reward bits", bits(np.array(rr)), bits(z[tag + "_reward"][k]))
            env.check_errors()


@pytest.mark.parametrize("path", golden_files("bookkeeping_"), ids=os.path.basename)
def test_dac(path):
    from gym_minigrid_b200 import wrappers as W
    z = np.load(path)
    L = int(z["dac_len"])
    for k, env in _envs(z):
        w = W.DACWrapper(env)
        w.reset()
        for t in range(L):
            obs, r, d, _ = w.step(torch.as_tensor(z["actions"][k, t:t + 1]))
            tag = "%s DAC[%d]@%d" % (os.path.basename(path), k, t)
            assert_same(tag + " image", _np(obs["image"])[0], z["dac_image"][k, t])
            assert int(obs["direction"][0]) == int(z["dac_dir"][k, t]), tag
            assert bits(np.array([float(r[0])]))[0] == bits(z["dac_reward"][k, t:t + 1])[0], tag
            assert bool(d[0]) == bool(z["dac_done"][k, t]), tag


@pytest.mark.parametrize("path", golden_files("bookkeeping_"), ids=os.path.basename)
def test_append_action_goal_policy_extra_info(path):
    from gym_minigrid_b200 import wrappers as W
    z = np.load(path)
    T = z["actions"].shape[1]
    for k, env in _envs(z):
        w = W.AppendActionWrapper(W.FullyObsOneHotWrapper(W.ImgObsWrapper(W.FullyObsWrapper(env)), flatten=True), 3)
        assert w.observation_space.shape == (z["app_obs"].shape[2],)
        w.reset()
        for t in range(T):
            obs, _, _, _ = w.step(torch.as_tensor(z["actions"][k, t:t + 1]))
            assert_same("%s AppendAction[%d]@%d" % (os.path.basename(path), k, t), _np(obs)[0], z["app_obs"][k, t])
    for k, env in _envs(z):
        w = W.GoalPolicyWrapper(W.FullyObsOneHotWrapper(W.ImgObsWrapper(W.FullyObsWrapper(env)), flatten=True))
        xi = W.AgentExtraInfoWrapper(env)
        w.reset()
        for t in range(T):
            obs, _, _, _ = w.step(torch.as_tensor(z["actions"][k, t:t + 1]))
            tag = "%s GoalPolicy[%d]@%d " % (os.path.basename(path), k, t)
            assert_same(tag + "observation", _np(obs["observation"])[0], z["gp_obs"][k, t])
            assert_same(tag + "achieved", _np(obs["achieved_goal"])[0], z["gp_achieved"][k, t])
            assert_same(tag + "desired", _np(obs["desired_goal"])[0], z["gp_desired"][k, t])
            assert bits(_np(w.compute_reward())[:1])[0] == bits(z["gp_reward"][k, t:t + 1])[0], tag
            x = xi.observation({})
            assert_same(tag + "pos", _np(x["pos"])[0], z["xi_pos"][k, t])
            assert int(x["dir"][0]) == int(z["xi_dir"][k, t])
            assert_same(tag + "map", _np(xi.get_map())[0], z["xi_map"][k, t])
            assert_same(tag + "full map", _np(xi.get_full_map())[0], z["xi_full"][k, t])


def test_bookkeeping_batch_properties():
    """size-independent properties at batch scale (2^16 envs, ragged): the visit counts of an env sum to the number
    of steps taken; under DACWrapper every env reports done exactly once, at step max_steps, and blank observations
    from its first inner done on; appended action planes are one-hot."""
    import gym_minigrid_b200 as mgb
    from gym_minigrid_b200 import wrappers as W
    N, T = (1 << 16) + 5, 40
    g = torch.Generator().manual_seed(3)
    acts = torch.randint(0, 3, (T, N), dtype=torch.uint8, generator=g).cuda()
    w = W.StateBonus(mgb.make("MiniGrid-Dynamic-Obstacles-5x5-v0", num_envs=N, seed=1))
    w.reset()
    total = torch.zeros(N, dtype=torch.float64, device="cuda")
    for t in range(T):
        _, r, _, _ = w.step(acts[t])
        total += r
    assert torch.all(w.counts.sum(1) == T)
    assert torch.all(w.counts >= 0) and float(total.min()) > -T

    env = mgb.make("MiniGrid-Dynamic-Obstacles-5x5-v0", num_envs=N, seed=1)
    d = W.DACWrapper(env)
    d.reset()
    ndone = torch.zeros(N, dtype=torch.int32, device="cuda")
    first_inner = torch.full((N,), -1, dtype=torch.int32, device="cuda")
    for t in range(env.max_steps):
        a = torch.randint(0, 3, (N,), dtype=torch.uint8, generator=g).cuda()
        obs, r, done, _ = d.step(a)
        blank = (obs["image"].reshape(N, -1) == 1).all(1)
        assert torch.equal(blank, d.env_done)
        assert torch.all(r[blank & (first_inner >= 0)] == 0)
        first_inner = torch.where((first_inner < 0) & d.env_done, torch.full_like(first_inner, t), first_inner)
        ndone += done.int()
        assert bool(done.any()) == (t == env.max_steps - 1)
    assert torch.all(ndone == 1)

    ap = W.AppendActionWrapper(W.FullyObsOneHotWrapper(W.ImgObsWrapper(W.FullyObsWrapper(
        mgb.make("MiniGrid-Dynamic-Obstacles-5x5-v0", num_envs=N, seed=1))), flatten=True), 4)
    o = ap.reset()
    D = o.shape[1] - 3 * 4
    assert int(o[:, D:].sum()) == 0
    for t in range(5):
        o, _, done, _ = ap.step(acts[t])
        tail = o[:, D:].reshape(N, 4, 3)
        assert torch.all(tail.sum(2) <= 1)
        live = ~done
        assert torch.all(tail[live, 3].argmax(1) == acts[t][live].long()) and torch.all(tail[live, 3].sum(1) == 1)
        assert int(tail[done].sum()) == 0


def test_terminal_observation_wrapper():
    """TerminalObservation: same trajectory as the plain auto-resetting env, plus the reference's done-step observation
    (= what a non-auto-resetting twin returns) in info."""
    import gym_minigrid_b200 as mgb
    from gym_minigrid_b200 import wrappers as W
    N, T = 4099, 60
    g = torch.Generator().manual_seed(11)
    plain = mgb.make("MiniGrid-Dynamic-Obstacles-8x8-v0", num_envs=N, seed=5)
    w = W.TerminalObservation(mgb.make("MiniGrid-Dynamic-Obstacles-8x8-v0", num_envs=N, seed=5))
    o0, o1 = plain.reset(), w.reset()
    assert torch.equal(o0["image"], o1["image"])
    seen = 0
    for t in range(T):
        a = torch.randint(0, 3, (N,), dtype=torch.uint8, generator=g)
        twin_state = plain.get_state()
        p_obs, p_r, p_d, _ = plain.step(a)
        w_obs, w_r, w_d, info = w.step(a)
        assert torch.equal(p_obs["image"], w_obs["image"]) and torch.equal(p_obs["direction"], w_obs["direction"])
        assert torch.equal(p_r, w_r) and torch.equal(p_d, w_d)
        # a twin that does not auto-reset, started from the same pre-step state, shows the terminal observation
        twin = mgb.make("MiniGrid-Dynamic-Obstacles-8x8-v0", num_envs=N, seed=5, autoreset=False)
        twin.reset(); twin.set_state(twin_state)
        t_obs, _, t_d, _ = twin.step(a)
        assert torch.equal(t_d, w_d)
        term = info["terminal_observation"]
        assert torch.equal(term["image"], t_obs["image"]) and torch.equal(term["direction"], t_obs["direction"])
        seen += int(w_d.sum())
    assert seen > 100


def test_episode_statistics_wrapper():
    """EpisodeStatistics: totals at done steps equal a host-side accumulation of the same rewards."""
    import gym_minigrid_b200 as mgb
    from gym_minigrid_b200 import wrappers as W
    N, T = 5000, 80
    w = W.EpisodeStatistics(mgb.make("MiniGrid-Dynamic-Obstacles-6x6-v0", num_envs=N, seed=4))
    w.reset()
    g = torch.Generator().manual_seed(5)
    ret, ln = np.zeros(N), np.zeros(N, np.int64)
    eps = steps = 0
    for t in range(T):
        _, r, d, info = w.step(torch.randint(0, 3, (N,), dtype=torch.uint8, generator=g))
        r, d = r.cpu().numpy(), d.cpu().numpy()
        ret += r; ln += 1
        assert np.array_equal(info['episode']['r'].cpu().numpy(), np.where(d, ret, 0.0))
        assert np.array_equal(info['episode']['l'].cpu().numpy(), np.where(d, ln, 0))
        eps += int(d.sum()); steps += int(ln[d].sum())
        ret[d] = 0; ln[d] = 0
    assert w.episodes == eps and eps > 1000 and abs(w.mean_length() - steps / eps) < 1e-12
