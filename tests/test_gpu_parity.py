"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C-ABI
(gym_minigrid_b200 -> ctypes -> libmgb200.so); the CPU oracle and the committed golden traces of
the live reference are only the checkers.  Integer/byte outputs must be bit-exact; rewards are
compared as fp64 bit patterns (tolerance: 0 ulp)."""
import os

import numpy as np
import pytest

from helpers import assert_same, bits, golden_files, load

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

HEADLINE = ["MiniGrid-Empty-8x8-v0", "MiniGrid-DoorKey-16x16-v0", "MiniGrid-FourRooms-v0",
            "MiniGrid-Dynamic-Obstacles-16x16-v0", "MiniGrid-KeyCorridorS6R3-v0"]


def _mgb():
    import gym_minigrid_b200 as mgb
    return mgb


def _np(t):
    return t.detach().cpu().numpy()


def _oracle_cfg(env_id):
    cfg = dict(_mgb().spec(env_id)["config"])
    cfg.pop("mission", None)
    cfg.pop("reward_range", None)
    return cfg


def _check_trace_gpu(d, k, mode):
    mgb = _mgb()
    T = d["actions"].shape[1]
    env = mgb.make(d["env_id"], num_envs=1, seed=int(d["seed"]), env_id_base=int(d["env_indices"][k]))
    # the product's own config table must agree with what the reference env reported
    for key, val in d["cfg"].items():
        assert int(env._cfg[key]) == val, (key, env._cfg[key], val)
    if mode == "tape":
        lo, hi = int(d["tape_offsets"][k]), int(d["tape_offsets"][k + 1])
        env.set_rng_tape(d["tape"][lo:hi], np.array([0, hi - lo]))
    tag = "%s[%d]" % (os.path.basename(d["path"]), k)
    obs = env.reset()
    assert_same(tag + " obs0", _np(obs["image"])[0], d["obs0"][k])
    assert int(obs["direction"][0]) == int(d["dir0"][k])
    assert obs["mission"][0] == str(d["missions0"][k])
    s = env.get_state()
    assert_same(tag + " grid0", _np(s["grid"])[0], d["grid0"][k])
    assert_same(tag + " agent0", _np(s["agent"])[0], d["agent0"][k])
    t0 = 0
    nob = d["cfg"]["n_obstacles"]
    for si, ts in enumerate(d["snap_t"][k]):
        a = torch.as_tensor(d["actions"][k, t0:ts + 1].reshape(-1, 1))
        if si % 2 == 0:      # alternate the persistent rollout kernel and single-step launches
            o, r, dn, dr = env.rollout(a)
            o, r, dn, dr = _np(o)[:, 0], _np(r)[:, 0], _np(dn)[:, 0], _np(dr)[:, 0]
        else:
            o, r, dn, dr = [], [], [], []
            for t in range(a.shape[0]):
                ob, rr, dd, _ = env.step(a[t])
                o.append(_np(ob["image"])[0].copy()); r.append(float(rr[0])); dn.append(bool(dd[0])); dr.append(int(ob["direction"][0]))
            o, r, dn, dr = np.stack(o), np.array(r), np.array(dn), np.array(dr, np.uint8)
        assert_same(tag + " done", dn.astype(np.uint8), d["done"][k, t0:ts + 1])
        assert_same(tag + " obs", o, d["obs"][k, t0:ts + 1])
        assert_same(tag + " dir", dr, d["dir"][k, t0:ts + 1])
        assert_same(tag + " reward bits", bits(r), bits(d["reward"][k, t0:ts + 1]))
        s = env.get_state()
        assert_same(tag + " snap_grid@%d" % ts, _np(s["grid"])[0], d["snap_grid"][k, si])
        assert_same(tag + " snap_agent@%d" % ts, _np(s["agent"])[0], d["snap_agent"][k, si])
        assert_same(tag + " snap_carrying@%d" % ts, _np(s["carrying"])[0], d["snap_carrying"][k, si])
        assert_same(tag + " snap_obst@%d" % ts, _np(s["obstacles"])[0, :nob], d["snap_obst"][k, si, :nob])
        assert_same(tag + " snap_target@%d" % ts, _np(s["target"])[0], d["snap_target"][k, si])
        t0 = ts + 1
    assert t0 == T
    env.check_errors()
    env.close()


@pytest.mark.parametrize("path", golden_files("philox_"), ids=os.path.basename)
def test_cuda_matches_reference_philox(path):
    d = load(path)
    d["path"] = path
    for k in range(d["actions"].shape[0]):
        _check_trace_gpu(d, k, "philox")


@pytest.mark.parametrize("path", golden_files("tape_"), ids=os.path.basename)
def test_cuda_matches_reference_tape(path):
    d = load(path)
    d["path"] = path
    for k in range(d["actions"].shape[0]):
        _check_trace_gpu(d, k, "tape")


@pytest.mark.parametrize("path", golden_files("scenes_"), ids=os.path.basename)
def test_cuda_matches_reference_scenes(path):
    """uploaded object soups (doors in all states, boxes, terminal goals, lava, carried objects,
    agents hugging every border), stepped without reset -- all envs of a file in ONE batch."""
    mgb = _mgb()
    d = load(path)
    n, T = d["actions"].shape
    env = mgb.make(d["env_id"], num_envs=n, autoreset=False)
    env.set_state(dict(grid=d["grid0"], aux=d["aux0"], agent=d["agent0"], carrying=d["carrying0"]))
    tag = os.path.basename(path)
    obs0 = env.reset(mask=np.zeros(n, np.uint8))          # observe without resetting anything
    assert_same(tag + " obs0", _np(obs0["image"]), d["obs0"])
    o, r, dn, dr = env.rollout(torch.as_tensor(d["actions"].T.copy()))
    assert_same(tag + " obs", _np(o).transpose(1, 0, 2, 3, 4), d["obs"])
    assert_same(tag + " dir", _np(dr).T, d["dir"])
    assert_same(tag + " done", _np(dn).T.astype(np.uint8), d["done"])
    assert_same(tag + " reward bits", bits(_np(r).T.copy()), bits(d["reward"]))
    s = env.get_state()
    assert_same(tag + " grid1", _np(s["grid"]), d["grid1"])
    assert_same(tag + " aux1", _np(s["aux"]), d["aux1"])
    assert_same(tag + " agent1", _np(s["agent"]), d["agent1"])
    assert_same(tag + " carrying1", _np(s["carrying"]), d["carrying1"])
    env.check_errors()


@pytest.mark.parametrize("env_id", HEADLINE + ["MiniGrid-DoorKey-5x5-v0", "MiniGrid-Dynamic-Obstacles-Random-6x6-v0",
                                               "MiniGrid-KeyCorridorS3R3-v0", "MiniGrid-Empty-Random-6x6-v0"])
def test_cuda_matches_oracle_batch(env_id):
    """thousands of envs (ragged tail group included), auto-reset on, TMA store path,
    against the oracle on the same seeded inputs; then the full state."""
    from oracle.oracle import OracleVec
    mgb = _mgb()
    cfg = _oracle_cfg(env_id)
    N, T, seed, base = 2048 + 32 * 3 + 7, 160, 99, 1000003
    rs = np.random.RandomState(5)
    actions = rs.randint(0, cfg["n_actions"], size=(T, N)).astype(np.uint8)
    env = mgb.make(env_id, num_envs=N, seed=seed, env_id_base=base)
    orc = OracleVec(cfg, N, seed=seed, env0=base)
    o0 = env.reset()
    ro0, rd0 = orc.reset()
    assert_same(env_id + " reset obs", _np(o0["image"]), ro0)
    assert_same(env_id + " reset dir", _np(o0["direction"]), rd0)
    half = T // 2
    got = [env.rollout(torch.as_tensor(actions[:half]))]
    # second half through T single-step launches
    obs_l, r_l, d_l, dir_l = [], [], [], []
    for t in range(half, T):
        ob, r, dn, _ = env.step(torch.as_tensor(actions[t]))
        obs_l.append(_np(ob["image"]).copy()); r_l.append(_np(r).copy()); d_l.append(_np(dn).copy()); dir_l.append(_np(ob["direction"]).copy())
    want = orc.rollout(actions, autoreset=True)
    o, r, dn, dr = got[0]
    o = np.concatenate([_np(o), np.stack(obs_l)]); r = np.concatenate([_np(r), np.stack(r_l)])
    dn = np.concatenate([_np(dn), np.stack(d_l)]); dr = np.concatenate([_np(dr), np.stack(dir_l)])
    assert_same(env_id + " done", dn.astype(np.uint8), want[2])
    assert_same(env_id + " obs", o, want[0])
    assert_same(env_id + " dir", dr, want[3])
    assert_same(env_id + " reward bits", bits(r), bits(want[1]))
    s, so = env.get_state(), orc.get_state()
    for key in ("grid", "aux", "agent", "carrying", "target"):
        assert_same(env_id + " state." + key, _np(s[key]), so[key])
    nob = cfg["n_obstacles"]
    assert_same(env_id + " state.obstacles", _np(s["obstacles"])[:, :nob], so["obstacles"][:, :nob])
    assert_same(env_id + " state.rng", _np(s["rng"]).view(np.uint32), so["rng"])
    env.check_errors()


def test_reward_formula_every_step_count():
    """_reward() = 1 - 0.9*(step_count/max_steps) for EVERY step count of every max_steps in the
    registry, against the values the reference itself computed (tests/golden/reward_table.npz)."""
    mgb = _mgb()
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "reward_table.npz"))
    for env_id in ("MiniGrid-Empty-8x8-v0", "MiniGrid-FourRooms-v0", "MiniGrid-DoorKey-16x16-v0", "MiniGrid-KeyCorridorS6R3-v0"):
        cfg = mgb.spec(env_id)["config"]
        ms, W, H = cfg["max_steps"], cfg["width"], cfg["height"]
        want = z["r_%d" % ms]                      # want[k] = reward when step_count == k after the step
        n = ms + 40                                # beyond max_steps too (stepping past done)
        grid = np.zeros((n, W, H, 3), np.uint8)
        grid[..., 0] = 1
        grid[:, 0, :, :] = grid[:, W - 1, :, :] = (2, 5, 0)
        grid[:, :, 0, :] = grid[:, :, H - 1, :] = (2, 5, 0)
        grid[:, 2, 1, :] = (8, 1, 0)
        aux = np.zeros((n, W, H), np.uint8)
        aux[:, 2, 1] = 1                           # Goal(toggletimes=0): overlap=True
        agent = np.zeros((n, 4), np.int32)
        agent[:, 0] = 1; agent[:, 1] = 1; agent[:, 2] = 0
        agent[:, 3] = np.arange(n)                 # step_count before the step
        env = mgb.make(env_id, num_envs=n, autoreset=False)
        env.set_state(dict(grid=grid, aux=aux, agent=agent))
        _, r, dn, _ = env.step(torch.full((n,), 2, dtype=torch.uint8))
        assert bool(dn.all())
        assert_same(env_id + " reward bits", bits(_np(r)), bits(want[1:n + 1]))
        env.check_errors()


def test_step_host_matches_step():
    mgb = _mgb()
    N = 32 * 1100 + 5                               # > 1024 groups: the chunked 3-stream pipeline is exercised
    a = np.random.RandomState(3).randint(0, 7, size=(6, N)).astype(np.uint8)
    e1 = mgb.make("MiniGrid-DoorKey-8x8-v0", num_envs=N, seed=11)
    e2 = mgb.make("MiniGrid-DoorKey-8x8-v0", num_envs=N, seed=11)
    e1.reset(); e2.reset()
    for t in range(a.shape[0]):
        o1, r1, d1, _ = e1.step(torch.as_tensor(a[t]))
        o2, r2, d2, _ = e2.step_host(torch.as_tensor(a[t]).pin_memory())
        assert_same("obs", _np(o1["image"]), o2["image"].numpy())
        assert_same("dir", _np(o1["direction"]), o2["direction"].numpy())
        assert_same("reward", bits(_np(r1)), bits(r2.numpy()))
        assert_same("done", _np(d1), d2.numpy())


def test_sharding_invariance():
    """(e) multi-GPU: env batches shard by global env id with no collective; two half-size handles
    with the right env_id_base must reproduce one full-size handle exactly."""
    mgb = _mgb()
    N, T = 4096, 64
    a = np.random.RandomState(8).randint(0, 3, size=(T, N)).astype(np.uint8)
    full = mgb.make("MiniGrid-Dynamic-Obstacles-16x16-v0", num_envs=N, seed=5)
    full.reset()
    fo, fr, fd, fdir = full.rollout(torch.as_tensor(a))
    for lo, hi in ((0, 1536), (1536, N)):
        part = mgb.make("MiniGrid-Dynamic-Obstacles-16x16-v0", num_envs=hi - lo, seed=5, env_id_base=lo)
        part.reset()
        po, pr, pd, pdir = part.rollout(torch.as_tensor(a[:, lo:hi].copy()))
        assert torch.equal(po, fo[:, lo:hi]) and torch.equal(pd, fd[:, lo:hi])
        assert torch.equal(pr.view(torch.int64), fr[:, lo:hi].contiguous().view(torch.int64))
        assert torch.equal(pdir, fdir[:, lo:hi])


def test_full_obs_and_invalid_action_flag():
    mgb = _mgb()
    env = mgb.make("MiniGrid-Empty-8x8-v0", num_envs=40)
    env.reset()
    f = _np(env.full_obs())
    s = env.get_state()
    g = _np(s["grid"]).copy()
    ag = _np(s["agent"])
    for i in range(40):
        g[i, ag[i, 0], ag[i, 1]] = (10, 0, ag[i, 2])      # wrappers.py:327-333
    assert_same("full_obs", f, g)
    env.step(torch.full((40,), 7, dtype=torch.uint8))      # reference: assert False, "unknown action"
    with pytest.raises(Exception, match="unknown action"):
        env.check_errors()
    env.check_errors()                                       # flags are cleared once reported


def test_seed_reproducible_and_reseed():
    mgb = _mgb()
    a = torch.as_tensor(np.random.RandomState(1).randint(0, 7, size=(50, 300)).astype(np.uint8))
    outs = []
    for _ in range(2):
        env = mgb.make("MiniGrid-FourRooms-v0", num_envs=300, seed=42)
        env.reset()
        outs.append(env.rollout(a)[0])
    assert torch.equal(outs[0], outs[1])
    env.seed(42)
    env.reset()
    assert torch.equal(env.rollout(a)[0], outs[0])
    env.seed(43)
    env.reset()
    assert not torch.equal(env.rollout(a)[0], outs[0])


def test_full_size_invariants():
    """BASELINE-size batch (2^20 envs): properties that need no oracle (run_tests.py:48-62 invariants):
    agent stays in bounds, the cell under the agent is empty-or-carried, rewards in range, the image
    only contains valid codes, and two launches with different step partitioning agree bit for bit."""
    mgb = _mgb()
    N, T = 1 << 20, 8
    g = torch.Generator(device="cuda").manual_seed(1234)
    a = torch.randint(0, 7, (T, N), dtype=torch.uint8, device="cuda", generator=g)
    env = mgb.make("MiniGrid-Empty-8x8-v0", num_envs=N, seed=0)
    env.reset()
    o, r, d, dr = env.rollout(a)
    env2 = mgb.make("MiniGrid-Empty-8x8-v0", num_envs=N, seed=0)
    env2.reset()
    o2 = torch.stack([env2.step(a[t])[0]["image"].clone() for t in range(T)])
    assert torch.equal(o, o2)
    assert int(o[..., 0].max()) <= 10 and int(o[..., 1].max()) <= 6 and int(o[..., 2].max()) <= 2
    assert bool((o[:, :, 3, 6, 0] == 1).all())               # nothing carried in Empty: agent cell shows 'empty'
    assert float(r.min()) >= 0.0 and float(r.max()) <= 1.0
    ag = env.get_state(("agent",))["agent"]
    assert bool(((ag[:, 0] >= 1) & (ag[:, 0] <= 6) & (ag[:, 1] >= 1) & (ag[:, 1] <= 6)).all())
    assert bool((ag[:, 3] == T).all())
    env.check_errors()


@pytest.mark.parametrize("path", golden_files("dynobs_boxed_"), ids=os.path.basename)
def test_cuda_matches_reference_boxed_obstacles(path):
    """Directed Dynamic-Obstacles cases: walled-in ball (101 tries = 202 draws per step, stays), one free neighbour,
    agent next to a ball -- exercises the speculative tries, the one-try-at-a-time continuation past the prefetched
    draw window and the on-demand Philox blocks; the draw counter is compared after every step."""
    mgb = _mgb()
    d = load(path)
    n, T = d["actions"].shape
    for k in range(n):
        env = mgb.make(d["env_id"], num_envs=1, seed=int(d["seed"][k]), env_id_base=int(d["env_index"][k]), autoreset=False)
        env.reset()
        env.set_state(dict(grid=d["grid0"][k:k + 1], aux=d["aux0"][k:k + 1], agent=d["agent0"][k:k + 1], carrying=d["carrying0"][k:k + 1],
                           obstacles=d["obstacles0"][k:k + 1], rng=np.array([[1, d["ndraws0"][k]]], np.int32)))
        tag = "%s[%d]" % (os.path.basename(path), k)
        for t in range(T):
            obs, r, dn, _ = env.step(torch.as_tensor(d["actions"][k, t:t + 1]))
            s = env.get_state(("rng", "obstacles"))
            assert_same("%s@%d obs" % (tag, t), _np(obs["image"])[0], d["obs"][k, t])
            assert int(obs["direction"][0]) == int(d["dir"][k, t]) and int(dn[0]) == int(d["done"][k, t])
            assert bits(_np(r))[0] == bits(d["reward"][k, t:t + 1])[0]
            assert int(_np(s["rng"])[0, 1]) == int(d["ndraws"][k, t]), "%s@%d draws" % (tag, t)
            assert_same("%s@%d obstacles" % (tag, t), _np(s["obstacles"])[0], d["obstacles"][k, t])
        assert_same(tag + " grid1", _np(env.get_state(("grid",))["grid"])[0], d["grid1"][k])
        env.check_errors()


@pytest.mark.parametrize("n", [1, 31, 32, 33, 65])
def test_small_and_ragged_batches(n):
    """edge sizes: a single env, one lane short of / exactly / one over a 32-env group; rollout == oracle, and the
    unaligned / ragged store path (no bulk copy) is the one exercised.  An empty batch is refused with an error."""
    from oracle.oracle import OracleVec
    mgb = _mgb()
    env_id = "MiniGrid-Dynamic-Obstacles-8x8-v0"
    cfg = _oracle_cfg(env_id)
    T = 40
    actions = np.random.RandomState(n).randint(0, cfg["n_actions"], size=(T, n)).astype(np.uint8)
    env = mgb.make(env_id, num_envs=n, seed=3, env_id_base=77)
    orc = OracleVec(cfg, n, seed=3, env0=77, threads=1)
    o0, d0 = orc.reset()
    g0 = env.reset()
    assert_same("obs0", _np(g0["image"]), o0)
    o, r, dn, dr = env.rollout(torch.as_tensor(actions))
    oo, orr, odn, odr = orc.rollout(actions)
    assert_same("obs", _np(o), oo)
    assert_same("done", _np(dn).astype(np.uint8), odn)
    assert_same("dir", _np(dr), odr)
    assert_same("reward bits", bits(_np(r)), bits(orr))
    env.check_errors()
    with pytest.raises(Exception):
        mgb.make(env_id, num_envs=0)


@pytest.mark.parametrize("env_id", ["MiniGrid-Empty-8x8-v0", "MiniGrid-Dynamic-Obstacles-16x16-v0", "MiniGrid-DoorKey-16x16-v0"])
def test_random_policy_rollout_matches_oracle(env_id):
    """mgb_rollout_random: the uniform random policy drawn on the device (no action input) -- actions and everything
    downstream equal the oracle's; at 2^20 envs the action histogram is uniform to 0.5 %."""
    from oracle.oracle import OracleVec
    mgb = _mgb()
    cfg = _oracle_cfg(env_id)
    N, T = 2048 + 17, 48
    env = mgb.make(env_id, num_envs=N, seed=21, env_id_base=123456789012)
    orc = OracleVec(cfg, N, seed=21, env0=123456789012)
    env.reset(); orc.reset()
    o, r, dn, dr, a = env.rollout_random(T)
    oo, orr, odn, odr, oa = orc.rollout_random(T)
    assert_same("actions", _np(a), oa)
    assert_same("obs", _np(o), oo)
    assert_same("done", _np(dn).astype(np.uint8), odn)
    assert_same("dir", _np(dr), odr)
    assert_same("reward bits", bits(_np(r)), bits(orr))
    env.check_errors()
    big = mgb.make(env_id, num_envs=1 << 20, seed=1)
    big.reset()
    _, _, _, _, a = big.rollout_random(8, want_obs=False)
    h = torch.bincount(a.flatten().long(), minlength=cfg["n_actions"]).double()
    assert h.numel() == cfg["n_actions"] and float((h / h.sum() - 1.0 / cfg["n_actions"]).abs().max()) < 0.005 / cfg["n_actions"] * cfg["n_actions"]


@pytest.mark.parametrize("env_id,misalign", [("MiniGrid-DoorKey-16x16-v0", 0), ("MiniGrid-FourRooms-v0", 0),
                                             ("MiniGrid-KeyCorridorS6R3-v0", 1), ("MiniGrid-Empty-8x8-v0", 0)])
def test_aligned_long_rollout_matches_oracle(env_id, misalign):
    """N a multiple of 32 and T > 64: the occluded kernels fetch 32 steps of actions at a time with 16-byte row loads
    (chunks 32 + 32 + 6 here), or byte-wise when the action array does not start on a 16-byte boundary (misalign=1);
    the state block comes and goes with bulk copies.  Everything against the oracle, then the full state."""
    from oracle.oracle import OracleVec
    mgb = _mgb()
    cfg = _oracle_cfg(env_id)
    N, T, seed, base = 1024, 70, 11, 4242
    actions = np.random.RandomState(17).randint(0, cfg["n_actions"], size=(T, N)).astype(np.uint8)
    env = mgb.make(env_id, num_envs=N, seed=seed, env_id_base=base)
    orc = OracleVec(cfg, N, seed=seed, env0=base)
    env.reset(); orc.reset()
    buf = torch.zeros(T * N + 16, dtype=torch.uint8, device="cuda")
    a_dev = buf[misalign:misalign + T * N].view(T, N)
    a_dev.copy_(torch.as_tensor(actions))
    assert (a_dev.data_ptr() % 16 == 0) == (misalign == 0)
    o, r, dn, dr = env.rollout(a_dev)
    wo, wr, wdn, wdr = orc.rollout(actions, autoreset=True)
    assert_same(env_id + " obs", _np(o), wo)
    assert_same(env_id + " done", _np(dn).astype(np.uint8), wdn)
    assert_same(env_id + " dir", _np(dr), wdr)
    assert_same(env_id + " reward bits", bits(_np(r)), bits(wr))
    s, so = env.get_state(), orc.get_state()
    for key in ("grid", "agent", "carrying", "target"):
        assert_same(env_id + " state." + key, _np(s[key]), so[key])
    assert_same(env_id + " state.rng", _np(s["rng"]).view(np.uint32), so["rng"])
    env.check_errors()


def test_invalid_action_inside_a_rollout_is_flagged():
    """an out-of-range action deep inside a rollout (step 40 of 48, second 32-step chunk of the packed action path)
    raises the same error flag as in a single step; 15 and 200 are both invalid."""
    mgb = _mgb()
    for bad in (7, 15, 200):
        env = mgb.make("MiniGrid-DoorKey-16x16-v0", num_envs=64)
        env.reset()
        a = torch.zeros((48, 64), dtype=torch.uint8, device="cuda")
        env.rollout(a)
        env.check_errors()
        a[40, 33] = bad
        env.rollout(a)
        with pytest.raises(Exception, match="unknown action"):
            env.check_errors()


# ---------------------------------------------------------------------------------------------------------------
# Batched replays of the golden traces: every trace of a fixture file in ONE handle of full 32-env groups, the whole
# trace in ONE persistent launch -- the bulk-copy (TMA) observation path, the packed action path and the
# warp-cooperative Dynamic-Obstacles reset are compared with the reference's bytes directly.
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", golden_files("tape_"), ids=os.path.basename)
def test_batched_tape_replay_matches_reference(path):
    """RNG-tape mode: env j of a 64-env handle replays trace j mod n (its own copy of the reference's MT19937 draws)."""
    mgb = _mgb()
    d = load(path)
    n, T = d["actions"].shape[:2]
    N = 64
    which = np.arange(N) % n
    tapes = [d["tape"][int(d["tape_offsets"][k]):int(d["tape_offsets"][k + 1])] for k in which]
    offs = np.concatenate([[0], np.cumsum([len(t) for t in tapes])]).astype(np.int64)
    env = mgb.make(d["env_id"], num_envs=N, seed=int(d["seed"]))
    env.set_rng_tape(np.concatenate(tapes) if offs[-1] else np.zeros(0, np.int32), offs)
    obs = env.reset()
    tag = os.path.basename(path)
    assert_same(tag + " obs0", _np(obs["image"]), d["obs0"][which])
    assert_same(tag + " dir0", _np(obs["direction"]), d["dir0"][which].astype(np.uint8))
    o, r, dn, dr = env.rollout(torch.as_tensor(d["actions"][which].T.copy()))
    assert_same(tag + " done", _np(dn).T.astype(np.uint8), d["done"][which])
    assert_same(tag + " obs", _np(o).transpose(1, 0, 2, 3, 4), d["obs"][which])
    assert_same(tag + " dir", _np(dr).T, d["dir"][which])
    assert_same(tag + " reward bits", bits(_np(r).T.copy()), bits(d["reward"][which]))
    env.check_errors()
    env.close()


@pytest.mark.parametrize("path", golden_files("philox_"), ids=os.path.basename)
def test_batched_philox_replay_matches_reference(path):
    """Philox mode: one handle covers the global env ids of all traces of the file (the others idle on action 'done'),
    so that the traces run on the device generators, the auto-reset and the bulk-copy path of full groups."""
    mgb = _mgb()
    d = load(path)
    ids = [int(i) for i in d["env_indices"]]
    keep = [k for k, i in enumerate(ids) if i < (1 << 20)]                     # the id beyond 2^32 has its own single-env test
    base = min(ids[k] for k in keep)
    N = (max(ids[k] for k in keep) - base + 1 + 31) // 32 * 32
    T = d["actions"].shape[1]
    n_act = int(d["cfg"]["n_actions"])
    acts = np.full((T, N), min(6, n_act - 1), np.uint8)                        # idle envs: 'done' (or the last action that exists)
    for k in keep:
        acts[:, ids[k] - base] = d["actions"][k]
    env = mgb.make(d["env_id"], num_envs=N, seed=int(d["seed"]), env_id_base=base)
    obs = env.reset()
    tag = os.path.basename(path)
    cols = [ids[k] - base for k in keep]
    assert_same(tag + " obs0", _np(obs["image"])[cols], d["obs0"][keep])
    o, r, dn, dr = env.rollout(torch.as_tensor(acts))
    assert_same(tag + " done", _np(dn)[:, cols].T.astype(np.uint8), d["done"][keep])
    assert_same(tag + " obs", _np(o)[:, cols].transpose(1, 0, 2, 3, 4), d["obs"][keep])
    assert_same(tag + " dir", _np(dr)[:, cols].T, d["dir"][keep])
    assert_same(tag + " reward bits", bits(_np(r)[:, cols].T.copy()), bits(d["reward"][keep]))
    env.check_errors()
    env.close()


def test_step_and_reset_return_fresh_tensors():
    """like the reference (fresh numpy arrays per call), obs from reset()/step() must survive later steps"""
    mgb = _mgb()
    env = mgb.make("MiniGrid-DoorKey-8x8-v0", num_envs=96, seed=3)
    o0 = env.reset()
    keep0 = o0["image"].clone()
    a = torch.randint(0, 7, (96,), dtype=torch.uint8)
    o1, r1, d1, _ = env.step(a)
    keep1, keepr = o1["image"].clone(), r1.clone()
    o2, r2, d2, _ = env.step(torch.full((96,), 2, dtype=torch.uint8))
    env.gen_obs(); env.agent_sees(1, 1)
    assert torch.equal(o0["image"], keep0) and torch.equal(o1["image"], keep1) and torch.equal(r1, keepr)
    assert o1["image"].data_ptr() != o2["image"].data_ptr() != o0["image"].data_ptr()
    # out= is the zero-allocation path: the caller's buffers are written
    out = env._new_out()
    o3, r3, d3, _ = env.step(a, out=out)
    assert o3["image"].data_ptr() == out[0].data_ptr() and r3.data_ptr() == out[1].data_ptr()


def test_env_grid_and_metadata():
    """env.grid (SURVEY 8b attribute list): lazy batched view with encode()/get(); render.modes lists rgb_array"""
    mgb = _mgb()
    env = mgb.make("MiniGrid-Empty-8x8-v0", num_envs=40)
    env.reset()
    g = env.grid
    assert (g.width, g.height) == (8, 8)
    enc = _np(g.encode())
    assert enc.shape == (40, 8, 8, 3)
    assert_same("grid == get_state", enc, _np(env.get_state(("grid",))["grid"]))
    assert_same("goal", _np(g.get(6, 6)), np.tile(np.array([8, 1, 0], np.uint8), (40, 1)))
    assert_same("wall", _np(g.get(0, 3)), np.tile(np.array([2, 5, 0], np.uint8), (40, 1)))
    assert ('green', 'goal') in g and ('red', 'key') not in g
    m = np.zeros((8, 8), bool); m[6, 6] = True
    assert int(_np(g.encode(vis_mask=m))[:, :, :, 0].sum()) == 40 * 8
    assert 'rgb_array' in env.metadata['render.modes']
    assert env.render('rgb_array').shape == (40, 64, 64, 3)


def test_handle_calls_leave_the_current_device_alone():
    """entry points run on the handle's device and restore the caller's current device"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mgb = _mgb()
    torch.cuda.set_device(0)
    env = mgb.make("MiniGrid-Empty-8x8-v0", num_envs=64, device="cuda:1")
    env.reset()
    env.step(torch.zeros(64, dtype=torch.uint8))
    assert torch.cuda.current_device() == 0
    x = torch.ones(4, device="cuda")
    assert x.device.index == 0


def _scatter_step_counts(env, orc, rs, lo_left, hi_left, max_steps):
    """give every env its own remaining time: step_count = max_steps - left, left drawn from [lo_left, hi_left)"""
    so = orc.get_state()
    left = rs.randint(lo_left, hi_left, size=so["agent"].shape[0])
    so["agent"][:, 3] = max_steps - left
    orc.set_state(so)
    env.set_state({"agent": torch.as_tensor(so["agent"])})


@pytest.mark.parametrize("env_id", ["MiniGrid-DoorKey-5x5-v0", "MiniGrid-DoorKey-16x16-v0", "MiniGrid-FourRooms-v0",
                                    "MiniGrid-KeyCorridorS6R3-v0", "MiniGrid-KeyCorridorS3R1-v0", "MiniGrid-LavaCrossingS9N2-v0",
                                    "MiniGrid-SimpleCrossingS11N5-v0", "MiniGrid-LavaGapS7-v0", "MiniGrid-MultiRoom-N4-S5-v0"])
def test_spread_out_episode_ends_match_oracle(env_id):
    """Episode ends that do NOT come in lock-step: each env reaches its time limit at its own step, so resets hit single
    lanes of a warp.  The kernels then generate next-episode layouts ahead of time for the other lanes of the warp (spare
    layouts) and later resets copy them; outputs and state must stay bit-identical to the oracle, which generates at the
    reset.  Three rounds so that spares are made, consumed, and made again; rollouts and single-step launches mixed; a
    grid upload in between must leave the spares usable, an rng upload must drop them."""
    from oracle.oracle import OracleVec
    mgb = _mgb()
    cfg = _oracle_cfg(env_id)
    N, T, seed, base = 1024 + 32 + 5, 48, 7, 5000
    rs = np.random.RandomState(11)
    env = mgb.make(env_id, num_envs=N, seed=seed, env_id_base=base)
    orc = OracleVec(cfg, N, seed=seed, env0=base)
    o0 = env.reset()
    ro0, _ = orc.reset()
    assert_same(env_id + " reset obs", _np(o0["image"]), ro0)
    for rnd in range(4):
        _scatter_step_counts(env, orc, rs, 1, 40, cfg["max_steps"])
        if rnd == 2:                                   # same state, uploaded in full (rng included): pre-generated layouts are dropped
            so = orc.get_state()
            env.set_state({k: torch.as_tensor(so[k].astype(np.int64) if k == "rng" else so[k]) for k in so})
        actions = rs.randint(0, cfg["n_actions"], size=(T, N)).astype(np.uint8)
        want = orc.rollout(actions, autoreset=True)
        if rnd % 2 == 0:
            o, r, dn, dr = [_np(x) for x in env.rollout(torch.as_tensor(actions))]
        else:
            res = []
            for t in range(T):
                ob, r, dn, _ = env.step(torch.as_tensor(actions[t]))
                res.append((_np(ob["image"]).copy(), _np(r).copy(), _np(dn).copy(), _np(ob["direction"]).copy()))
            o, r, dn, dr = [np.stack(x) for x in zip(*res)]
        tag = "%s round %d" % (env_id, rnd)
        assert want[2].sum() >= N, tag                 # every env ended an episode in this round
        assert_same(tag + " done", dn.astype(np.uint8), want[2])
        assert_same(tag + " obs", o, want[0])
        assert_same(tag + " dir", dr, want[3])
        assert_same(tag + " reward bits", bits(r), bits(want[1]))
        s, so = env.get_state(), orc.get_state()
        for key in ("grid", "aux", "agent", "carrying", "target"):
            assert_same(tag + " state." + key, _np(s[key]), so[key])
        assert_same(tag + " state.rng", _np(s["rng"]).view(np.uint32), so["rng"])
    env.check_errors()


def test_reseed_drops_pre_generated_layouts():
    """a layout generated ahead of time belongs to (seed, env, episode): after env.seed() the same episode numbers come round
    again under another seed and must not pick up the old layouts"""
    from oracle.oracle import OracleVec
    mgb = _mgb()
    env_id = "MiniGrid-DoorKey-6x6-v0"
    cfg = _oracle_cfg(env_id)
    N, T = 96, 40
    rs = np.random.RandomState(3)
    env = mgb.make(env_id, num_envs=N, seed=1)
    for seed in (1, 2, 1):
        env.seed(seed)
        orc = OracleVec(cfg, N, seed=seed)
        env.reset(); orc.reset()
        _scatter_step_counts(env, orc, rs, 1, 30, cfg["max_steps"])
        actions = rs.randint(0, cfg["n_actions"], size=(T, N)).astype(np.uint8)
        want = orc.rollout(actions, autoreset=True)
        got = [_np(x) for x in env.rollout(torch.as_tensor(actions))]
        assert_same("seed %d obs" % seed, got[0], want[0])
        assert_same("seed %d done" % seed, got[2].astype(np.uint8), want[2])
        assert_same("seed %d state.grid" % seed, _np(env.get_state()["grid"]), orc.get_state()["grid"])


def test_group_tickets_survive_odd_launch_shapes():
    """Groups beyond a warp's first are handed out by a device-side ticket counter that the kernel itself zeroes when its
    last warp leaves.  Mix launch kinds and sizes on one handle -- fewer groups than resident warps, exactly one group, many
    more groups than warps; resets with a mask, single steps, rollouts, the host pipeline -- and require the outputs to stay
    equal to the oracle's: a counter left non-zero by any of them would make the next launch skip or repeat groups."""
    from oracle.oracle import OracleVec
    mgb = _mgb()
    env_id = "MiniGrid-DoorKey-8x8-v0"
    cfg = _oracle_cfg(env_id)
    rs = np.random.RandomState(2)
    for N in (1, 33, 32 * 700 + 5, 32 * 5000):
        env = mgb.make(env_id, num_envs=N, seed=11)
        orc = OracleVec(cfg, N, seed=11)
        env.reset(); orc.reset()
        for rnd in range(3):
            a = rs.randint(0, cfg["n_actions"], size=(5, N)).astype(np.uint8)
            want = orc.rollout(a)
            got = [_np(x) for x in env.rollout(torch.as_tensor(a))]
            assert_same("N=%d rollout %d obs" % (N, rnd), got[0], want[0])
            a1 = rs.randint(0, cfg["n_actions"], size=N).astype(np.uint8)
            w1 = orc.step(a1)
            ob, r, dn, _ = env.step(torch.as_tensor(a1))
            assert_same("N=%d step %d obs" % (N, rnd), _np(ob["image"]), w1[0])
            a2 = rs.randint(0, cfg["n_actions"], size=N).astype(np.uint8)
            w2 = orc.step(a2)
            h = env.step_host(a2)
            assert_same("N=%d step_host %d obs" % (N, rnd), _np(h[0]["image"]), w2[0])
            m = rs.randint(0, 2, size=N).astype(np.uint8)
            wo, _ = orc.reset(m)
            go = env.reset(mask=torch.as_tensor(m))
            assert_same("N=%d masked reset %d obs" % (N, rnd), _np(go["image"])[m == 1], wo[m == 1])    # the oracle fills only the rows it reset
        env.check_errors()


def test_rollout_replayed_from_a_cuda_graph():
    """mgb_rollout keeps nothing per launch on the host (the group ticket counter is zeroed by the kernel itself), so a
    launch captured in a CUDA graph can be replayed: three replays must give what three launches give."""
    from oracle.oracle import OracleVec
    mgb = _mgb()
    env_id = "MiniGrid-FourRooms-v0"
    cfg = _oracle_cfg(env_id)
    N, T = 32 * 300 + 3, 8
    rs = np.random.RandomState(4)
    env = mgb.make(env_id, num_envs=N, seed=5)
    orc = OracleVec(cfg, N, seed=5)
    env.reset(); orc.reset()
    acts = torch.zeros((T, N), dtype=torch.uint8, device=env.device)
    out = env.rollout(acts)                                   # buffers to reuse; also warms up
    orc.rollout(np.zeros((T, N), np.uint8))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    a0 = rs.randint(0, cfg["n_actions"], size=(T, N)).astype(np.uint8)
    acts.copy_(torch.as_tensor(a0))
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        try:
            with torch.cuda.graph(g, stream=s):
                env.rollout(acts, out=out)
        except Exception as e:                                # pragma: no cover
            pytest.skip("capture not possible here: %r" % (e,))
    torch.cuda.synchronize()
    # the capture itself does not run the launch: replay it three times with fresh actions
    for rep in range(3):
        a = a0 if rep == 0 else rs.randint(0, cfg["n_actions"], size=(T, N)).astype(np.uint8)
        acts.copy_(torch.as_tensor(a))
        g.replay()
        torch.cuda.synchronize()
        want = orc.rollout(a)
        assert_same("replay %d obs" % rep, _np(out[0]), want[0])
        assert_same("replay %d done" % rep, _np(out[2]).astype(np.uint8), want[2])
    assert_same("state after replays", _np(env.get_state()["grid"]), orc.get_state()["grid"])
