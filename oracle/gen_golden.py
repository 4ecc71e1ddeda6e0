#!/usr/bin/env python
"""TEST INFRASTRUCTURE: generate tests/golden/*.npz from the LIVE, UNMODIFIED reference.

Run in the build container (needs /root/reference; uses oracle/ref_shim.py's gym stub):

    python oracle/gen_golden.py            # writes tests/golden/

The reference has no golden vectors of its own, so these files *are*
the pin for the C oracle and, through it, for the CUDA path.  Three families:

  philox_<id>.npz  RNG injection: env.np_random = PhiloxShim(seed, env_id, episode);
                   the reference's own _gen_grid / obstacle code consumes the shared
                   counter-based stream, so layouts + trajectories must match draw for draw.
  tape_<id>.npz    RNG tape: the reference runs with its own MT19937 RandomState, every
                   randint result is recorded and replayed by oracle/kernel.
  scenes_<name>.npz  state upload: random object soups (doors in all states, keys, balls,
                   boxes, default + terminal goals, lava, floor, carried objects, agents at
                   the borders) built with the reference's own classes, stepped with random
                   actions WITHOUT reset; covers the directed cases of SURVEY §8c.

Every file stores the config read off the reference env itself (width, height,
max_steps, see_through_walls, action_space.n, ...), per-step obs / direction /
reward (float64) / done, and periodic full-grid snapshots (Grid.encode layout).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim as R  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
MAX_OBST = 8
SNAP_EVERY = 50

HOOK_OF_CLASS = {"UnlockPickup": 1, "BlockedUnlockPickup": 1, "ObstructedMazeEnv": 1, "Unlock": 2, "FetchEnv": 3, "GoToDoorEnv": 4, "GoToObjectEnv": 5,
                 "PutNearEnv": 6, "RedBlueDoorEnv": 7, "MemoryEnv": 8}
GEN_OF_CLASS = {"EmptyEnv": 0, "DoorKeyEnv": 1, "FourRoomsEnv": 2, "DynamicObstaclesEnv": 3, "KeyCorridor": 4,
                "CrossingEnv": 6, "LavaGapEnv": 7, "MultiRoomEnv": 8, "DistShiftEnv": 9}


def gen_params_of(env, gen):
    """generator parameters (include/mgb200.h gen_param0/1) read off the live reference env"""
    mg = sys.modules["gym_minigrid.minigrid"]
    if gen == 6:      # crossing.py:11-22
        return int(env.num_crossings), int(env.ori) | (4 if env.obstacle_type is mg.Wall else 0)
    if gen == 7:      # lavagap.py:10-19
        return int(bool(env.const)), int(env.obstacle_type is mg.Wall)
    if gen == 8:      # multiroom.py:21-39
        assert env.minNumRooms == env.maxNumRooms
        return int(env.minNumRooms), int(env.maxRoomSize)
    if gen == 9:      # distshift.py:9-21
        assert tuple(env.agent_start_pos) == (1, 1) and env.agent_start_dir == 0
        return int(env.strip2_row), 0
    return 0, 0


def config_of(env):
    """Read the static config off a live reference env."""
    env = env.unwrapped
    gen = None
    # Empty variants whose generator uses the global np.random / sizetop (empty.py:34-46): level pool, not GEN_EMPTY
    if type(env).__name__ in ("EmptyEnv6x6Extra", "EmptyEnv6x6ExtraLava", "EmptyRandomEnv10x10"):
        gen = 5
    for klass in type(env).__mro__:
        if gen is None and klass.__name__ in GEN_OF_CLASS:
            gen = GEN_OF_CLASS[klass.__name__]
            break
    if gen is None:
        # no on-device generator for this class: level-pool mode, only legal when step() is the base step
        mg = sys.modules["gym_minigrid.minigrid"]
        names = [k.__name__ for k in type(env).__mro__]
        assert type(env).step is mg.MiniGridEnv.step or any(n in HOOK_OF_CLASS for n in names) or \
            any(n in ("LockedRoom", "PlaygroundV0") for n in names), \
            "%s overrides step() with an unknown hook: not a level-pool env" % type(env).__name__
        gen = 5
    gp0, gp1 = gen_params_of(env, gen)
    return dict(hook=HOOK_OF_CLASS.get(next((k.__name__ for k in type(env).__mro__ if k.__name__ in HOOK_OF_CLASS), ""), 0),
        gen_param0=gp0, gen_param1=gp1,
        **dict(
        gen=gen, width=env.width, height=env.height, max_steps=env.max_steps,
        see_through=int(bool(env.see_through_walls)), n_actions=env.action_space.n,
        n_obstacles=getattr(env, "n_obstacles", 0) if gen == 3 else 0,
        room_size=getattr(env, "room_size", 0), num_rows=getattr(env, "num_rows", 0),
        random_start=int(gen in (0, 3) and getattr(env, "agent_start_pos", 1) is None),
        lava_v1=int("v1" in type(env).__name__),     # minigrid.py:1263 -- e.g. "DoorKeyEnv16x16" contains "v1"
    ))


def pad_obst(o):
    out = np.zeros((MAX_OBST, 2), np.int16)
    out[:len(o)] = o
    return out


def run_trace(env_id, seed, env_index, T, act_seed, mode):
    """One env, T steps, auto-reset by the caller exactly like run_tests.py:37-46."""
    env = R.make(env_id)
    cfg = config_of(env)
    if mode == "philox":
        shim = R.PhiloxShim(seed, env_index, 0)
    else:
        shim = R.TapeRecorder(np.random.RandomState((seed * 7919 + env_index) % (2 ** 32)))
    env.np_random = shim
    episode = 0
    obs = env.reset()
    episode += 1
    rs = np.random.RandomState(act_seed)
    actions = rs.randint(0, cfg["n_actions"], size=T).astype(np.uint8)
    out = dict(obs=np.zeros((T, 7, 7, 3), np.uint8), dir=np.zeros(T, np.uint8),
               reward=np.zeros(T, np.float64), done=np.zeros(T, np.uint8))
    snaps, snap_t, missions = [], [], [obs["mission"]]
    first = dict(obs0=obs["image"].copy(), dir0=np.uint8(obs["direction"]))
    snap0 = R.snapshot(env)
    for t in range(T):
        obs, r, d, _ = env.step(int(actions[t]))
        if d:
            if mode == "philox":
                shim.new_episode(episode)
            obs = env.reset()
            episode += 1
            missions.append(obs["mission"])
        out["obs"][t] = obs["image"]
        out["dir"][t] = obs["direction"]
        out["reward"][t] = float(r)
        out["done"][t] = d
        if (t + 1) % SNAP_EVERY == 0 or t == T - 1:
            s = R.snapshot(env)
            snaps.append(s)
            snap_t.append(t)
    res = dict(cfg=cfg, actions=actions, **out, **first,
               snap_t=np.array(snap_t, np.int32),
               snap_grid=np.stack([s["grid"] for s in snaps]),
               snap_agent=np.stack([s["agent"] for s in snaps]),
               snap_carrying=np.stack([s["carrying"] for s in snaps]),
               snap_obst=np.stack([pad_obst(s["obstacles"]) for s in snaps]),
               snap_target=np.stack([s["target"] for s in snaps]),
               grid0=snap0["grid"], agent0=snap0["agent"], obst0=pad_obst(snap0["obstacles"]),
               target0=snap0["target"], n_episodes=episode, mission0=missions[0])
    if mode == "tape":
        res["tape"] = np.array(shim.tape, np.int32)
    return res


def save_traces(name, env_id, traces, seed, env_indices, act_seeds):
    cfg = traces[0]["cfg"]
    d = dict(env_id=env_id, seed=np.uint64(seed), env_indices=np.array(env_indices, np.int64),
             act_seeds=np.array(act_seeds, np.int64),
             cfg_keys=np.array(list(cfg.keys())), cfg_vals=np.array(list(cfg.values()), np.int32),
             missions0=np.array([t["mission0"] for t in traces]),
             n_episodes=np.array([t["n_episodes"] for t in traces], np.int32))
    for k in ("actions", "obs", "dir", "reward", "done", "obs0", "dir0", "snap_t", "snap_grid", "snap_agent",
              "snap_carrying", "snap_obst", "snap_target", "grid0", "agent0", "obst0", "target0"):
        d[k] = np.stack([t[k] for t in traces])
    if "tape" in traces[0]:
        d["tape"] = np.concatenate([t["tape"] for t in traces])
        d["tape_offsets"] = np.concatenate([[0], np.cumsum([len(t["tape"]) for t in traces])]).astype(np.int64)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    return path


# ---------------------------------------------------------------------------
# scene fuzz: object soups built from the reference's own classes
# ---------------------------------------------------------------------------
def build_scene(env, rs, density):
    mg = sys.modules["gym_minigrid.minigrid"]
    env = env.unwrapped
    W, H = env.width, env.height
    env.grid = mg.Grid(W, H)
    env.grid.wall_rect(0, 0, W, H)
    colors = list(mg.COLOR_TO_IDX.keys())

    def rand_obj():
        k = rs.randint(0, 12)
        c = colors[rs.randint(0, len(colors))]
        if k == 0:
            return mg.Wall(c)
        if k == 1:
            return mg.Wall()
        if k == 2:
            return mg.Floor(c)
        if k in (3, 4, 5):
            st = rs.randint(0, 3)
            return mg.Door(c, is_open=(st == 0), is_locked=(st == 2))
        if k == 6:
            return mg.Key(c)
        if k == 7:
            return mg.Ball(c)
        if k == 8:
            return mg.Box(c)
        if k == 9:
            return mg.Goal()
        if k == 10:
            return mg.Goal(toggletimes=0)       # the only kind of goal that terminates (minigrid.py:157-160,1259)
        return mg.Lava()

    for x in range(1, W - 1):
        for y in range(1, H - 1):
            if rs.rand() < density:
                env.grid.set(x, y, rand_obj())
    # agent on an overlappable cell, biased towards the borders so every view clip is hit
    while True:
        if rs.rand() < 0.5:
            ax = [1, W - 2][rs.randint(0, 2)] if rs.rand() < 0.5 else rs.randint(1, W - 1)
            ay = [1, H - 2][rs.randint(0, 2)] if rs.rand() < 0.5 else rs.randint(1, H - 1)
        else:
            ax, ay = rs.randint(1, W - 1), rs.randint(1, H - 1)
        c = env.grid.get(ax, ay)
        if c is None or c.can_overlap():
            break
        env.grid.set(ax, ay, None)
        break
    env.agent_pos = np.array([ax, ay])
    env.agent_dir = int(rs.randint(0, 4))
    env.carrying = None
    k = rs.randint(0, 5)
    if k == 1:
        env.carrying = mg.Key(colors[rs.randint(0, len(colors))])
    elif k == 2:
        env.carrying = mg.Ball(colors[rs.randint(0, len(colors))])
    elif k == 3:
        env.carrying = mg.Box(colors[rs.randint(0, len(colors))])
    env.step_count = int(rs.randint(0, env.max_steps))


def run_scene(env_id, scene_seed, T, density):
    env = R.make(env_id)
    rs = np.random.RandomState(scene_seed)
    build_scene(env, rs, density)
    cfg = config_of(env)
    s0 = R.snapshot(env)
    actions = rs.randint(0, 7, size=T).astype(np.uint8)
    obs0 = env.gen_obs()
    out = dict(obs=np.zeros((T, 7, 7, 3), np.uint8), dir=np.zeros(T, np.uint8),
               reward=np.zeros(T, np.float64), done=np.zeros(T, np.uint8))
    for t in range(T):
        obs, r, d, _ = env.step(int(actions[t]))       # never reset: stepping past done is legal in the reference
        out["obs"][t] = obs["image"]
        out["dir"][t] = obs["direction"]
        out["reward"][t] = float(r)
        out["done"][t] = d
    s1 = R.snapshot(env)
    return dict(cfg=cfg, actions=actions, **out, obs0=obs0["image"], dir0=np.uint8(obs0["direction"]),
                grid0=s0["grid"], aux0=s0["aux"], agent0=s0["agent"], carrying0=s0["carrying"],
                grid1=s1["grid"], aux1=s1["aux"], agent1=s1["agent"], carrying1=s1["carrying"])


def save_scenes(name, env_id, scenes):
    cfg = scenes[0]["cfg"]
    d = dict(env_id=env_id, cfg_keys=np.array(list(cfg.keys())), cfg_vals=np.array(list(cfg.values()), np.int32))
    for k in scenes[0]:
        if k != "cfg":
            d[k] = np.stack([s[k] for s in scenes])
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    return path


HEADLINE = [
    # id, (n_long, T_long), (n_short, T_short)
    ("MiniGrid-Empty-8x8-v0", (2, 800), (6, 300)),
    ("MiniGrid-DoorKey-16x16-v0", (1, 7800), (7, 400)),
    ("MiniGrid-FourRooms-v0", (2, 1600), (6, 300)),
    ("MiniGrid-Dynamic-Obstacles-16x16-v0", (2, 1500), (6, 400)),
    ("MiniGrid-KeyCorridorS6R3-v0", (1, 3400), (11, 300)),
]
VARIANTS = [
    "MiniGrid-Empty-5x5-v0", "MiniGrid-Empty-6x6-v0", "MiniGrid-Empty-16x16-v0",
    "MiniGrid-Empty-Random-5x5-v0", "MiniGrid-Empty-Random-6x6-v0", "MiniGrid-Empty-Random-8x8-v0",
    "MiniGrid-DoorKey-5x5-v0", "MiniGrid-DoorKey-6x6-v0", "MiniGrid-DoorKey-8x8-v0",
    "MiniGrid-Dynamic-Obstacles-5x5-v0", "MiniGrid-Dynamic-Obstacles-Random-5x5-v0",
    "MiniGrid-Dynamic-Obstacles-6x6-v0", "MiniGrid-Dynamic-Obstacles-Random-6x6-v0",
    "MiniGrid-Dynamic-Obstacles-8x8-v0",
    "MiniGrid-KeyCorridorS3R1-v0", "MiniGrid-KeyCorridorS3R2-v0", "MiniGrid-KeyCorridorS3R3-v0",
    "MiniGrid-KeyCorridorS4R3-v0", "MiniGrid-KeyCorridorS5R3-v0",
    # on-device generators of round 2 (crossing.py, lavagap.py, multiroom.py): np_random.shuffle / choice / randint injected
    "MiniGrid-LavaCrossingS9N1-v0", "MiniGrid-LavaCrossingS9N0-v0", "MiniGrid-LavaCrossingS9N3-v0", "MiniGrid-LavaCrossingS11N5-v0",
    "MiniGrid-SimpleCrossingS9N2-v0", "MiniGrid-SimpleCrossingS11N5-v0",
    "MiniGrid-LavaGapS5-v0", "MiniGrid-LavaGapS7-v0", "MiniGrid-NormalGapS6-v0", "MiniGrid-LavaGapS6-v1",
    "MiniGrid-MultiRoom-N2-S4-v0", "MiniGrid-MultiRoom-N4-S5-v0", "MiniGrid-MultiRoom-N6-v0",
    "MiniGrid-DistShift1-v0", "MiniGrid-DistShift1-v1", "MiniGrid-DistShift2-v0",
]
# RNG-tape traces (the reference's own MT19937 draws through TapeRecorder.randint / shuffle / choice) for the same files
TAPE_VARIANTS = ["MiniGrid-LavaCrossingS9N2-v0", "MiniGrid-SimpleCrossingS9N3-v0", "MiniGrid-LavaGapS6-v0", "MiniGrid-MultiRoom-N6-v0"]


def short(env_id):
    return env_id.replace("MiniGrid-", "").replace("-v0", "").replace("-", "_").lower()


def main():
    os.makedirs(OUT, exist_ok=True)
    R.load_reference()
    t0 = time.time()
    seed = 20261018
    for env_id, (nl, Tl), (ns, Ts) in HEADLINE:
        for mode in ("philox", "tape"):
            # equal-length traces are stacked per file: long and short go to separate files
            for tag, n, T, base in (("long", nl, Tl, 0), ("short", ns, Ts, 100)):
                if mode == "tape" and tag == "short":
                    n = min(n, 3)
                idx = [base + 37 * k + (1 << 33) * (k == 1) for k in range(n)]   # one id beyond 32 bits
                acts = [1000 + base + k for k in range(n)]
                traces = [run_trace(env_id, seed, i, T, a, mode) for i, a in zip(idx, acts)]
                p = save_traces("%s_%s_%s" % (mode, short(env_id), tag), env_id, traces, seed, idx, acts)
                print("%-60s %7.1f KB  episodes=%s  [%.0fs]" % (os.path.basename(p), os.path.getsize(p) / 1024,
                      [t["n_episodes"] for t in traces], time.time() - t0), flush=True)
    for env_id in VARIANTS:
        idx = [5, 77, 4242]
        acts = [2000, 2001, 2002]
        traces = [run_trace(env_id, seed + 1, i, 300, a, "philox") for i, a in zip(idx, acts)]
        p = save_traces("philox_%s" % short(env_id), env_id, traces, seed + 1, idx, acts)
        print("%-60s %7.1f KB  episodes=%s  [%.0fs]" % (os.path.basename(p), os.path.getsize(p) / 1024,
              [t["n_episodes"] for t in traces], time.time() - t0), flush=True)
    for env_id in TAPE_VARIANTS:
        idx = [9, 31, 640]
        acts = [2100, 2101, 2102]
        traces = [run_trace(env_id, seed + 2, i, 300, a, "tape") for i, a in zip(idx, acts)]
        p = save_traces("tape_%s" % short(env_id), env_id, traces, seed + 2, idx, acts)
        print("%-60s %7.1f KB  episodes=%s  [%.0fs]" % (os.path.basename(p), os.path.getsize(p) / 1024,
              [t["n_episodes"] for t in traces], time.time() - t0), flush=True)
    # scenes: see-through (Empty classes) and occluded (DoorKey classes, FourRooms 19x19)
    for name, env_id, n, T, dens in (
        ("scenes_seethrough_8x8", "MiniGrid-Empty-8x8-v0", 48, 120, 0.30),
        ("scenes_seethrough_16x16", "MiniGrid-Empty-16x16-v0", 16, 200, 0.25),
        ("scenes_occluded_8x8", "MiniGrid-DoorKey-8x8-v0", 48, 120, 0.30),
        ("scenes_occluded_16x16", "MiniGrid-DoorKey-16x16-v0", 24, 200, 0.25),
        ("scenes_occluded_19x19", "MiniGrid-FourRooms-v0", 24, 200, 0.20),
        ("scenes_occluded_5x5", "MiniGrid-DoorKey-5x5-v0", 32, 80, 0.35),
    ):
        scenes = [run_scene(env_id, 9000 + k, T, dens) for k in range(n)]
        p = save_scenes(name, env_id, scenes)
        dn = sum(int(s["done"].sum()) for s in scenes)
        rw = sum(int((s["reward"] != 0).sum()) for s in scenes)
        print("%-60s %7.1f KB  done-steps=%d reward-steps=%d [%.0fs]" % (os.path.basename(p), os.path.getsize(p) / 1024,
              dn, rw, time.time() - t0), flush=True)


def wrapper_traces():
    """observation wrappers: outputs of the reference's own wrapper classes
    (FullyObsWrapper, FlatObsWrapper, FullyObsOneHotWrapper over ImgObsWrapper(FullyObsWrapper))
    evaluated on Philox-injected trajectories."""
    W = sys.modules["gym_minigrid.wrappers"]
    seed = 4242
    for env_id in ("MiniGrid-DoorKey-8x8-v0", "MiniGrid-KeyCorridorS3R3-v0", "MiniGrid-Empty-8x8-v0"):
        idx, T = [3, 11], 40
        full, flat, foh, acts, missions = [], [], [], [], []
        for k, i in enumerate(idx):
            env = R.make(env_id)
            shim = R.PhiloxShim(seed, i, 0)
            env.np_random = shim
            obs = env.reset()
            ep = 1
            w_full = W.FullyObsWrapper(env)
            w_flat = W.FlatObsWrapper(env)
            w_foh = W.FullyObsOneHotWrapper(W.ImgObsWrapper(W.FullyObsWrapper(env)), flatten=True)
            rs = np.random.RandomState(77 + k)
            a = rs.randint(0, env.action_space.n, size=T).astype(np.uint8)
            f1, f2, f3, ms = [], [], [], []
            for t in range(T):
                obs, r, d, _ = env.step(int(a[t]))
                if d:
                    shim.new_episode(ep)
                    obs = env.reset()
                    ep += 1
                fo = w_full.observation(obs)["image"]
                f1.append(fo.copy())
                f2.append(np.asarray(w_flat.observation(obs)).copy())
                f3.append(np.asarray(w_foh.observation(fo)).copy())
                ms.append(obs["mission"])
            full.append(np.stack(f1)); flat.append(np.stack(f2)); foh.append(np.stack(f3)); acts.append(a); missions.append(ms)
        path = os.path.join(OUT, "wrappers_%s.npz" % short(env_id))
        np.savez_compressed(path, env_id=env_id, seed=np.uint64(seed), env_indices=np.array(idx, np.int64),
                            actions=np.stack(acts), full=np.stack(full), flat=np.stack(flat).astype(np.float32),
                            full_onehot=np.stack(foh).astype(np.uint8), missions=np.array(missions))
        print("%-50s %7.1f KB flat dtype %s" % (os.path.basename(path), os.path.getsize(path) / 1024, flat[0].dtype))


def pool_traces():
    """Level-pool mode: envs whose step() is the base MiniGridEnv.step.  K layouts are
    generated BY THE REFERENCE (its own _gen_grid and RNG); the trace restores the layout that the shared
    Philox pick selects at every (auto-)reset and steps the unmodified reference."""
    import copy
    seed, K = 777, 6
    for env_id, T, n_env in (("MiniGrid-LavaCrossingS9N2-v0", 700, 3), ("MiniGrid-SimpleCrossingS11N5-v0", 600, 2),
                             ("MiniGrid-LavaGapS7-v1", 500, 3), ("MiniGrid-MultiRoom-N4-S5-v0", 400, 3),
                             ("MiniGrid-DistShift2-v0", 500, 2), ("MiniGrid-SimpleRoom-v0", 300, 2),
                             ("MiniGrid-Empty-6x6-v2", 400, 2), ("MiniGrid-Empty-Random-10x10-v0", 450, 2)):
        env = R.make(env_id)
        cfg = dict(config_of(env), gen=5, gen_param0=0, gen_param1=0)      # played from an uploaded pool, whatever generator the id has
        levels = []
        for k in range(K):
            env.seed(100 + k)
            obs = env.reset()
            s = R.snapshot(env)
            levels.append(dict(grid=s["grid"], aux=s["aux"], agent=s["agent"][:3].copy(), mission=obs["mission"],
                               obj=(copy.deepcopy(env.unwrapped.grid), tuple(int(v) for v in env.unwrapped.agent_pos), int(env.unwrapped.agent_dir))))

        def restore(lvl):
            u = env.unwrapped
            env.reset()
            g, pos, d = levels[lvl]["obj"]
            u.grid = copy.deepcopy(g)
            u.agent_pos = np.array(pos)
            u.agent_dir = d
            u.carrying = None
            u.step_count = 0
            return u.gen_obs()

        idx = [4, 90, (1 << 33) + 5][:n_env]
        tr = dict(obs=[], dir=[], reward=[], done=[], actions=[], obs0=[], dir0=[], lvl=[], grid_end=[], agent_end=[])
        for k, i in enumerate(idx):
            pick = R.PhiloxShim(seed, i, 0)
            lv = [pick.randint(0, K)]
            obs = restore(lv[-1])
            rs = np.random.RandomState(500 + k)
            a = rs.randint(0, 7, size=T).astype(np.uint8)
            O, D, RW, DN = [obs["image"].copy()], [obs["direction"]], [], []
            ep = 1
            for t in range(T):
                obs, r, d, _ = env.step(int(a[t]))
                if d:
                    pick.new_episode(ep)
                    ep += 1
                    lv.append(pick.randint(0, K))
                    obs = restore(lv[-1])
                O.append(obs["image"].copy()); D.append(obs["direction"]); RW.append(float(r)); DN.append(int(d))
            s = R.snapshot(env)
            tr["obs0"].append(O[0]); tr["dir0"].append(D[0]); tr["obs"].append(np.stack(O[1:])); tr["dir"].append(np.array(D[1:], np.uint8))
            tr["reward"].append(np.array(RW)); tr["done"].append(np.array(DN, np.uint8)); tr["actions"].append(a)
            tr["lvl"].append(np.array(lv + [-1] * (T + 1 - len(lv)), np.int32)); tr["grid_end"].append(s["grid"]); tr["agent_end"].append(s["agent"])
        path = os.path.join(OUT, "pool_%s.npz" % short(env_id).replace("-v1", "_v1").replace("-v2", "_v2"))
        np.savez_compressed(path, env_id=env_id, seed=np.uint64(seed), env_indices=np.array(idx, np.int64),
                            cfg_keys=np.array(list(cfg.keys())), cfg_vals=np.array(list(cfg.values()), np.int32),
                            level_grid=np.stack([l["grid"] for l in levels]), level_aux=np.stack([l["aux"] for l in levels]),
                            level_agent=np.stack([l["agent"] for l in levels]).astype(np.int32),
                            level_mission=np.array([l["mission"] for l in levels]),
                            **{k: np.stack(v) for k, v in tr.items()})
        print("%-50s %7.1f KB episodes=%s rewards=%d" % (os.path.basename(path), os.path.getsize(path) / 1024,
              [int((l >= 0).sum()) for l in tr["lvl"]], sum(int((r != 0).sum()) for r in tr["reward"])))


def viewsize_traces():
    """ViewSizeWrapper(env, V) (wrappers.py:579-608) on Philox-injected trajectories, V in {3, 5, 9, 11}."""
    W = sys.modules["gym_minigrid.wrappers"]
    seed = 31337
    for env_id, V in (("MiniGrid-DoorKey-8x8-v0", 5), ("MiniGrid-FourRooms-v0", 9), ("MiniGrid-Dynamic-Obstacles-8x8-v0", 5),
                      ("MiniGrid-KeyCorridorS3R3-v0", 3), ("MiniGrid-Empty-16x16-v0", 11), ("MiniGrid-DoorKey-16x16-v0", 11)):
        idx, T = [2, 40], 250
        tr = dict(obs=[], dir=[], reward=[], done=[], actions=[], obs0=[])
        for k, i in enumerate(idx):
            env = W.ViewSizeWrapper(R.make(env_id), V)
            assert env.observation_space.spaces["image"].shape == (V, V, 3)
            shim = R.PhiloxShim(seed, i, 0)
            env.unwrapped.np_random = shim
            obs = env.reset()
            ep = 1
            n_act = env.unwrapped.action_space.n
            a = np.random.RandomState(900 + k).randint(0, n_act, size=T).astype(np.uint8)
            O, D, RW, DN = [obs["image"].copy()], [], [], []
            for t in range(T):
                obs, r, d, _ = env.step(int(a[t]))
                if d:
                    shim.new_episode(ep)
                    ep += 1
                    obs = env.reset()
                O.append(obs["image"].copy()); D.append(obs["direction"]); RW.append(float(r)); DN.append(int(d))
            tr["obs0"].append(O[0]); tr["obs"].append(np.stack(O[1:])); tr["dir"].append(np.array(D, np.uint8))
            tr["reward"].append(np.array(RW)); tr["done"].append(np.array(DN, np.uint8)); tr["actions"].append(a)
        cfg = config_of(env)
        cfg["view_size"] = V
        path = os.path.join(OUT, "viewsize%d_%s.npz" % (V, short(env_id)))
        np.savez_compressed(path, env_id=env_id, seed=np.uint64(seed), env_indices=np.array(idx, np.int64), view=np.int32(V),
                            cfg_keys=np.array(list(cfg.keys())), cfg_vals=np.array(list(cfg.values()), np.int32),
                            **{k: np.stack(v) for k, v in tr.items()})
        print("%-50s %7.1f KB" % (os.path.basename(path), os.path.getsize(path) / 1024))


# ---------------------------------------------------------------------------
# level-pool envs with step() hooks
# ---------------------------------------------------------------------------
def _find(u, pred):
    for x in range(u.width):
        for y in range(u.height):
            c = u.grid.get(x, y)
            if c is not None and pred(c):
                return (x, y)
    return None


def hook_params_of(u, hook):
    """the per-level attributes each hook reads, as the 16 int32 of MGB_HOOK_PARAMS"""
    mg = sys.modules["gym_minigrid.minigrid"]
    T, C = mg.OBJECT_TO_IDX, mg.COLOR_TO_IDX
    hp = [0] * 16
    if hook == 1:
        hp[0], hp[1] = T[u.obj.type], C[u.obj.color]
    elif hook == 2:
        hp[6], hp[7] = _find(u, lambda c: c is u.door)
    elif hook == 3:
        hp[0], hp[1] = T[u.targetType], C[u.targetColor]
    elif hook == 4:
        hp[4], hp[5] = u.target_pos
        for k, (x, y) in enumerate(u.doorPos):
            hp[6 + 2 * k], hp[7 + 2 * k] = x, y
    elif hook == 5:
        hp[4], hp[5] = u.target_pos
    elif hook == 6:
        hp[2], hp[3] = T[u.move_type], C[u.moveColor]
        hp[4], hp[5] = u.target_pos
    elif hook == 7:
        hp[6], hp[7] = _find(u, lambda c: c is u.red_door)
        hp[8], hp[9] = _find(u, lambda c: c is u.blue_door)
    elif hook == 8:
        hp[6], hp[7] = u.success_pos
        hp[8], hp[9] = u.failure_pos
    return [int(v) for v in hp]


def _face(u, P, rs):
    """teleport the agent to an empty cell 4-adjacent to P, facing P; returns False if there is none"""
    mg = sys.modules["gym_minigrid.minigrid"]
    cand = []
    for d, (dx, dy) in enumerate([(1, 0), (0, 1), (-1, 0), (0, -1)]):
        x, y = P[0] - dx, P[1] - dy
        if 0 < x < u.width - 1 and 0 < y < u.height - 1 and u.grid.get(x, y) is None:
            cand.append((x, y, d))
    if not cand:
        return False
    x, y, d = cand[rs.randint(len(cand))]
    u.agent_pos = np.array([x, y])
    u.agent_dir = d
    return True


def directed_scenarios(u0, hook, rs):
    """(env copy, actions) pairs that drive each hook through its success AND failure branch"""
    import copy
    A = u0.actions
    out = []

    def fresh():
        u = copy.deepcopy(u0)
        u.step_count = int(rs.randint(0, u.max_steps - 5))
        return u

    tail = [int(v) for v in rs.randint(0, 7, size=3)]
    if hook == 1:
        u = fresh()
        if _face(u, _find(u, lambda c: c is u.obj), rs):
            out.append((u, [A.pickup] + tail))
        box = _find(u0, lambda c: c.type == "box" and c.contains is not None)
        if box:                                          # ObstructedMaze: keys hidden in boxes (Box.contains)
            for acts in ([A.toggle, A.pickup, A.drop], [A.pickup, A.left, A.drop, A.toggle, A.pickup], [A.pickup, A.right, A.right, A.drop]):
                u = fresh()
                if _face(u, box, rs):
                    out.append((u, list(acts) + tail))
    elif hook == 2:
        u = fresh()
        key = _find(u, lambda c: c.type == "key" and c.color == u.door.color)
        if key:
            u.carrying = u.grid.get(*key)
            u.grid.set(*key, None)
            if _face(u, _find(u, lambda c: c is u.door), rs):
                out.append((u, [A.toggle] + tail))
                u2 = copy.deepcopy(u)
                u2.carrying = None                       # no key: the toggle must not open it
                out.append((u2, [A.toggle] + tail))
    elif hook == 3:
        for want in (True, False):
            u = fresh()
            P = _find(u, lambda c: c.can_pickup() and ((c.type == u.targetType and c.color == u.targetColor) == want))
            if P and _face(u, P, rs):
                out.append((u, [A.pickup] + tail))
    elif hook == 4:
        for k in range(4):
            u = fresh()
            if _face(u, u.doorPos[k], rs):
                out.append((u, [A.done] + tail))
    elif hook == 5:
        u = fresh()
        if _face(u, u.target_pos, rs):
            out.append((u, [A.done] + tail))
            u2 = copy.deepcopy(u)
            out.append((u2, [A.toggle] + tail))
        u3 = fresh()
        out.append((u3, [A.done] + tail))                 # wherever the agent starts: done without reward (usually)
    elif hook == 6:
        for near in (True, False):
            u = fresh()
            mv = _find(u, lambda c: c.type == u.move_type and c.color == u.moveColor)
            if not mv:
                continue
            u.carrying = u.grid.get(*mv)
            u.grid.set(*mv, None)
            tx, ty = u.target_pos
            cells = [(x, y) for x in range(1, u.width - 1) for y in range(1, u.height - 1) if u.grid.get(x, y) is None
                     and ((abs(x - tx) <= 1 and abs(y - ty) <= 1) == near)]
            rs.shuffle(cells)
            for F in cells:
                if _face(u, F, rs):
                    out.append((u, [A.drop] + tail))
                    break
        u = fresh()                                      # wrong pickup
        P = _find(u, lambda c: c.can_pickup() and not (c.type == u.move_type and c.color == u.moveColor))
        if P and _face(u, P, rs):
            out.append((u, [A.pickup] + tail))
    elif hook == 7:
        for red_open in (True, False):
            u = fresh()
            u.red_door.is_open = red_open
            if _face(u, _find(u, lambda c: c is u.blue_door), rs):
                out.append((u, [A.toggle] + tail))
        u = fresh()
        u.blue_door.is_open = True                        # blue already open, then red
        u.red_door.is_open = False
        # (blue_after -> failure branch fires on any action)
        out.append((u, [A.left] + tail))
    elif hook == 8:
        for P in (u0.success_pos, u0.failure_pos):
            u = fresh()
            if _face(u, P, rs):
                out.append((u, [A.forward] + tail))
        u = fresh()
        out.append((u, [A.pickup, A.pickup] + tail))      # pickup acts as toggle
    return out


def hook_traces():
    """SURVEY §8f rank 2, env files whose step() adds a success/failure rule (MGB_HOOK_*).  Levels and their
    hook attributes come from the reference; random-action traces (pool restore by deepcopy of the reference
    env) plus directed scenarios that force every hook branch."""
    import copy
    seed, K = 909, 6
    for env_id, T in (("MiniGrid-Unlock-v0", 400), ("MiniGrid-UnlockPickup-v0", 300), ("MiniGrid-BlockedUnlockPickup-v0", 300),
                      ("MiniGrid-Fetch-8x8-N3-v0", 500), ("MiniGrid-GoToDoor-6x6-v0", 500), ("MiniGrid-GoToObject-8x8-N2-v0", 400),
                      ("MiniGrid-PutNear-8x8-N3-v0", 400), ("MiniGrid-RedBlueDoors-6x6-v0", 800), ("MiniGrid-MemoryS7-v0", 700),
                      ("MiniGrid-MemoryS13Random-v0", 300), ("MiniGrid-LockedRoom-v0", 300), ("MiniGrid-Playground-v0", 250),
                      ("MiniGrid-ObstructedMaze-1Dlhb-v0", 400), ("MiniGrid-ObstructedMaze-Full-v0", 300)):
        env = R.make(env_id)
        cfg = config_of(env)
        hook = cfg["hook"]
        levels = []
        for k in range(K):
            env.seed(200 + k)
            obs = env.reset()
            s = R.snapshot(env)
            # the reference's own FlatObsWrapper on the level's first observation: image ++ one-hot of THIS level's mission
            flat0 = np.asarray(sys.modules["gym_minigrid.wrappers"].FlatObsWrapper(env).observation(obs), np.float32)
            levels.append(dict(grid=s["grid"], aux=s["aux"], agent=s["agent"][:3].copy(), mission=obs["mission"], flat0=flat0,
                               hp=hook_params_of(env.unwrapped, hook), env=copy.deepcopy(env.unwrapped)))
        idx = [8, 123]
        tr = dict(obs=[], dir=[], reward=[], done=[], actions=[], obs0=[], dir0=[], lvl=[], grid_end=[], agent_end=[])
        for k, i in enumerate(idx):
            pick = R.PhiloxShim(seed, i, 0)
            lv = [pick.randint(0, K)]
            u = copy.deepcopy(levels[lv[-1]]["env"])
            obs = u.gen_obs()
            a = np.random.RandomState(600 + k).randint(0, 7, size=T).astype(np.uint8)
            O, D, RW, DN = [obs["image"].copy()], [obs["direction"]], [], []
            ep = 1
            for t in range(T):
                obs, r, d, _ = u.step(int(a[t]))
                if d:
                    pick.new_episode(ep)
                    ep += 1
                    lv.append(pick.randint(0, K))
                    u = copy.deepcopy(levels[lv[-1]]["env"])
                    obs = u.gen_obs()
                O.append(obs["image"].copy()); D.append(obs["direction"]); RW.append(float(r)); DN.append(int(d))
            s = R.snapshot(u)
            tr["obs0"].append(O[0]); tr["dir0"].append(D[0]); tr["obs"].append(np.stack(O[1:])); tr["dir"].append(np.array(D[1:], np.uint8))
            tr["reward"].append(np.array(RW)); tr["done"].append(np.array(DN, np.uint8)); tr["actions"].append(a)
            tr["lvl"].append(np.array(lv + [-1] * (T + 1 - len(lv)), np.int32)); tr["grid_end"].append(s["grid"]); tr["agent_end"].append(s["agent"])
        # directed scenarios (state upload, no reset)
        rs = np.random.RandomState(4321)
        sc = dict(sc_level=[], sc_grid=[], sc_aux=[], sc_agent=[], sc_carrying=[], sc_actions=[], sc_obs=[], sc_dir=[], sc_reward=[], sc_done=[])
        for li, L in enumerate(levels):
            for u, acts in directed_scenarios(L["env"], hook, rs):
                s0 = R.snapshot(u)
                O, D, RW, DN = [], [], [], []
                for a_ in acts:
                    obs, r, d, _ = u.step(int(a_))
                    O.append(obs["image"].copy()); D.append(obs["direction"]); RW.append(float(r)); DN.append(int(d))
                sc["sc_level"].append(li); sc["sc_grid"].append(s0["grid"]); sc["sc_aux"].append(s0["aux"]); sc["sc_agent"].append(s0["agent"])
                sc["sc_carrying"].append(s0["carrying"]); sc["sc_actions"].append(np.array(acts, np.uint8)); sc["sc_obs"].append(np.stack(O))
                sc["sc_dir"].append(np.array(D, np.uint8)); sc["sc_reward"].append(np.array(RW)); sc["sc_done"].append(np.array(DN, np.uint8))
        maxA = max([len(a_) for a_ in sc["sc_actions"]] + [1])
        sc_len = np.array([len(a_) for a_ in sc["sc_actions"]], np.int32)        # true lengths (before padding)
        for key in ("sc_actions", "sc_obs", "sc_dir", "sc_reward", "sc_done"):       # pad to a common length with action 6 (done = no-op for base)
            sc[key] = [np.concatenate([v, np.repeat(v[-1:], maxA - len(v), axis=0)]) if len(v) < maxA else v for v in sc[key]]
        path = os.path.join(OUT, "hook_%s.npz" % short(env_id))
        extra = {k: (np.stack(v) if len(v) else np.zeros((0,))) for k, v in sc.items()}
        np.savez_compressed(path, env_id=env_id, seed=np.uint64(seed), env_indices=np.array(idx, np.int64),
                            cfg_keys=np.array(list(cfg.keys())), cfg_vals=np.array(list(cfg.values()), np.int32),
                            level_grid=np.stack([l["grid"] for l in levels]), level_aux=np.stack([l["aux"] for l in levels]),
                            level_agent=np.stack([l["agent"] for l in levels]).astype(np.int32),
                            level_hook=np.array([l["hp"] for l in levels], np.int32),
                            level_mission=np.array([l["mission"] for l in levels]), sc_len=sc_len,
                            level_flat=np.stack([l["flat0"] for l in levels]),
                            **{k: np.stack(v) for k, v in tr.items()}, **extra)
        print("%-44s %6.1f KB hook=%d episodes=%s rewards(random)=%d scenarios=%d (reward>0: %d, done: %d)" % (
            os.path.basename(path), os.path.getsize(path) / 1024, hook, [int((l >= 0).sum()) for l in tr["lvl"]],
            sum(int((r != 0).sum()) for r in tr["reward"]), len(sc["sc_level"]),
            sum(int((r[:1] > 0).sum()) for r in sc["sc_reward"]), sum(int(d[0]) for d in sc["sc_done"])))


def rgb_traces():
    """RGBImgPartialObsWrapper / RGBImgObsWrapper (wrappers.py:245-309) outputs of the reference on Philox-injected
    trajectories: pixel-exact targets for the tile-atlas kernels."""
    W = sys.modules["gym_minigrid.wrappers"]
    seed = 2468
    for env_id in ("MiniGrid-DoorKey-8x8-v0", "MiniGrid-KeyCorridorS3R3-v0", "MiniGrid-Dynamic-Obstacles-8x8-v0", "MiniGrid-FourRooms-v0"):
        idx, T = [1, 17], 24
        part, full, fullhl, acts = [], [], [], []
        for k, i in enumerate(idx):
            env = R.make(env_id)
            shim = R.PhiloxShim(seed, i, 0)
            env.np_random = shim
            obs = env.reset()
            wp, wf = W.RGBImgPartialObsWrapper(env), W.RGBImgObsWrapper(env)
            a = np.random.RandomState(300 + k).randint(0, env.action_space.n, size=T).astype(np.uint8)
            P, F = [wp.observation(obs)["image"].copy()], [wf.observation(obs)["image"].copy()]
            H = [env.render('rgb_array', highlight=True, tile_size=8).copy()]        # MiniGridEnv.render, view highlighted
            ep = 1
            for t in range(T):
                obs, r, d, _ = env.step(int(a[t]))
                if d:
                    shim.new_episode(ep)
                    ep += 1
                    obs = env.reset()
                P.append(wp.observation(obs)["image"].copy()); F.append(wf.observation(obs)["image"].copy())
                H.append(env.render('rgb_array', highlight=True, tile_size=8).copy())
            part.append(np.stack(P)); full.append(np.stack(F)); fullhl.append(np.stack(H)); acts.append(a)
        path = os.path.join(OUT, "rgb_%s.npz" % short(env_id))
        np.savez_compressed(path, env_id=env_id, seed=np.uint64(seed), env_indices=np.array(idx, np.int64), actions=np.stack(acts),
                            partial=np.stack(part), full=np.stack(full), full_highlight=np.stack(fullhl))
        print("%-44s %6.1f KB partial %s full %s" % (os.path.basename(path), os.path.getsize(path) / 1024, part[0].shape, full[0].shape))


def bookkeeping_traces():
    """bookkeeping wrappers: the reference's own ActionBonus, StateBonus, DACWrapper,
    AppendActionWrapper, GoalPolicyWrapper and AgentExtraInfoWrapper on Philox-injected trajectories.  A finished
    episode is followed by reset() of the same wrapper (the batched wrappers auto-reset), except under DACWrapper,
    which runs one fixed-length episode."""
    W = sys.modules["gym_minigrid.wrappers"]
    seed = 5151
    for env_id, T in (("MiniGrid-Empty-5x5-v0", 230), ("MiniGrid-DoorKey-5x5-v0", 300), ("MiniGrid-Dynamic-Obstacles-5x5-v0", 120),
                      ("MiniGrid-DoorKey-8x8-v0", 60)):
        idx = [2, 9]
        out = {k: [] for k in ("actions", "ab_reward", "ab_done", "ab_image", "sb_reward", "sb_done", "dac_image", "dac_dir",
                               "dac_reward", "dac_done", "app_obs", "gp_obs", "gp_achieved", "gp_desired", "gp_reward",
                               "xi_pos", "xi_dir", "xi_map", "xi_full")}
        for k, i in enumerate(idx):
            def fresh():
                env = R.make(env_id)
                shim = R.PhiloxShim(seed, i, 0)
                env.np_random = shim
                return env, shim
            n_act = R.make(env_id).action_space.n
            a = np.random.RandomState(500 + k).randint(0, n_act, size=T).astype(np.uint8)
            out["actions"].append(a)
            # --- ActionBonus / StateBonus: the counts survive reset() ---
            for tag, cls in (("ab", W.ActionBonus), ("sb", W.StateBonus)):
                env, shim = fresh()
                w = cls(env)
                w.reset()
                ep, rr, dd, im = 1, [], [], []
                for t in range(T):
                    obs, r, d, _ = w.step(int(a[t]))
                    if d:
                        shim.new_episode(ep); ep += 1
                        obs = w.reset()
                    rr.append(float(r)); dd.append(bool(d)); im.append(obs["image"].copy())
                out[tag + "_reward"].append(np.array(rr, np.float64)); out[tag + "_done"].append(np.array(dd))
                if tag == "ab":
                    out["ab_image"].append(np.stack(im))
            # --- DACWrapper: one episode of exactly max_steps steps (plus a few steps beyond) ---
            env, shim = fresh()
            w = W.DACWrapper(env)
            w.reset()
            Td = min(T, env.max_steps + 4)
            im, di, rr, dd = [], [], [], []
            for t in range(Td):
                obs, r, d, _ = w.step(int(a[t % T]))
                im.append(np.asarray(obs["image"]).copy()); di.append(int(obs["direction"])); rr.append(float(r)); dd.append(bool(d))
            pad = T - Td
            out["dac_image"].append(np.concatenate([np.stack(im), np.zeros((pad,) + im[0].shape, np.uint8)]))
            out["dac_dir"].append(np.array(di + [0] * pad, np.uint8)); out["dac_reward"].append(np.array(rr + [0.0] * pad, np.float64))
            out["dac_done"].append(np.array(dd + [False] * pad))
            dac_len = Td
            # --- AppendActionWrapper(K=3) over the flat full one-hot observation ---
            env, shim = fresh()
            w = W.AppendActionWrapper(W.FullyObsOneHotWrapper(W.ImgObsWrapper(W.FullyObsWrapper(env)), flatten=True), 3)
            w.reset()
            ep, oo = 1, []
            for t in range(T):
                obs, r, d, _ = w.step(int(a[t]))
                if d:
                    shim.new_episode(ep); ep += 1
                    obs = w.reset()
                oo.append(np.asarray(obs).copy())
            out["app_obs"].append(np.stack(oo).astype(np.uint8))
            # --- GoalPolicyWrapper + AgentExtraInfoWrapper ---
            env, shim = fresh()
            w = W.GoalPolicyWrapper(W.FullyObsOneHotWrapper(W.ImgObsWrapper(W.FullyObsWrapper(env)), flatten=True))
            xi = W.AgentExtraInfoWrapper(env)
            w.reset()
            ep, o1, o2, o3, o4, p1, p2, p3, p4 = 1, [], [], [], [], [], [], [], []
            for t in range(T):
                obs, r, d, _ = w.step(int(a[t]))
                if d:
                    shim.new_episode(ep); ep += 1
                    obs = w.reset()
                o1.append(obs["observation"].copy()); o2.append(obs["achieved_goal"].copy()); o3.append(obs["desired_goal"].copy())
                o4.append(float(w.compute_reward(None, None, None)))
                x = xi.observation({})
                p1.append(np.array(x["pos"], np.int32)); p2.append(int(x["dir"])); p3.append(xi.get_map().copy()); p4.append(xi.get_full_map().copy())
            out["gp_obs"].append(np.stack(o1).astype(np.uint8)); out["gp_achieved"].append(np.stack(o2).astype(np.uint8))
            out["gp_desired"].append(np.stack(o3).astype(np.uint8)); out["gp_reward"].append(np.array(o4, np.float64))
            out["xi_pos"].append(np.stack(p1)); out["xi_dir"].append(np.array(p2, np.int32)); out["xi_map"].append(np.stack(p3).astype(np.uint8))
            out["xi_full"].append(np.stack(p4).astype(np.uint8))
        path = os.path.join(OUT, "bookkeeping_%s.npz" % short(env_id))
        np.savez_compressed(path, env_id=env_id, seed=np.uint64(seed), env_indices=np.array(idx, np.int64), dac_len=np.int64(dac_len),
                            **{k: np.stack(v) for k, v in out.items()})
        nd = int(np.stack(out["ab_done"]).sum())
        print("%-50s %7.1f KB  done-steps=%d dac done at %s  max bonus count>1: %s" % (
            os.path.basename(path), os.path.getsize(path) / 1024, nd, [int(np.argmax(x)) for x in out["dac_done"]],
            bool((np.stack(out["sb_reward"]) < 0.6).any())))


def dynobs_boxed_traces():
    """Directed Dynamic-Obstacles cases: a ball walled in completely (its place_obj makes 101 tries =
    202 draws every step and it stays), a ball with exactly one free neighbour, a ball whose window is clipped by the
    grid border, and the agent standing next to a ball (the 'not onto the agent' rejection).  Philox-injected; the draw
    counter after every step is part of the fixture."""
    mg = sys.modules["gym_minigrid.minigrid"]
    seed = 7171
    for env_id in ("MiniGrid-Dynamic-Obstacles-8x8-v0", "MiniGrid-Dynamic-Obstacles-16x16-v0"):
        cases = []
        for k in range(6):
            env = R.make(env_id)
            u = env.unwrapped
            shim = R.PhiloxShim(seed, k, 0)
            u.np_random = shim
            env.reset()
            rs = np.random.RandomState(100 + k)
            W, H = u.width, u.height

            def free_nb(b):
                bx, by = (int(v) for v in b.cur_pos)
                out = []
                for x in range(bx - 1, bx + 2):
                    for y in range(by - 1, by + 2):
                        if (x, y) != (bx, by) and 0 < x < W - 1 and 0 < y < H - 1 and u.grid.get(x, y) is None \
                                and (x, y) != tuple(int(v) for v in u.agent_pos):
                            out.append((x, y))
                return out
            # ball 0: walled in completely; ball 1: one free neighbour left
            for bi, leave in ((0, 0), (1, 1)):
                nb = free_nb(u.obstacles[bi])
                rs.shuffle(nb)
                for (x, y) in nb[leave:]:
                    u.grid.set(x, y, mg.Wall())
            # agent next to the last ball when possible
            nb = free_nb(u.obstacles[-1])
            if nb and k % 2 == 0:
                u.agent_pos = np.array(nb[rs.randint(0, len(nb))])
                u.agent_dir = int(rs.randint(0, 4))
            s0 = R.snapshot(env)
            nd0 = shim.ndraws
            T = 10
            actions = rs.randint(0, 2, size=T).astype(np.uint8)          # turn only: the episode goes on
            if k == 5:
                actions[-1] = 2
            o = dict(obs=np.zeros((T, 7, 7, 3), np.uint8), dir=np.zeros(T, np.uint8), reward=np.zeros(T, np.float64),
                     done=np.zeros(T, np.uint8), ndraws=np.zeros(T, np.int64), obstacles=np.zeros((T, MAX_OBST, 2), np.int16))
            for t in range(T):
                ob, r, d, _ = env.step(int(actions[t]))
                o["obs"][t] = ob["image"]; o["dir"][t] = ob["direction"]; o["reward"][t] = float(r); o["done"][t] = d
                o["ndraws"][t] = shim.ndraws
                o["obstacles"][t] = pad_obst(R.snapshot(env)["obstacles"])
            s1 = R.snapshot(env)
            cases.append(dict(cfg=config_of(env), actions=actions, env_index=np.int64(k), seed=np.uint64(seed), ndraws0=np.int64(nd0), **o,
                              grid0=s0["grid"], aux0=s0["aux"], agent0=s0["agent"], carrying0=s0["carrying"],
                              obstacles0=pad_obst(s0["obstacles"]), grid1=s1["grid"], agent1=s1["agent"]))
        path = save_scenes("dynobs_boxed_%s" % short(env_id), env_id, cases)
        per_step = np.diff(np.stack([c["ndraws"] for c in cases]), axis=1)
        print("%-50s %6.1f KB seed %d draws/step min %d max %d" % (os.path.basename(path), os.path.getsize(path) / 1024, seed,
                                                                 per_step.min(), per_step.max()))


def helper_traces():
    """MiniGridEnv geometry helpers (minigrid.py:1092-1225) evaluated by the reference on Philox-injected trajectories:
    dir_vec, right_vec, front/left/right_pos, get_view_exts, and get_view_coords / in_view / agent_sees for EVERY cell."""
    seed = 3131
    for env_id in ("MiniGrid-DoorKey-8x8-v0", "MiniGrid-KeyCorridorS3R3-v0", "MiniGrid-Dynamic-Obstacles-6x6-v0"):
        idx, T = [4, 12], 30
        out = {k: [] for k in ("actions", "dir_vec", "right_vec", "front_pos", "left_pos", "right_pos", "view_exts", "view_coords",
                               "in_view", "agent_sees")}
        for k, i in enumerate(idx):
            env = R.make(env_id)
            u = env.unwrapped
            shim = R.PhiloxShim(seed, i, 0)
            u.np_random = shim
            env.reset()
            a = np.random.RandomState(900 + k).randint(0, env.action_space.n, size=T).astype(np.uint8)
            rec = {k2: [] for k2 in out if k2 != "actions"}
            ep = 1

            def snap():
                W_, H_ = u.width, u.height
                rec["dir_vec"].append(np.array(u.dir_vec, np.int32)); rec["right_vec"].append(np.array(u.right_vec, np.int32))
                rec["front_pos"].append(np.array(u.front_pos, np.int32)); rec["left_pos"].append(np.array(u.left_pos, np.int32))
                rec["right_pos"].append(np.array(u.right_pos, np.int32)); rec["view_exts"].append(np.array(u.get_view_exts(), np.int32))
                vc, iv, sees = np.zeros((W_, H_, 2), np.int32), np.zeros((W_, H_), bool), np.zeros((W_, H_), bool)
                for x in range(W_):
                    for y in range(H_):
                        vc[x, y] = u.get_view_coords(x, y)
                        iv[x, y] = u.in_view(x, y)
                        try:
                            sees[x, y] = bool(u.agent_sees(x, y))
                        except AttributeError:            # world cell None under a carried object (minigrid.py:1223-1225)
                            sees[x, y] = False
                rec["view_coords"].append(vc); rec["in_view"].append(iv); rec["agent_sees"].append(sees)
            snap()
            for t in range(T):
                _, _, d, _ = env.step(int(a[t]))
                if d:
                    shim.new_episode(ep); ep += 1
                    env.reset()
                snap()
            out["actions"].append(a)
            for k2 in rec:
                out[k2].append(np.stack(rec[k2]))
        path = os.path.join(OUT, "helpers_%s.npz" % short(env_id))
        np.savez_compressed(path, env_id=env_id, seed=np.uint64(seed), env_indices=np.array(idx, np.int64), **{k2: np.stack(v) for k2, v in out.items()})
        print("%-44s %6.1f KB  agent_sees true: %d" % (os.path.basename(path), os.path.getsize(path) / 1024, int(np.stack(out["agent_sees"]).sum())))


def reward_table():
    """_reward() (minigrid.py:933-937) evaluated BY THE REFERENCE for every step_count of every
    max_steps in the registry (and 50 beyond): r_<max_steps>[k] = reward at step_count == k."""
    d = {}
    for env_id in [h[0] for h in HEADLINE] + VARIANTS:
        env = R.make(env_id).unwrapped
        ms = env.max_steps
        if "r_%d" % ms in d:
            continue
        out = np.zeros(ms + 51, np.float64)
        for k in range(ms + 51):
            env.step_count = k
            out[k] = env._reward()
        d["r_%d" % ms] = out
    np.savez_compressed(os.path.join(OUT, "reward_table.npz"), **d)
    print("reward_table.npz", sorted(d.keys()))


if __name__ == "__main__":
    R.load_reference()
    if "--bookkeeping-only" in sys.argv:
        bookkeeping_traces()
        sys.exit(0)
    if "--helpers-only" in sys.argv:
        helper_traces()
        sys.exit(0)
    if "--rgb-only" in sys.argv:
        rgb_traces()
        sys.exit(0)
    if "--dynobs-boxed-only" in sys.argv:
        dynobs_boxed_traces()
        sys.exit(0)
    if "--extras-only" not in sys.argv:
        main()
    reward_table()
    wrapper_traces()
    pool_traces()
    viewsize_traces()
    hook_traces()
    rgb_traces()
    bookkeeping_traces()
    dynobs_boxed_traces()
    helper_traces()
