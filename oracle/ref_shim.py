"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Runs the *unmodified* reference (rohitrango/gym-minigrid, mounted read-only at
/root/reference) inside this container, where `gym`, `matplotlib` and `skimage`
are not installed.  We inject ~100 lines of stub modules into ``sys.modules``
(only the surface the reference touches at import time and on the hot path:
gym_minigrid/minigrid.py:2,5-7; wrappers.py; register.py:1) and then import
``gym_minigrid`` from the reference tree.  No reference source is copied.

The reference tree does NOT travel to the GPU box; everything here is used to
(a) generate the committed fixtures under tests/golden/ (oracle/gen_golden.py)
and (b) live-fuzz the C oracle against the reference in the CPU test-suite
(skipped when the tree is absent).

Also defines ``PhiloxShim``: a drop-in for ``env.np_random`` exposing the only
RNG call the five BASELINE configs make, ``randint(low, high)``
(minigrid.py:944,999-1000), backed by the same counter-based Philox4x32-10
stream + bounded-int mapping the CUDA kernel and the C oracle use, and
``TapeRecorder`` which wraps the reference's own RandomState and records every
draw (RNG-tape parity mode).
"""
import importlib
import os
import sys
import types

import numpy as np

# search order (SURVEY headline fact 4): $MGB_REFERENCE -> /root/reference (the build container) -> baseline/_ref (the
# unmodified reference pip-installed there by `python -m pip install --no-index --no-build-isolation --no-deps --target
# baseline/_ref <copy of /root/reference>`; git-ignored, but it travels to the GPU box with the repo snapshot)
REFERENCE_ROOT_CANDIDATES = [
    os.environ.get("MGB_REFERENCE", ""),
    "/root/reference",
    os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref"),
]


def reference_root():
    for p in REFERENCE_ROOT_CANDIDATES:
        if p and os.path.isdir(os.path.join(p, "gym_minigrid")):
            return p
    return None


def reference_available():
    return reference_root() is not None


# --------------------------------------------------------------------------
# gym / matplotlib / skimage stubs
# --------------------------------------------------------------------------
class _Space:
    pass


class _Box(_Space):
    def __init__(self, low, high, shape=None, dtype=None):
        self.low, self.high, self.dtype = low, high, dtype
        self.shape = tuple(shape) if shape is not None else np.shape(low)


class _Discrete(_Space):
    def __init__(self, n):
        self.n = int(n)

    def sample(self):
        return int(np.random.randint(self.n))


class _Dict(_Space):
    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def __getitem__(self, k):
        return self.spaces[k]


class _Env:
    metadata = {}
    reward_range = (-float("inf"), float("inf"))

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


class _Wrapper(_Env):
    def __init__(self, env):
        self.env = env
        self.action_space = getattr(env, "action_space", None)
        self.observation_space = getattr(env, "observation_space", None)
        self.reward_range = getattr(env, "reward_range", None)
        self.metadata = getattr(env, "metadata", None)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kw):
        return self.env.reset(**kw)

    def render(self, *a, **kw):
        return self.env.render(*a, **kw)

    def seed(self, seed=None):
        return self.env.seed(seed)

    def close(self):
        return self.env.close()


class _ObservationWrapper(_Wrapper):
    def reset(self, **kw):
        return self.observation(self.env.reset(**kw))

    def step(self, action):
        obs, r, d, info = self.env.step(action)
        return self.observation(obs), r, d, info


_REGISTRY = {}


def _register(id, entry_point=None, reward_threshold=None, **kw):
    _REGISTRY[id] = entry_point


def _make(id, **kw):
    mod, cls = _REGISTRY[id].split(":")
    return getattr(importlib.import_module(mod), cls)(**kw)


def _np_random(seed=None):
    # gym<=0.21 hashes the seed before seeding MT19937; the exact stream is
    # third-party arithmetic we never reproduce (parity goes through RNG
    # injection or the RNG tape), so a plain RandomState is sufficient here.
    rs = np.random.RandomState(None if seed is None else int(seed) % (2 ** 32))
    return rs, seed


def install_stubs():
    """Idempotent: safe across fork and repeated calls."""
    if "gym" in sys.modules and getattr(sys.modules["gym"], "_mgb_stub", False):
        return
    gym = types.ModuleType("gym")
    gym._mgb_stub = True
    core = types.ModuleType("gym.core")
    spaces = types.ModuleType("gym.spaces")
    error = types.ModuleType("gym.error")
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")

    class GoalEnv(_Env):
        pass

    for m in (gym, core):
        m.Env, m.Wrapper, m.ObservationWrapper = _Env, _Wrapper, _ObservationWrapper
    core.GoalEnv = GoalEnv
    gym.GoalEnv = GoalEnv
    spaces.Box, spaces.Discrete, spaces.Dict, spaces.Space = _Box, _Discrete, _Dict, _Space
    seeding.np_random = _np_random
    utils.seeding = seeding
    registration.register = _register
    envs.registration = registration
    gym.core, gym.spaces, gym.error, gym.utils, gym.envs = core, spaces, error, utils, envs
    gym.make = _make
    gym.register = _register
    mods = {
        "gym": gym, "gym.core": core, "gym.spaces": spaces, "gym.error": error,
        "gym.utils": utils, "gym.utils.seeding": seeding, "gym.envs": envs,
        "gym.envs.registration": registration,
    }
    for name in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.measure"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                mods[name] = types.ModuleType(name)
    if "matplotlib" in mods:
        mods["matplotlib"].pyplot = mods.get("matplotlib.pyplot")
    if "skimage" in mods:
        mods["skimage"].measure = mods.get("skimage.measure")
    sys.modules.update(mods)


_ref = None


def load_reference():
    """Import the untouched reference package; returns the `gym_minigrid` module."""
    global _ref
    if _ref is not None:
        return _ref
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not found (looked at $MGB_REFERENCE, /root/reference)")
    install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _ref = importlib.import_module("gym_minigrid")
    return _ref


def make(env_id, **kw):
    load_reference()
    import warnings
    import io
    import contextlib
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        return sys.modules["gym"].make(env_id, **kw)


# --------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., SC'11) -- the one RNG shared by kernel,
# oracle and the injected shim.  Pure-python ints; only used for fixtures.
# --------------------------------------------------------------------------
_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


class PhiloxShim:
    """Assigned to ``env.np_random`` of a reference env *after* construction.

    Stream = (seed, global env id, episode).  Every Philox word serves two consecutive draws: draw n uses word
    (n>>1)&3 of block n>>3, as it is for even n and times DRAW_ODD_MULT (mod 2^32) for odd n;
    ``randint(low, high) = low + mulhi32(u32, high-low)``.
    """

    DRAW_ODD_MULT = 0x9E3779B1

    def __init__(self, seed, env_id, episode=0):
        self.key = (seed & _MASK, (seed >> 32) & _MASK)
        self.env_id = int(env_id)
        self.episode = int(episode)
        self.ndraws = 0
        self._blk = None
        self._blk_idx = -1

    def new_episode(self, episode):
        self.episode = int(episode)
        self.ndraws = 0
        self._blk_idx = -1

    def _u32(self):
        b = self.ndraws >> 3
        if b != self._blk_idx:
            ctr = (b & _MASK, self.episode & _MASK, self.env_id & _MASK, (self.env_id >> 32) & _MASK)
            self._blk = philox4x32_10(ctr, self.key)
            self._blk_idx = b
        v = self._blk[(self.ndraws >> 1) & 3]
        if self.ndraws & 1:
            v = (v * self.DRAW_ODD_MULT) & _MASK
        self.ndraws += 1
        return v

    def randint(self, low, high=None):
        if high is None:
            low, high = 0, low
        span = int(high) - int(low)
        assert span > 0
        return int(low) + ((self._u32() * span) >> 32)

    # The two other RandomState methods the stock env files call (crossing.py:53,66,75-80), expressed through randint so
    # that the device generators can mirror them draw for draw: numpy's own algorithms (Fisher-Yates from the back;
    # one bounded integer for a 1-D choice) on this object's stream.
    def shuffle(self, x):
        _shuffle(self, x)

    def choice(self, a):
        return _choice(self, a)


def _shuffle(rng, x):
    for i in reversed(range(1, len(x))):
        j = rng.randint(0, i + 1)
        x[i], x[j] = x[j], x[i]


def _choice(rng, a):
    a = list(range(a)) if isinstance(a, (int, np.integer)) else list(a)
    return a[rng.randint(0, len(a))]


class TapeRecorder:
    """Wraps the reference's own RandomState; records every ``randint`` result.  shuffle / choice run the same
    algorithms as PhiloxShim on the recorded draws."""

    def __init__(self, rs):
        self.rs = rs
        self.tape = []

    def randint(self, low, high=None):
        v = int(self.rs.randint(low, high))
        self.tape.append(v)
        return v

    def shuffle(self, x):
        _shuffle(self, x)

    def choice(self, a):
        return _choice(self, a)


# --------------------------------------------------------------------------
# State snapshot of a reference env
# --------------------------------------------------------------------------
def _box_aux(c):
    """aux bits 1-3 of a Box: 0 = contains None, k+1 = contains Key(colour k).  Only what obstructedmaze.py:68-73
    builds is representable: a grey box holding a default key."""
    mg = sys.modules["gym_minigrid.minigrid"]
    if c.triage_color is not None or c.toggletimes != 1:
        raise NotImplementedError("non-default box")
    if c.contains is None:
        return 0
    k = c.contains
    if k.type != "key" or k.contains is not None or c.color != "grey":
        raise NotImplementedError("box contents other than a key in a grey box")
    return (mg.COLOR_TO_IDX[k.color] + 1) << 1


def snapshot(env):
    """-> dict(grid uint8[W,H,3], aux uint8[W,H], agent int32[4]=(x,y,dir,step_count),
    carrying uint8[3] (0,0,0 = none), obstacles int16[n,2], target uint8[2]=(type,color))"""
    env = env.unwrapped
    mg = sys.modules["gym_minigrid.minigrid"]
    W, H = env.grid.width, env.grid.height
    grid = env.grid.encode()
    aux = np.zeros((W, H), np.uint8)
    for x in range(W):
        for y in range(H):
            c = env.grid.get(x, y)
            if c is None:
                continue
            if c.type == "goal":
                if c.triage_color is not None:
                    raise NotImplementedError("triage_color")
                if c.overlap:
                    aux[x, y] |= 1
                if c.toggletimes not in (0, 1) or (c.toggletimes <= 0) != bool(c.overlap):
                    raise NotImplementedError("goal toggletimes")
            if c.type == "box":
                aux[x, y] |= _box_aux(c)
    carrying = np.zeros(3, np.uint8)
    if env.carrying is not None:
        carrying[:] = env.carrying.encode()
        if env.carrying.type == "box":
            carrying[2] = _box_aux(env.carrying)        # a carried box keeps its contents (state byte is otherwise 0)
    obst = np.zeros((0, 2), np.int16)
    if hasattr(env, "obstacles"):
        obst = np.array([tuple(int(v) for v in o.cur_pos) for o in env.obstacles], np.int16).reshape(-1, 2)
    target = np.zeros(2, np.uint8)
    if hasattr(env, "obj") and hasattr(env, "room_grid"):
        target[:] = (mg.OBJECT_TO_IDX[env.obj.type], mg.COLOR_TO_IDX[env.obj.color])
    agent = np.array([int(env.agent_pos[0]), int(env.agent_pos[1]), int(env.agent_dir), int(env.step_count)], np.int32)
    return dict(grid=grid, aux=aux, agent=agent, carrying=carrying, obstacles=obst, target=target)
