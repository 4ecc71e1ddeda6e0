"""Batched observation wrappers: SURVEY §8(f) rank 1 -- the part of gym_minigrid/wrappers.py that
the north star names as kept API surface.  Same class names and constructor arguments as the
reference; they wrap a VecMiniGridEnv and transform whole batches on the GPU through libmgb200
(mgb_full_obs / mgb_onehot / mgb_flat_obs).

Built (with reference lines):      ReseedWrapper (wrappers.py:12-32), ImgObsWrapper (:156-166),
    OneHotPartialObsWrapper (:203-243), FullyObsWrapper (:311-338), FullyObsOneHotWrapper (:340-415),
    FlatObsWrapper (:528-577).
    ViewSizeWrapper (:579-608; odd sizes 3..11, the kernel is a template on the view size).
    RGBImgPartialObsWrapper (:283-309), RGBImgObsWrapper (:245-281): tile-atlas gathers, pixel-exact (tile_size 8).
    Bookkeeping wrappers: DACWrapper (:35-84), ActionBonus (:87-119), StateBonus (:121-154),
    AgentExtraInfoWrapper (:169-200), AppendActionWrapper (:418-458), GoalPolicyWrapper (:461-526).
    ActionBonus / StateBonus / DACWrapper need the TERMINAL state of a finished env (its last position, or that it
    stays finished), so they switch the wrapped env's in-kernel auto-reset off and do it themselves afterwards.

Reference quirks, kept or documented:
  * OneHotPartialObsWrapper.observation reads `self.observation_space.shape` of a Dict space
    (wrappers.py:228), which has no usable shape, so the reference class cannot actually run; the
    intended (7, 7, 21) encoding of its loop body (:230-238) is what is implemented.
  * FlatObsWrapper concatenates a uint8 image with a float32 one-hot, so the result is float32 and
    1-D although its declared space is uint8 with shape (1, n) (wrappers.py:543-548,575).  Kept.
  * FullyObsOneHotWrapper uses 4 state planes (wrappers.py:368) and expects an *array* observation
    (e.g. ImgObsWrapper(FullyObsWrapper(env))).  Kept.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, spaces
from .vec_env import COLOR_TO_IDX, OBJECT_TO_IDX, STATE_TO_IDX, IDX_TO_COLOR, IDX_TO_OBJECT


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class Wrapper:
    """gym.core.Wrapper for batched envs: forwards everything it does not override."""

    def __init__(self, env):
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space
        self.reward_range = env.reward_range
        self.metadata = getattr(env, "metadata", {})

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self, **kw):
        return self.env.reset(**kw)

    def step(self, action):
        return self.env.step(action)

    def seed(self, seed=None):
        return self.env.seed(seed)

    def close(self):
        return self.env.close()


class ObservationWrapper(Wrapper):
    def reset(self, **kw):
        return self.observation(self.env.reset(**kw))

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        return self.observation(obs), reward, done, info

    def observation(self, obs):
        raise NotImplementedError


class ReseedWrapper(Wrapper):
    """wrappers.py:12-32: every reset re-seeds with the next seed of the list, so the whole batch
    regenerates the same layouts (per global env id) each time."""

    def __init__(self, env, seeds=[0], seed_idx=0):
        self.seeds = list(seeds)
        self.seed_idx = seed_idx
        super().__init__(env)

    def reset(self, **kw):
        seed = self.seeds[self.seed_idx]
        self.seed_idx = (self.seed_idx + 1) % len(self.seeds)
        self.env.seed(seed)
        return self.env.reset(**kw)


class ImgObsWrapper(ObservationWrapper):
    """wrappers.py:156-166: the image is the only observation output."""

    def __init__(self, env):
        super().__init__(env)
        self.observation_space = env.observation_space.spaces['image']

    def observation(self, obs):
        return obs['image']


def _onehot(cells, n_classes, n_colors, n_states, class_map=None):
    L = _lib.load()
    cells = cells.contiguous()
    assert cells.dtype == torch.uint8 and cells.shape[-1] == 3 and cells.is_cuda
    nbits = n_classes + n_colors + n_states
    out = torch.empty(cells.shape[:-1] + (nbits,), dtype=torch.uint8, device=cells.device)
    cm = None
    if class_map is not None:
        cm = (C.c_uint8 * 11)(*class_map)
    with torch.cuda.device(cells.device):
        _lib.check(L.mgb_onehot(_ptr(cells), _ptr(out), cells.numel() // 3, cm, n_classes, n_colors, n_states,
                                C.c_void_p(torch.cuda.current_stream(cells.device).cuda_stream)))
    return out


class OneHotPartialObsWrapper(ObservationWrapper):
    """wrappers.py:203-243: 7x7x3 view -> 7x7x(11+7+3) one-hot planes."""

    def __init__(self, env, tile_size=8):
        super().__init__(env)
        self.tile_size = tile_size
        shape = env.observation_space['image'].shape
        self.num_bits = len(OBJECT_TO_IDX) + len(COLOR_TO_IDX) + len(STATE_TO_IDX)
        self.observation_space = spaces.Dict(dict(env.observation_space.spaces))
        self.observation_space.spaces["image"] = spaces.Box(0, 255, (shape[0], shape[1], self.num_bits), 'uint8')

    def observation(self, obs):
        out = _onehot(obs['image'], len(OBJECT_TO_IDX), len(COLOR_TO_IDX), len(STATE_TO_IDX))
        return {'mission': obs['mission'], 'image': out}


class FullyObsWrapper(ObservationWrapper):
    """wrappers.py:311-338: full grid encoding with the agent drawn as (10, 0, dir)."""

    def __init__(self, env):
        super().__init__(env)
        self.observation_space = spaces.Dict(dict(env.observation_space.spaces))
        self.observation_space.spaces["image"] = spaces.Box(0, 255, (self.env.width, self.env.height, 3), 'uint8')

    def observation(self, obs):
        return {'mission': obs['mission'], 'image': self.unwrapped.full_obs()}


class FullyObsOneHotWrapper(ObservationWrapper):
    """wrappers.py:340-415.  Expects an array observation, e.g. ImgObsWrapper(FullyObsWrapper(env))."""

    def __init__(self, env, drop_color=False, keep_classes=None, flatten=True):
        super().__init__(env)
        if not keep_classes:
            keep_classes = list(OBJECT_TO_IDX.keys())
        keep_classes.sort(key=lambda x: OBJECT_TO_IDX[x])
        self.num_classes = len(keep_classes)
        self.object_to_new_idx = {OBJECT_TO_IDX[k]: i for i, k in enumerate(keep_classes)}
        self.num_colors = 0 if drop_color else len(COLOR_TO_IDX)
        self.num_states = 4
        self.N = self.num_classes + self.num_colors + self.num_states
        o = self.env.observation_space
        try:
            shp = o['image'].shape
        except Exception:
            shp = o.shape
        self._cells_shape = tuple(shp[:2])
        self.flatten = flatten
        n = int(np.prod(self._cells_shape))
        shape = (n * self.N,) if flatten else self._cells_shape + (self.N,)
        self.obsshape = n
        self.observation_space = spaces.Box(0, 1, shape, 'uint8')
        # types that are not kept have no plane: the reference would raise KeyError; map them to an
        # out-of-range index so that no class bit is set
        self._class_map = [self.object_to_new_idx.get(t, 255) for t in range(11)]

    def observation(self, obs):
        out = _onehot(obs, self.num_classes, self.num_colors, self.num_states, self._class_map)
        n = out.shape[0]
        return out.reshape(n, -1) if self.flatten else out.reshape((n,) + self._cells_shape + (self.N,))


class FlatObsWrapper(ObservationWrapper):
    """wrappers.py:528-577: image.flatten() ++ 27x96 one-hot of the lower-cased mission, float32."""

    def __init__(self, env, maxStrLen=96):
        super().__init__(env)
        self.maxStrLen = maxStrLen
        self.numCharCodes = 27
        img = env.observation_space.spaces['image']
        self._img_bytes = int(np.prod(img.shape))
        self.observation_space = spaces.Box(0, 255, (1, self._img_bytes + self.numCharCodes * self.maxStrLen), 'uint8')
        u = env.unwrapped
        tmpl = u._mission
        # every mission string this env can produce: one for the static ones, one per target colour for
        # KeyCorridor ("pick up the <colour> ball", keycorridor.py:49; row index = colour id), and for the level-pool
        # ids whose mission names the level's target (Fetch, GoToDoor, GoToObject, PutNear ...) the distinct
        # strings of the uploaded pool (row index = _lvl2row[level the env is playing])
        self._per_colour = False
        self._lvl2row = None
        if u._pool_missions is not None:
            uniq = sorted(set(u._pool_missions))
            if len(uniq) > 256:
                raise _lib.MgbError("FlatObsWrapper: %d distinct missions in the level pool (at most 256)" % len(uniq))
            row = {m: k for k, m in enumerate(uniq)}
            self._lvl2row = torch.as_tensor(np.array([row[m] for m in u._pool_missions], np.uint8)).to(u.device)
            missions = uniq
        elif u._pool_n and "(per level)" in tmpl:
            raise _lib.MgbError("FlatObsWrapper: this level-pool env has per-level missions but none were uploaded "
                                "(set_level_pool(..., missions=[...]))")
        elif "%s" in tmpl:
            self._per_colour = True
            missions = [tmpl % (IDX_TO_COLOR[c], IDX_TO_OBJECT[6]) for c in range(len(COLOR_TO_IDX))]
        else:
            missions = [tmpl]
        table = np.stack([self._encode(m) for m in missions])
        self._table = torch.as_tensor(table).to(u.device)

    def _encode(self, mission):
        assert len(mission) <= self.maxStrLen, 'mission string too long ({} chars)'.format(len(mission))
        arr = np.zeros((self.maxStrLen, self.numCharCodes), dtype='float32')
        chNo = 0
        for idx, ch in enumerate(mission.lower()):
            if 'a' <= ch <= 'z':
                chNo = ord(ch) - ord('a')
            elif ch == ' ':
                chNo = ord('z') - ord('a') + 1
            assert chNo < self.numCharCodes
            arr[idx, chNo] = 1
        return arr.reshape(-1)

    def observation(self, obs):
        L = _lib.load()
        img = obs['image'].contiguous()
        u = self.unwrapped
        N = img.shape[0]
        midx = None
        if self._per_colour:
            midx = u.get_state(("target",))["target"][:, 1].contiguous()
        elif self._lvl2row is not None:
            midx = self._lvl2row[u.level_index_device().long()].contiguous()
        mlen = self._table.shape[1]
        out = torch.empty((N, self._img_bytes + mlen), dtype=torch.float32, device=img.device)
        with torch.cuda.device(img.device):
            _lib.check(L.mgb_flat_obs(_ptr(img), self._img_bytes, _ptr(self._table), mlen, _ptr(midx), _ptr(out), N,
                                      C.c_void_p(torch.cuda.current_stream(img.device).cuda_stream)))
        return out


class ViewSizeWrapper(Wrapper):
    """wrappers.py:579-608: customise the agent's field of view.  The reference mutates
    `env.unwrapped.agent_view_size`; here the view size is a template parameter of the kernel, so the wrapped
    env is rebuilt with the new size and its complete state (grids, agents, RNG positions) is carried over.
    Odd sizes 3, 5, 7, 9, 11.  Like the reference, not to be combined with the fully-observable wrappers."""

    def __init__(self, env, agent_view_size=7):
        base = env.unwrapped.with_view_size(agent_view_size)
        env.unwrapped.close()
        super().__init__(base)
        self.observation_space = spaces.Dict({
            'image': spaces.Box(0, 255, (agent_view_size, agent_view_size, 3), 'uint8')})


_ATLAS = {}


def _atlas(tile_size, device):
    """tile atlas rendered by the reference's own Grid.render_tile (oracle/gen_atlas.py), shipped as data"""
    import os
    key = (tile_size, str(device))
    if key not in _ATLAS:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "tile_atlas_t%d.npz" % tile_size)
        if not os.path.exists(path):
            raise _lib.MgbError("no tile atlas for tile_size=%d (shipped: 8); build one with oracle/gen_atlas.py %d "
                                "on a machine that has the reference" % (tile_size, tile_size))
        z = np.load(path)
        _ATLAS[key] = (torch.as_tensor(z["atlas"]).to(device).contiguous(), z["valid"])
    return _ATLAS[key][0]


class RGBImgPartialObsWrapper(ObservationWrapper):
    """wrappers.py:283-309 -> MiniGridEnv.get_obs_render (minigrid.py:1383-1398): the agent's view as RGB."""

    def __init__(self, env, tile_size=8):
        super().__init__(env)
        self.tile_size = tile_size
        shape = env.observation_space['image'].shape
        self.observation_space = spaces.Dict(dict(env.observation_space.spaces))
        self.observation_space.spaces['image'] = spaces.Box(0, 255, (shape[0] * tile_size, shape[1] * tile_size, 3), 'uint8')
        self._atlas = _atlas(tile_size, env.unwrapped.device)

    def observation(self, obs):
        L = _lib.load()
        img = obs['image'].contiguous()
        N, V = img.shape[0], img.shape[1]
        ts = self.tile_size
        out = torch.empty((N, V * ts, V * ts, 3), dtype=torch.uint8, device=img.device)
        with torch.cuda.device(img.device):
            _lib.check(L.mgb_render_partial(_ptr(img), V, _ptr(self._atlas), ts, _ptr(out), N,
                                            C.c_void_p(torch.cuda.current_stream(img.device).cuda_stream)))
        return {'mission': obs['mission'], 'image': out}


class RGBImgObsWrapper(ObservationWrapper):
    """wrappers.py:245-281 -> env.render('rgb_array', highlight=False): the whole grid as RGB.
    The declared space keeps the reference's (width*ts, height*ts, 3); the array is [N, height*ts, width*ts, 3]
    like the reference's render output."""

    def __init__(self, env, tile_size=8):
        super().__init__(env)
        self.tile_size = tile_size
        self.observation_space = spaces.Dict(dict(env.observation_space.spaces))
        self.observation_space.spaces['image'] = spaces.Box(0, 255, (self.env.width * tile_size, self.env.height * tile_size, 3), 'uint8')
        self._atlas = _atlas(tile_size, env.unwrapped.device)

    def observation(self, obs):
        return {'mission': obs['mission'], 'image': self.unwrapped.render('rgb_array', highlight=False, tile_size=self.tile_size)}


# ------------------------------------------------------------------------------------------------
# bookkeeping wrappers
# ------------------------------------------------------------------------------------------------
def _stream(env):
    return env.unwrapped._stream()


class _TakesOverAutoReset(Wrapper):
    """The reference wrappers below read the env *after* its step and *before* any reset (terminal agent position).
    The batched env auto-resets inside the step kernel, so these wrappers turn that off on the wrapped env and, when
    `autoreset` is True (default: what the wrapped env was doing), reset the finished envs themselves right after the
    bookkeeping -- the caller still sees the gym.vector convention (obs at a done step = first obs of the next episode)."""

    def __init__(self, env, autoreset=None):
        super().__init__(env)
        u = self.unwrapped
        if env is not u:
            raise _lib.MgbError("%s must wrap the batched env directly (observation wrappers go on top of it): it resets "
                                "finished envs after its bookkeeping, which an inner observation wrapper would not see"
                                % type(self).__name__)
        self._autoreset = u.autoreset if autoreset is None else bool(autoreset)
        u.set_autoreset(False)

    def rollout(self, *a, **kw):
        raise NotImplementedError("%s does its bookkeeping between steps: call step(); rollout() of the wrapped env would "
                                  "skip it (and no longer auto-resets)" % type(self).__name__)

    def _reset_done(self, obs, done):
        if not self._autoreset:
            return obs
        u = self.unwrapped
        # in place: the step's own (fresh) observation tensors receive the first observation of the new episodes
        _lib.check(u._L.mgb_reset(u._h, _ptr(done.view(torch.uint8)), _ptr(obs['image']), _ptr(obs['direction']), u._stream()))
        return obs


class TerminalObservation(_TakesOverAutoReset):
    """Not a reference class: the gym.vector companion of auto-reset.  The batched env returns, at a done step, the
    first observation of the next episode; this wrapper also hands out the observation the reference
    would have returned at that step: info['terminal_observation'] = {'image' uint8 [N,V,V,3], 'direction' [N]},
    meaningful where `done` is set (elsewhere it equals the returned observation)."""

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        info = dict(info)
        info['terminal_observation'] = {'image': obs['image'].clone(), 'direction': obs['direction'].clone()}
        return self._reset_done(obs, done), reward, done, info


class _VisitBonus(_TakesOverAutoReset):
    _BY_ACTION = 0

    def __init__(self, env, autoreset=None):
        super().__init__(env, autoreset)
        u = self.unwrapped
        self._table = u.width * u.height * (4 * u.action_space.n if self._BY_ACTION else 1)
        # the reference's `self.counts` dict, one per env instance, never cleared by reset (wrappers.py:94-96,118-119)
        self.counts = torch.zeros((u.num_envs, self._table), dtype=torch.int32, device=u.device)

    def step(self, action):
        u = self.unwrapped
        a = u._actions(action, (u.num_envs,))
        obs, reward, done, info = self.env.step(a)
        _lib.check(u._L.mgb_visit_bonus(u._h, self._BY_ACTION, _ptr(a), _ptr(self.counts), self._table, _ptr(reward), u._stream()))
        return self._reset_done(obs, done), reward, done, info


class ActionBonus(_VisitBonus):
    """wrappers.py:87-119: reward += 1/sqrt(number of times this env saw (agent_pos, agent_dir, action))."""
    _BY_ACTION = 1


class StateBonus(_VisitBonus):
    """wrappers.py:121-154: reward += 1/sqrt(number of times this env's agent stood on agent_pos)."""
    _BY_ACTION = 0


class DACWrapper(_TakesOverAutoReset):
    """wrappers.py:35-84: once an env's episode ends it returns a blank observation (image*0+1, the reset-time
    direction), reward 0, and done only when `count >= max_steps` -- every episode lasts exactly max_steps, so the
    whole batch finishes on the same step and the caller resets it, as with the reference.  A finished env is still
    stepped by the kernel (its outputs are discarded; its state is regenerated by the next reset)."""

    def __init__(self, env):
        super().__init__(env, autoreset=False)
        u = self.unwrapped
        N, dev = u.num_envs, u.device
        self._envdone = [torch.zeros(N, dtype=torch.uint8, device=dev) for _ in range(2)]
        self._reset_dir = torch.zeros(N, dtype=torch.uint8, device=dev)
        self._done_out = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.count = 0

    @property
    def env_done(self):
        return self._envdone[0].view(torch.bool)

    def reset(self, **kw):
        obs = self.env.reset(**kw)
        self._envdone[0].zero_()
        self._reset_dir.copy_(obs['direction'])
        self.count = 0
        return obs

    def step(self, action):
        self.count += 1
        obs, reward, done, info = self.env.step(action)
        u = self.unwrapped
        _lib.check(u._L.mgb_dac(u._h, int(self.count >= u.max_steps), _ptr(done.view(torch.uint8)), _ptr(self._envdone[0]),
                                _ptr(self._envdone[1]), _ptr(self._reset_dir), _ptr(obs['image']), _ptr(reward),
                                _ptr(self._done_out), _ptr(obs['direction']), u._stream()))
        self._envdone.reverse()
        return obs, reward, self._done_out.view(torch.bool), info


class AgentExtraInfoWrapper(ObservationWrapper):
    """wrappers.py:169-200: adds 'pos' int32 [N,2] and 'dir' int32 [N] to the observation dict."""

    def __init__(self, env):
        super().__init__(env)
        self.observation_space = spaces.Dict({
            'image': env.observation_space.spaces['image'],
            'pos': spaces.Box(-1, 10000, (2,), 'float32'),
            'dir': spaces.Box(0, 5, (), 'float32')})

    def observation(self, obs):
        agent = self.unwrapped.get_state(("agent",))["agent"]
        out = {'pos': agent[:, :2], 'dir': agent[:, 2]}
        out.update(obs)
        return out

    def get_map(self):
        """type plane of grid.encode() (wrappers.py:188-191): uint8 [N,W,H]."""
        return self.unwrapped.get_state(("grid",))["grid"][..., 0]

    def get_full_map(self):
        """grid.encode() with the agent cell := (10, red, dir) (wrappers.py:193-201) -- the FullyObsWrapper image."""
        return self.unwrapped.full_obs()


class AppendActionWrapper(Wrapper):
    """wrappers.py:418-458: appends the one-hot encodings of the last K actions to a flat uint8 observation
    (e.g. FullyObsOneHotWrapper(..., flatten=True)).  An env that finished starts over with an empty history."""

    def __init__(self, env, K):
        super().__init__(env)
        self.K = int(K)
        self.actsize = env.action_space.n
        self._D = int(self.env.observation_space.shape[0])
        u = self.unwrapped
        self._hist = torch.full((u.num_envs, self.K), 255, dtype=torch.uint8, device=u.device)
        self.observation_space = spaces.Box(0, 1, (self._D + self.actsize * self.K,), 'uint8')

    def _append(self, obs, actions, done):
        u = self.unwrapped
        obs = obs.contiguous()
        assert obs.dtype == torch.uint8 and tuple(obs.shape) == (u.num_envs, self._D)
        out = torch.empty((u.num_envs, self._D + self.actsize * self.K), dtype=torch.uint8, device=u.device)
        _lib.check(u._L.mgb_append_action(u.num_envs, self._D, self.actsize, self.K, _ptr(obs), _ptr(actions),
                                          _ptr(None if done is None else done.view(torch.uint8)), _ptr(self._hist), _ptr(out), u._stream()))
        return out

    def reset(self, **kw):
        return self._append(self.env.reset(**kw), None, None)

    def step(self, action):
        u = self.unwrapped
        a = u._actions(action, (u.num_envs,))
        obs, reward, done, info = self.env.step(a)
        return self._append(obs, a, done), reward, done, info


class GoalPolicyWrapper(Wrapper):
    """wrappers.py:461-526 (a gym GoalEnv over FullyObsOneHotWrapper): observation dict with 'achieved_goal' (goal
    plane erased) and 'desired_goal' (agent moved onto the goal cell)."""

    def __init__(self, env):
        assert isinstance(env, FullyObsOneHotWrapper)
        super().__init__(env)
        self.observation_space = spaces.Dict({
            'observation': env.observation_space, 'achieved_goal': env.observation_space, 'desired_goal': env.observation_space})

    def _get_goals(self, obs):
        e, u = self.env, self.unwrapped
        obs = obs.contiguous()
        achieved, desired = torch.empty_like(obs), torch.empty_like(obs)
        _lib.check(u._L.mgb_goal_policy(u.num_envs * e.obsshape, e.N, e.object_to_new_idx[OBJECT_TO_IDX['agent']],
                                        e.object_to_new_idx[OBJECT_TO_IDX['empty']], e.object_to_new_idx[OBJECT_TO_IDX['goal']],
                                        _ptr(obs), _ptr(achieved), _ptr(desired), u._stream()))
        return achieved, desired

    def compute_reward(self, achieved_goal=None, desired_goal=None, info=None):
        return self.unwrapped._reward()                     # wrappers.py:499-505: the first `_reward` down the chain

    def _pack(self, obs):
        achieved, desired = self._get_goals(obs)
        return {'observation': obs, 'achieved_goal': achieved, 'desired_goal': desired}

    def reset(self, **kw):
        return self._pack(self.env.reset(**kw))

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        return self._pack(obs), reward, done, info


class EpisodeStatistics(Wrapper):
    """Not a reference class (the reference only prints; SURVEY §5 "metrics"): running return and length per env on the
    device.  After every step info['episode'] = {'r': float64 [N], 'l': int32 [N]} holds the totals of the episodes that
    just ended (0 elsewhere, use `done` as the mask); `episodes` / `mean_length()` give batch totals without a host loop."""

    def __init__(self, env):
        super().__init__(env)
        u = self.unwrapped
        N, dev = u.num_envs, u.device
        self._ret = torch.zeros(N, dtype=torch.float64, device=dev)
        self._len = torch.zeros(N, dtype=torch.int32, device=dev)
        self._out_ret = torch.zeros(N, dtype=torch.float64, device=dev)
        self._out_len = torch.zeros(N, dtype=torch.int32, device=dev)
        self._totals = torch.zeros(2, dtype=torch.int64, device=dev)

    def reset(self, **kw):
        self._ret.zero_(); self._len.zero_()
        return self.env.reset(**kw)

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        u = self.unwrapped
        _lib.check(u._L.mgb_episode_stats(u.num_envs, _ptr(reward), _ptr(done.view(torch.uint8)), _ptr(self._ret), _ptr(self._len),
                                          _ptr(self._out_ret), _ptr(self._out_len), _ptr(self._totals), u._stream()))
        info = dict(info)
        info['episode'] = {'r': self._out_ret, 'l': self._out_len}
        return obs, reward, done, info

    @property
    def episodes(self):
        return int(self._totals[0])

    def mean_length(self):
        t = self._totals.tolist()
        return t[1] / t[0] if t[0] else float('nan')
