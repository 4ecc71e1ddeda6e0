"""VecMiniGridEnv: the batched, GPU-resident mirror of the reference's MiniGridEnv interface
(gym_minigrid/minigrid.py:720-1381).  Same attribute and method names, batched:

    env = make('MiniGrid-Empty-8x8-v0', num_envs=1 << 20)
    obs = env.reset()                              # {'image': uint8[N,7,7,3] (cuda), 'direction': uint8[N], 'mission': seq}
    obs, reward, done, info = env.step(actions)    # reward float64[N], done bool[N]; auto-reset on done

PyTorch is plumbing only (device memory, streams); all computation happens in libmgb200.so.
"""
from enum import IntEnum

import ctypes as C
import numpy as np
import torch

from . import _lib, spaces

# minigrid.py:27-61
COLOR_TO_IDX = {'red': 0, 'green': 1, 'blue': 2, 'purple': 3, 'yellow': 4, 'grey': 5, 'white': 6}
IDX_TO_COLOR = {v: k for k, v in COLOR_TO_IDX.items()}
OBJECT_TO_IDX = {'unseen': 0, 'empty': 1, 'wall': 2, 'floor': 3, 'door': 4, 'key': 5, 'ball': 6, 'box': 7,
                 'goal': 8, 'lava': 9, 'agent': 10}
IDX_TO_OBJECT = {v: k for k, v in OBJECT_TO_IDX.items()}
STATE_TO_IDX = {'open': 0, 'closed': 1, 'locked': 2}


class Actions(IntEnum):          # MiniGridEnv.Actions, minigrid.py:731-745
    left = 0
    right = 1
    forward = 2
    pickup = 3
    drop = 4
    toggle = 5
    done = 6


class MissionBatch:
    """obs['mission'] for N envs without materialising N Python strings per step."""

    def __init__(self, env):
        self._env = env
        self._targets = None

    def __len__(self):
        return self._env.num_envs

    def __getitem__(self, i):
        tmpl = self._env._mission
        if self._env._pool_missions is not None:        # level-pool env with per-level mission strings
            return self._env._pool_missions[int(self._env.level_index(i, 1)[0])]
        if "%s" not in tmpl:
            return tmpl
        if self._targets is None:          # KeyCorridor: "pick up the <colour> <type>" (keycorridor.py:49)
            self._targets = self._env.get_state(("target",))["target"].cpu().numpy()
        t, c = self._targets[i]
        return tmpl % (IDX_TO_COLOR[int(c)], IDX_TO_OBJECT[int(t)])

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _philox4x32_10(ctr, key):
    """host copy of the device RNG (Salmon et al., SC'11); only used to map envs to pool levels lazily"""
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    M = 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M, p1 & M, ((p0 >> 32) ^ c3 ^ k1) & M, p0 & M
        k0, k1 = (k0 + 0x9E3779B9) & M, (k1 + 0xBB67AE85) & M
    return c0, c1, c2, c3


class BatchedGrid:
    """`env.grid` of a batched env: nothing is copied until a method is called."""

    def __init__(self, env):
        self._env = env
        self.width, self.height = env.width, env.height

    def encode(self, vis_mask=None):
        """Grid.encode (minigrid.py:571-594) for every env: uint8 [N, width, height, 3]; vis_mask: bool [N, W, H] or [W, H]"""
        g = self._env.get_state(("grid",))["grid"]
        if vis_mask is not None:
            m = torch.as_tensor(vis_mask, device=g.device).to(torch.bool)
            g = g * m.expand(g.shape[:3]).unsqueeze(-1).to(g.dtype)
        return g

    def get(self, i, j):
        """Grid.get (minigrid.py:417-420) as encodings: uint8 [N, 3] = (type, colour, state) of cell (i, j); empty = (1, 0, 0)"""
        assert 0 <= i < self.width and 0 <= j < self.height
        return self.encode()[:, i, j]

    def __contains__(self, key):
        """(colour, type) in grid (minigrid.py:377-391) -- true if ANY env of the batch holds such an object"""
        c, t = key
        g = self.encode()
        m = torch.ones(g.shape[:3], dtype=torch.bool, device=g.device)
        if t is not None:
            m &= g[..., 0] == OBJECT_TO_IDX[t]
        if c is not None:
            m &= g[..., 1] == COLOR_TO_IDX[c]
        return bool(m.any())


class VecMiniGridEnv:
    metadata = {'render.modes': ['rgb_array'], 'video.frames_per_second': 10}    # minigrid.py:725-728; no window ('human')
    Actions = Actions

    def __init__(self, spec, num_envs=1, device=None, seed=1337, env_id_base=0, autoreset=True, agent_view_size=7):
        self.spec = spec
        cfg = dict(spec["config"])
        cfg["agent_view_size"] = int(agent_view_size)
        self._ctor = dict(num_envs=num_envs, device=device, seed=seed, env_id_base=env_id_base, autoreset=autoreset)
        self._cfg = cfg
        self._L = _lib.load()                      # raises loudly when libmgb200.so is missing
        if not torch.cuda.is_available():
            raise _lib.MgbError("gym_minigrid_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise _lib.MgbError("device must be a CUDA device, got %s" % self.device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_envs = int(num_envs)
        # reference attribute names (minigrid.py:785-819)
        self.actions = Actions
        self.action_space = spaces.Discrete(cfg["n_actions"])
        V = self.agent_view_size = int(agent_view_size)          # minigrid.py:776,795
        self.observation_space = spaces.Dict({'image': spaces.Box(0, 255, (V, V, 3), 'uint8')})
        self.reward_range = cfg["reward_range"]
        self.width, self.height = cfg["width"], cfg["height"]
        self.max_steps = cfg["max_steps"]
        self.see_through_walls = bool(cfg["see_through"])
        self._mission = cfg["mission"]
        c = _lib.MgbConfig(**{k: int(cfg[k]) for k, _ in _lib.MgbConfig._fields_})
        h = C.c_void_p()
        _lib.check(self._L.mgb_create(C.byref(c), self.num_envs, self.device.index, C.c_uint64(int(seed) & (2 ** 64 - 1)),
                                      int(env_id_base), C.byref(h)))
        self._h = h
        self._seed = seed
        self._env_id_base = int(env_id_base)
        self._pool_n = 0
        self._pool_missions = None
        self.autoreset = True
        if not autoreset:
            self.set_autoreset(False)
        self._tape = None
        self._host = None
        self._render_scratch = None

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _actions(self, actions, shape):
        a = actions
        if not torch.is_tensor(a):
            a = torch.as_tensor(np.asarray(a))
        if a.dtype != torch.uint8:
            a = a.to(torch.uint8)
        a = a.to(self.device, non_blocking=True).contiguous()
        if tuple(a.shape) != tuple(shape):
            raise ValueError("actions must have shape %s, got %s" % (tuple(shape), tuple(a.shape)))
        return a

    def _new_out(self):
        """fresh (obs, reward, done, dir) tensors: like the reference, reset() / step() hand out new arrays on every
        call, so an `obs` kept by the caller is never overwritten by a later step.  Allocation only (the kernel writes
        straight into them); pass `out=` to step() / rollout() to reuse buffers instead."""
        N, V, dev = self.num_envs, self.agent_view_size, self.device
        return (torch.empty((N, V, V, 3), dtype=torch.uint8, device=dev), torch.empty((N,), dtype=torch.float64, device=dev),
                torch.empty((N,), dtype=torch.uint8, device=dev), torch.empty((N,), dtype=torch.uint8, device=dev))

    def _obs_dict(self, image, direction):
        return {'image': image, 'direction': direction, 'mission': MissionBatch(self)}

    def with_view_size(self, agent_view_size):
        """ViewSizeWrapper support (wrappers.py:579-608): the view size is a compile-time parameter of the
        kernel, so a new handle is created and the complete env state is carried over."""
        if self._pool_n:
            raise _lib.MgbError("change the view size before uploading the level pool (make(..., agent_view_size=V))")
        new = VecMiniGridEnv(self.spec, agent_view_size=agent_view_size, **self._ctor)
        new.set_state({k: v for k, v in self.get_state().items()})
        return new

    def close(self):
        if getattr(self, "_h", None):
            self._L.mgb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def unwrapped(self):
        return self

    def set_autoreset(self, on):
        """Auto-reset on done (gym.vector convention) on/off.  Off = the reference's behaviour: a finished env keeps
        its terminal state until reset(mask) -- what the bookkeeping wrappers (ActionBonus, DACWrapper ...) build on."""
        _lib.check(self._L.mgb_set_autoreset(self._h, int(bool(on))))
        self.autoreset = bool(on)

    def _reward(self):
        """MiniGridEnv._reward (minigrid.py:933-937) for every env at its current step_count: float64 [N]."""
        steps = self.step_count.to(torch.float64)
        # three separately rounded fp64 operations, as in Python.  The divisor is a tensor on purpose: torch divides a
        # CUDA tensor by a Python scalar as a multiplication by its reciprocal, which is 1 ulp off for some step counts.
        return 1 - 0.9 * (steps / torch.full_like(steps, float(self.max_steps)))

    # ------------------------------------------------------------------ gym protocol
    def seed(self, seed=1337):
        """MiniGridEnv.seed (minigrid.py:860-863): re-key the RNG; effective at the next reset."""
        _lib.check(self._L.mgb_seed(self._h, C.c_uint64(int(seed) & (2 ** 64 - 1))))
        self._seed = seed
        return [seed]

    def reset(self, mask=None):
        """MiniGridEnv.reset for every env (or those with mask != 0); returns the batched obs dict."""
        m = None
        if mask is not None:
            m = torch.as_tensor(mask).to(self.device).to(torch.uint8).contiguous()
            assert m.shape == (self.num_envs,)
        obs, _, _, d = self._new_out()
        _lib.check(self._L.mgb_reset(self._h, _ptr(m), _ptr(obs), _ptr(d), self._stream()))
        return self._obs_dict(obs, d)

    def step(self, actions, out=None):
        """MiniGridEnv.step for every env.  Returns fresh tensors on every call (as the reference returns fresh arrays);
        `out` = optional (obs, reward, done, dir) tensors to write into instead (zero-allocation loop: the caller then
        owns the aliasing)."""
        a = self._actions(actions, (self.num_envs,))
        obs, reward, done, d = out if out is not None else self._new_out()
        _lib.check(self._L.mgb_step(self._h, _ptr(a), _ptr(obs), _ptr(reward), _ptr(done), _ptr(d), self._stream()))
        return self._obs_dict(obs, d), reward, done.view(torch.bool), {}

    def rollout(self, actions, out=None, want_obs=True):
        """T steps in one persistent kernel.  actions [T,N] -> obs [T,N,7,7,3], reward, done, dir [T,N]."""
        T = int(actions.shape[0])
        a = self._actions(actions, (T, self.num_envs))
        N = self.num_envs
        if out is None:
            V = self.agent_view_size
            obs = torch.empty((T, N, V, V, 3), dtype=torch.uint8, device=self.device) if want_obs else None
            reward = torch.empty((T, N), dtype=torch.float64, device=self.device)
            done = torch.empty((T, N), dtype=torch.uint8, device=self.device)
            d = torch.empty((T, N), dtype=torch.uint8, device=self.device)
        else:
            obs, reward, done, d = out
        _lib.check(self._L.mgb_rollout(self._h, T, _ptr(a), _ptr(obs), _ptr(reward), _ptr(done), _ptr(d), self._stream()))
        return obs, reward, done.view(torch.bool), d

    def rollout_random(self, T, want_obs=True):
        """T steps under the uniform random policy (the reference's `env.action_space.sample()` loop), drawn on the device
        from a counter-based stream: no action input.  Returns (obs, reward, done, dir, actions) with actions uint8 [T,N]."""
        T, N, V = int(T), self.num_envs, self.agent_view_size
        obs = torch.empty((T, N, V, V, 3), dtype=torch.uint8, device=self.device) if want_obs else None
        reward = torch.empty((T, N), dtype=torch.float64, device=self.device)
        done = torch.empty((T, N), dtype=torch.uint8, device=self.device)
        d = torch.empty((T, N), dtype=torch.uint8, device=self.device)
        a = torch.empty((T, N), dtype=torch.uint8, device=self.device)
        _lib.check(self._L.mgb_rollout_random(self._h, T, _ptr(a), _ptr(obs), _ptr(reward), _ptr(done), _ptr(d), self._stream()))
        return obs, reward, done.view(torch.bool), d, a

    def step_host(self, actions_host):
        """End-to-end step with HOST buffers: actions (pinned uint8[N]) in, pinned numpy-viewable
        obs/reward/done/dir out; H2D + kernel + D2H are pipelined inside libmgb200 (mgb_step_host)."""
        N = self.num_envs
        if self._host is None:
            self._host = dict(
                obs=torch.empty((N, self.agent_view_size, self.agent_view_size, 3), dtype=torch.uint8).pin_memory(),
                reward=torch.empty((N,), dtype=torch.float64).pin_memory(),
                done=torch.empty((N,), dtype=torch.uint8).pin_memory(),
                dir=torch.empty((N,), dtype=torch.uint8).pin_memory())
        a = actions_host
        if not torch.is_tensor(a):
            a = torch.as_tensor(np.asarray(a))
        a = a.to(torch.uint8).contiguous()
        assert a.device.type == "cpu" and a.shape == (N,)
        hb = self._host
        _lib.check(self._L.mgb_step_host(self._h, _ptr(a), _ptr(hb["obs"]), _ptr(hb["reward"]), _ptr(hb["done"]), _ptr(hb["dir"])))
        return self._obs_dict(hb["obs"], hb["dir"]), hb["reward"], hb["done"].view(torch.bool), {}

    # ------------------------------------------------------------------ state / checkpoint
    _STATE_FIELDS = ("grid", "aux", "agent", "carrying", "obstacles", "target", "rng")

    def _state_shapes(self, n):
        W, H = self.width, self.height
        return dict(grid=((n, W, H, 3), torch.uint8), aux=((n, W, H), torch.uint8), agent=((n, 4), torch.int32),
                    carrying=((n, 3), torch.uint8), obstacles=((n, _lib.MAX_OBSTACLES, 2), torch.int16),
                    target=((n, 2), torch.uint8), rng=((n, 2), torch.int32))

    def get_state(self, fields=None, first=0, count=None):
        """Snapshot in the reference's encoding (Grid.encode layout); also the checkpoint format."""
        count = self.num_envs - first if count is None else count
        fields = self._STATE_FIELDS if fields is None else fields
        shapes = self._state_shapes(count)
        out = {k: torch.zeros(shapes[k][0], dtype=shapes[k][1], device=self.device) for k in fields}
        args = [_ptr(out.get(k)) for k in self._STATE_FIELDS]
        _lib.check(self._L.mgb_get_state(self._h, first, count, *args, self._stream()))
        return out

    def set_state(self, state, first=0):
        shapes = self._state_shapes(0)
        count = None
        ts = {}
        for k in self._STATE_FIELDS:
            if k in state and state[k] is not None:
                v = state[k]
                if not torch.is_tensor(v):
                    v = torch.as_tensor(np.ascontiguousarray(v))
                if k == "rng":
                    v = v.to(torch.int64).to(torch.int32) if v.dtype not in (torch.int32,) else v
                v = v.to(self.device).to(shapes[k][1]).contiguous()
                n = v.shape[0]
                if k == "obstacles" and v.shape[1] != _lib.MAX_OBSTACLES:
                    pad = torch.zeros((n, _lib.MAX_OBSTACLES, 2), dtype=torch.int16, device=self.device)
                    pad[:, :v.shape[1]] = v
                    v = pad
                assert tuple(v.shape[1:]) == tuple(self._state_shapes(n)[k][0][1:]), (k, v.shape)
                count = n if count is None else count
                assert n == count
                ts[k] = v
        if count is None:
            return
        args = [_ptr(ts.get(k)) for k in self._STATE_FIELDS]
        _lib.check(self._L.mgb_set_state(self._h, first, count, *args, self._stream()))
        torch.cuda.current_stream(self.device).synchronize()     # ts goes out of scope
        self.check_errors()

    def set_level_pool(self, grid, agent, aux=None, missions=None, hook_params=None):
        """Level-pool mode: upload K reference-generated layouts; reset / auto-reset of
        env e in episode k picks level mulhi32(philox(seed, e, k).word0, K)."""
        g = torch.as_tensor(np.ascontiguousarray(grid, dtype=np.uint8)).to(self.device)
        K = g.shape[0]
        assert tuple(g.shape[1:]) == (self.width, self.height, 3), g.shape
        a = torch.as_tensor(np.ascontiguousarray(np.asarray(agent)[:, :3], dtype=np.int32)).to(self.device)
        x = None if aux is None else torch.as_tensor(np.ascontiguousarray(aux, dtype=np.uint8)).to(self.device)
        hp = None if hook_params is None else torch.as_tensor(np.ascontiguousarray(hook_params, dtype=np.int32).reshape(K, 16)).to(self.device)
        _lib.check(self._L.mgb_set_level_pool(self._h, K, _ptr(g), _ptr(x), _ptr(a), _ptr(hp), self._stream()))
        self._pool_n = K
        self._pool_missions = list(missions) if missions is not None else None
        self.check_errors()

    def level_index_device(self):
        """which pool level each env is currently playing, int32 [N] on the device"""
        lv = torch.empty((self.num_envs,), dtype=torch.int32, device=self.device)
        _lib.check(self._L.mgb_get_levels(self._h, _ptr(lv), self._stream()))
        return lv

    def level_index(self, first=0, count=None):
        """which pool level each env is currently playing"""
        lv = torch.empty((self.num_envs,), dtype=torch.int32, device=self.device)
        _lib.check(self._L.mgb_get_levels(self._h, _ptr(lv), self._stream()))
        count = self.num_envs - first if count is None else count
        return lv[first:first + count].cpu().numpy().astype(np.int64)

    def set_levels(self, levels):
        """restore which level each env is playing (checkpoint restore, after set_state)"""
        lv = torch.as_tensor(np.ascontiguousarray(levels, dtype=np.int32)).to(self.device)
        assert lv.shape == (self.num_envs,)
        _lib.check(self._L.mgb_set_levels(self._h, _ptr(lv), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()

    def set_rng_tape(self, draws, offsets):
        """RNG-tape parity mode (SURVEY §8c mode 2)."""
        if draws is None:
            self._tape = None
            _lib.check(self._L.mgb_set_rng_tape(self._h, None, None))
            return
        draws = np.ascontiguousarray(draws, dtype=np.int32)
        if draws.size == 0:
            draws = np.zeros(1, np.int32)          # Empty-8x8 draws nothing; keep the pointer non-NULL
        d = torch.as_tensor(draws).to(self.device)
        o = torch.as_tensor(np.ascontiguousarray(offsets, dtype=np.int64)).to(self.device)
        assert o.shape == (self.num_envs + 1,)
        self._tape = (d, o)
        _lib.check(self._L.mgb_set_rng_tape(self._h, _ptr(d), _ptr(o)))

    def full_obs(self):
        """FullyObsWrapper.observation (wrappers.py:311-338) for every env: uint8[N,W,H,3]."""
        out = torch.empty((self.num_envs, self.width, self.height, 3), dtype=torch.uint8, device=self.device)
        _lib.check(self._L.mgb_full_obs(self._h, _ptr(out), self._stream()))
        return out

    def render(self, mode='rgb_array', close=False, highlight=True, tile_size=8):
        """MiniGridEnv.render (minigrid.py:1400-1466), mode 'rgb_array': uint8 [N, height*ts, width*ts, 3], the cells
        the agent currently sees highlighted (read off the env's current partial observation).  Only tile sizes with a
        shipped atlas (8); the reference default is 32.  There is no window ('human' mode)."""
        if close:
            return None
        if mode != 'rgb_array':
            raise NotImplementedError("batched envs render to arrays only (mode='rgb_array')")
        from .wrappers import _atlas
        atlas = _atlas(tile_size, self.device)
        out = torch.empty((self.num_envs, self.height * tile_size, self.width * tile_size, 3), dtype=torch.uint8, device=self.device)
        view = None
        if highlight:                                         # observe the current state (no reset) into a private buffer
            if self._render_scratch is None:
                V = self.agent_view_size
                self._render_scratch = (torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device),
                                        torch.empty((self.num_envs, V, V, 3), dtype=torch.uint8, device=self.device),
                                        torch.empty(self.num_envs, dtype=torch.uint8, device=self.device))
            none, view, vdir = self._render_scratch
            _lib.check(self._L.mgb_reset(self._h, _ptr(none), _ptr(view), _ptr(vdir), self._stream()))
        _lib.check(self._L.mgb_render_full(self._h, _ptr(view), _ptr(atlas), tile_size, _ptr(out), self._stream()))
        return out

    def check_errors(self):
        """Synchronise and raise if the device flagged anything the reference would have asserted on."""
        f = C.c_uint32(0)
        _lib.check(self._L.mgb_error_flags(self._h, self._stream(), C.byref(f)))
        if f.value:
            msgs = [m for b, m in _lib.ERROR_BITS.items() if f.value & b]
            raise _lib.MgbError("device error flags 0x%x: %s" % (f.value, "; ".join(msgs)))

    # ------------------------------------------------------------------ reference attributes, batched
    @property
    def agent_pos(self):
        return self.get_state(("agent",))["agent"][:, :2]

    @property
    def agent_dir(self):
        return self.get_state(("agent",))["agent"][:, 2]

    @property
    def step_count(self):
        return self.get_state(("agent",))["agent"][:, 3]

    @property
    def carrying(self):
        return self.get_state(("carrying",))["carrying"]

    @property
    def grid(self):
        """env.grid, batched (minigrid.py:366-615): a lazy view of the device state with the Grid methods that make sense
        for a batch -- width, height, encode(), get(i, j)."""
        return BatchedGrid(self)

    # ------------------------------------------------------------------ MiniGridEnv geometry helpers, batched
    # (minigrid.py:1092-1225).  Scalars broadcast over the batch; "None" results become a boolean mask.
    def _agent(self):
        return self.get_state(("agent",))["agent"]

    @staticmethod
    def _vec(d):
        """DIR_TO_VEC (minigrid.py:63-73): right, down, left, up"""
        return torch.stack([(1 - d) * (1 - (d & 1)), (2 - d) * (d & 1)], dim=1)

    @property
    def dir_vec(self):
        return self._vec(self._agent()[:, 2])

    @property
    def right_vec(self):
        v = self.dir_vec
        return torch.stack([-v[:, 1], v[:, 0]], dim=1)

    @property
    def front_pos(self):
        a = self._agent()
        return a[:, :2] + self._vec(a[:, 2])

    @property
    def left_pos(self):
        a = self._agent()
        return a[:, :2] + self._vec((a[:, 2] - 1) % 4)

    @property
    def right_pos(self):
        a = self._agent()
        return a[:, :2] + self._vec((a[:, 2] + 1) % 4)

    def get_view_coords(self, i, j, agent=None):
        """absolute (i, j) -> the agent's view coordinates (vx, vy), possibly outside the view (minigrid.py:1135-1160)"""
        a = self._agent() if agent is None else agent
        d = self._vec(a[:, 2])
        rx, ry = -d[:, 1], d[:, 0]
        sz, hs = self.agent_view_size, self.agent_view_size // 2
        tx = a[:, 0] + d[:, 0] * (sz - 1) - rx * hs
        ty = a[:, 1] + d[:, 1] * (sz - 1) - ry * hs
        lx, ly = torch.as_tensor(i, device=self.device) - tx, torch.as_tensor(j, device=self.device) - ty
        return rx * lx + ry * ly, -(d[:, 0] * lx + d[:, 1] * ly)

    def get_view_exts(self):
        """(topX, topY, botX, botY) of the agent's view window, int32 [N,4] (minigrid.py:1162-1189)"""
        a = self._agent()
        V, h, d = self.agent_view_size, self.agent_view_size // 2, a[:, 2]
        top_x = a[:, 0] - torch.where(d == 0, 0, torch.where(d == 2, V - 1, h))
        top_y = a[:, 1] - torch.where(d == 1, 0, torch.where(d == 3, V - 1, h))
        return torch.stack([top_x, top_y, top_x + V, top_y + V], dim=1)

    def relative_coords(self, x, y, agent=None):
        """(vx, vy, valid): valid is False where the reference returns None (minigrid.py:1191-1201)"""
        vx, vy = self.get_view_coords(x, y, agent)
        V = self.agent_view_size
        return vx, vy, (vx >= 0) & (vy >= 0) & (vx < V) & (vy < V)

    def in_view(self, x, y):
        return self.relative_coords(x, y)[2]

    def gen_obs(self):
        """MiniGridEnv.gen_obs (minigrid.py:1359-1381): the observation of the current state, without stepping"""
        return self.reset(mask=torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device))

    def agent_sees(self, x, y):
        """minigrid.py:1210-1225: (x, y) is in view, shows a non-empty cell there, and its type equals the world cell's.
        (Where the reference would fail on an empty world cell under a carried object, this returns False.)"""
        s = self.get_state(("agent", "grid"))
        vx, vy, ok = self.relative_coords(x, y, s["agent"])
        img = self.gen_obs()['image']
        n = torch.arange(self.num_envs, device=self.device)
        V = self.agent_view_size
        seen = img[n, vx.clamp(0, V - 1).long(), vy.clamp(0, V - 1).long(), 0]
        xs = torch.as_tensor(x, device=self.device).expand(self.num_envs).clamp(0, self.width - 1).long()
        ys = torch.as_tensor(y, device=self.device).expand(self.num_envs).clamp(0, self.height - 1).long()
        world = s["grid"][n, xs, ys, 0]
        return ok & (seen > 1) & (seen == world)

    @property
    def mission(self):
        return MissionBatch(self)

    @property
    def kernel_launches(self):
        return int(self._L.mgb_kernel_launches(self._h))

    def set_kernel_timing(self, on=True):
        _lib.check(self._L.mgb_set_kernel_timing(self._h, int(on)))

    def last_kernel_ms(self):
        return float(self._L.mgb_last_kernel_ms(self._h))
