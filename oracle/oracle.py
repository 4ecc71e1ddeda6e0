"""TEST INFRASTRUCTURE ONLY: ctypes binding of the C oracle (oracle/minigrid_oracle.c).

Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs.  The product package never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libminigrid_oracle.so")

GEN_EMPTY, GEN_DOORKEY, GEN_FOURROOMS, GEN_DYNOBS, GEN_KEYCORRIDOR, GEN_POOL, GEN_CROSSING, GEN_LAVAGAP, GEN_MULTIROOM, GEN_DISTSHIFT = range(10)
OBS_BYTES = 147
MAX_OBST = 8


class OrcConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "gen", "width", "height", "max_steps", "see_through", "n_actions",
        "n_obstacles", "room_size", "num_rows", "random_start", "lava_v1", "view_size", "hook", "gen_param0", "gen_param1")]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("minigrid_oracle.c", "minigrid_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_vec_create.restype = C.c_void_p
        L.orc_vec_create.argtypes = [C.POINTER(OrcConfig), C.c_uint64, C.c_int64, C.c_int32]
        L.orc_vec_destroy.argtypes = [C.c_void_p]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_vec_set_tape.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_vec_set_level_pool.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_vec_get_levels.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_vec_set_levels.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_vec_reset.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        L.orc_vec_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.orc_vec_rollout.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.orc_vec_rollout_random.argtypes = [C.c_void_p, C.c_int32, C.c_int] + [C.c_void_p] * 5
        L.orc_policy_action.argtypes = [C.c_uint64, C.c_int64, C.c_uint32, C.c_uint32, C.c_int]
        L.orc_vec_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.orc_vec_set_state.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.orc_philox4x32_10.argtypes = [C.c_void_p] * 3
        L.orc_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def philox(ctr, key):
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(c), _p(k), _p(out))
    return out


class OracleVec:
    """n independent reference-semantics envs stepped on the CPU."""

    def __init__(self, cfg: dict, n: int, seed: int = 0, env0: int = 0, threads: int = 0):
        self.cfg = dict(cfg)
        c = OrcConfig(**{k: int(cfg.get(k, 0)) for k, _ in OrcConfig._fields_})
        self.n = int(n)
        self.W, self.H = c.width, c.height
        self.V = c.view_size or 7
        self._L = lib()
        self._L.orc_set_threads(int(threads))
        self._h = self._L.orc_vec_create(C.byref(c), C.c_uint64(seed), C.c_int64(env0), self.n)
        if not self._h:
            raise RuntimeError(self._L.orc_last_error().decode())
        self._tape = None

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_vec_destroy(self._h)
            self._h = None

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self._L.orc_last_error().decode())

    def set_tape(self, draws, offsets):
        self._tape = (np.ascontiguousarray(draws, np.int32), np.ascontiguousarray(offsets, np.int64))
        self._chk(self._L.orc_vec_set_tape(self._h, _p(self._tape[0]), _p(self._tape[1])))

    def set_level_pool(self, grid, aux, agent, hook_params=None):
        g = np.ascontiguousarray(grid, np.uint8)
        K = g.shape[0]
        a = None if aux is None else np.ascontiguousarray(aux, np.uint8)
        ag = np.ascontiguousarray(np.asarray(agent)[:, :3], np.int32)
        hp = None if hook_params is None else np.ascontiguousarray(hook_params, np.int32).reshape(K, 16)
        self._chk(self._L.orc_vec_set_level_pool(self._h, K, _p(g), _p(a), _p(ag), _p(hp)))

    def get_levels(self):
        out = np.zeros(self.n, np.int32)
        self._chk(self._L.orc_vec_get_levels(self._h, _p(out)))
        return out

    def set_levels(self, levels):
        lv = np.ascontiguousarray(levels, np.int32)
        self._chk(self._L.orc_vec_set_levels(self._h, _p(lv)))

    def reset(self, mask=None):
        obs = np.zeros((self.n, self.V, self.V, 3), np.uint8)
        d = np.zeros(self.n, np.uint8)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self._chk(self._L.orc_vec_reset(self._h, _p(m), _p(obs), _p(d)))
        return obs, d

    def step(self, actions, autoreset=True):
        o, r, dn, d = self.rollout(np.asarray(actions, np.uint8).reshape(1, self.n), autoreset)
        return o[0], r[0], dn[0], d[0]

    def rollout(self, actions, autoreset=True, want_obs=True, out=None):
        """out = (obs, reward, done, dir) from an earlier call: reuse the (already touched) buffers"""
        a = np.ascontiguousarray(actions, np.uint8)
        T = a.shape[0]
        assert a.shape == (T, self.n)
        if out is not None:
            obs, r, dn, d = out
            assert obs.shape == (T, self.n, self.V, self.V, 3) and r.shape == dn.shape == d.shape == (T, self.n)
        else:
            obs = np.zeros((T, self.n, self.V, self.V, 3), np.uint8) if want_obs else None
            r = np.zeros((T, self.n), np.float64)
            dn = np.zeros((T, self.n), np.uint8)
            d = np.zeros((T, self.n), np.uint8)
        self._chk(self._L.orc_vec_rollout(self._h, T, _p(a), int(autoreset), _p(obs), _p(r), _p(dn), _p(d)))
        return obs, r, dn, d

    def rollout_random(self, T, autoreset=True, want_obs=True):
        """T steps under the counter-based uniform random policy; returns (obs, reward, done, dir, actions)"""
        obs = np.zeros((T, self.n, self.V, self.V, 3), np.uint8) if want_obs else None
        r = np.zeros((T, self.n), np.float64)
        dn = np.zeros((T, self.n), np.uint8)
        d = np.zeros((T, self.n), np.uint8)
        a = np.zeros((T, self.n), np.uint8)
        self._chk(self._L.orc_vec_rollout_random(self._h, T, int(autoreset), _p(a), _p(obs), _p(r), _p(dn), _p(d)))
        return obs, r, dn, d, a

    def get_state(self):
        n, W, H = self.n, self.W, self.H
        s = dict(grid=np.zeros((n, W, H, 3), np.uint8), aux=np.zeros((n, W, H), np.uint8),
                 agent=np.zeros((n, 4), np.int32), carrying=np.zeros((n, 3), np.uint8),
                 obstacles=np.zeros((n, MAX_OBST, 2), np.int16), target=np.zeros((n, 2), np.uint8),
                 rng=np.zeros((n, 2), np.uint32))
        self._chk(self._L.orc_vec_get_state(self._h, *[_p(s[k]) for k in
                  ("grid", "aux", "agent", "carrying", "obstacles", "target", "rng")]))
        return s

    def set_state(self, s):
        n, W, H = self.n, self.W, self.H
        g = np.ascontiguousarray(s["grid"], np.uint8).reshape(n, W, H, 3)
        aux = np.ascontiguousarray(s.get("aux", np.zeros((n, W, H), np.uint8)), np.uint8)
        ag = np.ascontiguousarray(s["agent"], np.int32).reshape(n, 4)
        ca = np.ascontiguousarray(s.get("carrying", np.zeros((n, 3), np.uint8)), np.uint8)
        ob = np.zeros((n, MAX_OBST, 2), np.int16)
        if "obstacles" in s and np.size(s["obstacles"]):
            o = np.asarray(s["obstacles"], np.int16).reshape(n, -1, 2)
            ob[:, :o.shape[1]] = o
        tg = np.ascontiguousarray(s.get("target", np.zeros((n, 2), np.uint8)), np.uint8)
        rng = np.ascontiguousarray(s.get("rng", np.zeros((n, 2), np.uint32)), np.uint32)
        self._chk(self._L.orc_vec_set_state(self._h, _p(g), _p(aux), _p(ag), _p(ca), _p(ob), _p(tg), _p(rng)))
