"""GPU: the C-ABI is self-sufficient -- a binding written from include/mgb200.h alone (ctypes + raw device pointers, none
of the Python facade) drives the library and gets the same bytes as the facade.  This is the stub INTEGRATION.md §2 shows
a reference maintainer."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Config(C.Structure):      # mgb_config, include/mgb200.h
    _fields_ = [(n, C.c_int32) for n in ("gen", "width", "height", "max_steps", "see_through", "n_actions", "n_obstacles",
                                          "room_size", "num_rows", "random_start", "lava_v1", "agent_view_size", "hook")]


def test_raw_ctypes_binding_matches_facade():
    lib = C.CDLL(os.environ.get("MGB_LIB") or os.path.join(ROOT, "gym_minigrid_b200", "libmgb200.so"))
    lib.mgb_last_error.restype = C.c_char_p
    lib.mgb_create.argtypes = [C.POINTER(Config), C.c_int64, C.c_int, C.c_uint64, C.c_int64, C.POINTER(C.c_void_p)]
    lib.mgb_reset.argtypes = [C.c_void_p] * 5
    lib.mgb_step.argtypes = [C.c_void_p] * 7
    lib.mgb_destroy.argtypes = [C.c_void_p]
    # MiniGrid-DoorKey-8x8-v0: DoorKeyEnv(size=8), max_steps = 10*size*size (envs/doorkey.py:10-13)
    cfg = Config(gen=1, width=8, height=8, max_steps=640, see_through=0, n_actions=7, agent_view_size=7)
    N = 1000
    h = C.c_void_p()
    assert lib.mgb_create(C.byref(cfg), N, 0, 42, 5, C.byref(h)) == 0, lib.mgb_last_error()
    obs = torch.empty((N, 7, 7, 3), dtype=torch.uint8, device="cuda")
    d = torch.empty(N, dtype=torch.uint8, device="cuda")
    rew = torch.empty(N, dtype=torch.float64, device="cuda")
    done = torch.empty(N, dtype=torch.uint8, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def p(t):
        return C.c_void_p(t.data_ptr())

    import gym_minigrid_b200 as mgb
    env = mgb.make("MiniGrid-DoorKey-8x8-v0", num_envs=N, seed=42, env_id_base=5)
    assert lib.mgb_reset(h, None, p(obs), p(d), stream) == 0, lib.mgb_last_error()
    ref = env.reset()
    assert torch.equal(obs, ref["image"]) and torch.equal(d, ref["direction"])
    g = torch.Generator().manual_seed(0)
    for _ in range(30):
        a = torch.randint(0, 7, (N,), dtype=torch.uint8, generator=g).cuda()
        assert lib.mgb_step(h, p(a), p(obs), p(rew), p(done), p(d), stream) == 0, lib.mgb_last_error()
        o2, r2, d2, _ = env.step(a)
        assert torch.equal(obs, o2["image"]) and torch.equal(d, o2["direction"])
        assert torch.equal(rew.view(torch.int64), r2.view(torch.int64)) and torch.equal(done.bool(), d2)
    # errors come back as codes + message, never as exceptions across the boundary
    assert lib.mgb_step(h, None, p(obs), p(rew), p(done), p(d), stream) != 0
    assert b"actions" in lib.mgb_last_error()
    assert lib.mgb_destroy(h) == 0
