"""ctypes binding of libmgb200.so (include/mgb200.h).  Fails loudly when the CUDA extension
is missing -- there is no CPU fallback and nothing here ever imports oracle/."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("MGB_LIB") or os.path.join(_HERE, "libmgb200.so")   # MGB_LIB: kernel-variant experiments

OBS_BYTES = 147
MAX_OBSTACLES = 8

ERROR_BITS = {
    1: "unknown action (reference: assert False, minigrid.py:1316-1318)",
    2: "RNG tape exhausted",
    4: "RNG tape value outside [low, high)",
    8: "rejection sampling gave up (reference: RecursionError)",
    16: "agent / cell index out of bounds",
    32: "unsupported cell code in set_state",
    64: "reset without a level pool",
}


class MgbConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "gen", "width", "height", "max_steps", "see_through", "n_actions",
        "n_obstacles", "room_size", "num_rows", "random_start", "lava_v1", "agent_view_size", "hook", "gen_param0", "gen_param1")]


# name -> (restype, argtypes); the single source of truth checked against include/mgb200.h by the tests
_P = C.c_void_p
SIGNATURES = {
    "mgb_version": (C.c_char_p, []),
    "mgb_last_error": (C.c_char_p, []),
    "mgb_create": (C.c_int, [C.POINTER(MgbConfig), C.c_int64, C.c_int, C.c_uint64, C.c_int64, C.POINTER(_P)]),
    "mgb_destroy": (C.c_int, [_P]),
    "mgb_num_envs": (C.c_int64, [_P]),
    "mgb_set_autoreset": (C.c_int, [_P, C.c_int]),
    "mgb_seed": (C.c_int, [_P, C.c_uint64]),
    "mgb_reset": (C.c_int, [_P, _P, _P, _P, _P]),
    "mgb_step": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "mgb_rollout": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "mgb_rollout_random": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "mgb_step_host": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "mgb_set_state": (C.c_int, [_P, C.c_int64, C.c_int64] + [_P] * 8),
    "mgb_get_state": (C.c_int, [_P, C.c_int64, C.c_int64] + [_P] * 8),
    "mgb_set_level_pool": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, _P]),
    "mgb_get_levels": (C.c_int, [_P, _P, _P]),
    "mgb_set_levels": (C.c_int, [_P, _P, _P]),
    "mgb_set_rng_tape": (C.c_int, [_P, _P, _P]),
    "mgb_full_obs": (C.c_int, [_P, _P, _P]),
    "mgb_onehot": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "mgb_flat_obs": (C.c_int, [_P, C.c_int32, _P, C.c_int32, _P, _P, C.c_int64, _P]),
    "mgb_render_partial": (C.c_int, [_P, C.c_int32, _P, C.c_int32, _P, C.c_int64, _P]),
    "mgb_render_full": (C.c_int, [_P, _P, _P, C.c_int32, _P, _P]),
    "mgb_visit_bonus": (C.c_int, [_P, C.c_int32, _P, _P, C.c_int64, _P, _P]),
    "mgb_dac": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mgb_append_action": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "mgb_goal_policy": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "mgb_episode_stats": (C.c_int, [C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mgb_error_flags": (C.c_int, [_P, _P, C.POINTER(C.c_uint32)]),
    "mgb_kernel_launches": (C.c_int64, [_P]),
    "mgb_set_kernel_timing": (C.c_int, [_P, C.c_int]),
    "mgb_last_kernel_ms": (C.c_double, [_P]),
}

_lib = None


class MgbError(RuntimeError):
    pass


def load():
    """dlopen libmgb200.so and bind every entry point of include/mgb200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise MgbError(
            "libmgb200.so not found at %s. Build it with `python -m gym_minigrid_b200.build` "
            "(needs nvcc; sm_100a only). There is no CPU fallback." % SO_PATH)
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise MgbError(load().mgb_last_error().decode())
