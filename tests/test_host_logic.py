"""CPU tests of the host side: registry (same ids / contract as the reference), C-ABI surface
(library loads, exports exactly what include/mgb200.h declares, fails loudly without a GPU),
config table vs the live reference (when mounted), sharding logic under a 2-process gloo group."""
import ctypes as C
import os
import re
import socket

import pytest

import gym_minigrid_b200 as mgb
from gym_minigrid_b200 import _lib, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADLINE = ["MiniGrid-Empty-8x8-v0", "MiniGrid-DoorKey-16x16-v0", "MiniGrid-FourRooms-v0",
            "MiniGrid-Dynamic-Obstacles-16x16-v0", "MiniGrid-KeyCorridorS6R3-v0"]


def header_functions():
    src = open(os.path.join(ROOT, "include", "mgb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mgb_[a-z_0-9]+)\s*\(", src)))


def test_registry_contract():
    for i in HEADLINE:
        assert i in mgb.env_list
    assert len(mgb.env_list) == len(set(mgb.env_list)) == 81      # 45 with device generators + 36 level-pool-only ids
    assert sum(1 for i in mgb.env_list if mgb.spec(i)["config"]["gen"] != 5) == 45
    with pytest.raises(AssertionError):            # register.py:12  id must start with "MiniGrid-"
        mgb.register("Foo-v0", "gym_minigrid.envs:EmptyEnv")
    with pytest.raises(AssertionError):            # register.py:13  ids are unique
        mgb.register("MiniGrid-Empty-8x8-v0", "gym_minigrid.envs:EmptyEnv")
    with pytest.raises(KeyError):
        mgb.spec("MiniGrid-MinimapForSparky-v0")    # out of scope (SAR env): loud, not silent
    assert mgb.spec("MiniGrid-SimpleRoom-v0")["config"]["gen"] == 5       # level-pool id (its generator draws from the global np.random)
    c = mgb.spec("MiniGrid-DistShift2-v0")["config"]                      # distshift.py:62-64: fixed layout
    assert (c["gen"], c["gen_param0"], c["see_through"]) == (9, 5, 1)
    c = mgb.spec("MiniGrid-MultiRoom-N6-v0")["config"]                    # multiroom.py:255-262: on-device generator
    assert (c["gen"], c["gen_param0"], c["gen_param1"], c["width"], c["max_steps"]) == (8, 6, 10, 25, 120)
    c = mgb.spec("MiniGrid-SimpleCrossingS11N5-v0")["config"]             # crossing.py:135-137
    assert (c["gen"], c["gen_param0"], c["gen_param1"]) == (6, 5, 2 | 4)
    c = mgb.spec("MiniGrid-LavaGapS6-v1")["config"]                       # lavagap.py:78-80 ('v1' in the class name)
    assert (c["gen"], c["gen_param0"], c["gen_param1"], c["lava_v1"]) == (7, 1, 0, 1)
    c = mgb.spec("MiniGrid-Dynamic-Obstacles-16x16-v0")["config"]
    assert (c["n_actions"], c["n_obstacles"], c["reward_range"], c["lava_v1"]) == (3, 8, (-1, 1), 1)
    assert mgb.spec("MiniGrid-KeyCorridorS6R3-v0")["config"]["max_steps"] == 1080
    assert mgb.spec("MiniGrid-FourRooms-v0")["config"]["max_steps"] == 500


def test_cabi_exports_match_header():
    names = header_functions()
    assert names == sorted(_lib.SIGNATURES), (names, sorted(_lib.SIGNATURES))
    lib = _lib.load()                               # dlopen works on a box without a GPU
    for n in names:
        assert hasattr(lib, n), n
    assert b"sm_100a" in lib.mgb_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.MgbError, match="no CPU fallback"):
        mgb.make("MiniGrid-Empty-8x8-v0", num_envs=4)
    lib = _lib.load()
    cfg = _lib.MgbConfig(**{k: int(mgb.spec("MiniGrid-Empty-8x8-v0")["config"].get(k, 0)) for k, _ in _lib.MgbConfig._fields_})
    h = C.c_void_p()
    assert lib.mgb_create(C.byref(cfg), 4, 0, 0, 0, C.byref(h)) != 0
    assert b"no CUDA device" in lib.mgb_last_error()


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under gym_minigrid_b200/ may import, link or load it"""
    pkg = os.path.join(ROOT, "gym_minigrid_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "minigrid_oracle" not in txt and "ref_shim" not in txt, f


def test_config_table_matches_live_reference():
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree not mounted")
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gen_golden
    for env_id in mgb.env_list:
        env = ref_shim.make(env_id)
        want = gen_golden.config_of(env)
        got = mgb.spec(env_id)["config"]
        for k, v in want.items():
            if got["gen"] == 5 and k in ("room_size", "num_rows"):
                continue                            # RoomGrid attributes are irrelevant without a device generator
            assert int(got[k]) == int(v), (env_id, k, got[k], v)
        assert tuple(got["reward_range"]) == tuple(env.reward_range), env_id
        obs = env.reset()
        if got["mission"] == "(per level)":
            pass
        elif "%s" not in got["mission"]:
            assert obs["mission"] == got["mission"], env_id
        else:
            assert obs["mission"].startswith("pick up the "), env_id


def test_shard_range():
    for total in (1, 7, 32, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(r, world, total) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (b0, c0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + c0 == b1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base, count = sharding.shard_range(rank, world, 1001)
    # rank r "processes" count*10 env-steps in (1 + r) seconds
    value, units, secs = sharding.aggregate_throughput(count * 10, 1.0 + rank)
    q.put((rank, base, count, value, units, secs))
    dist.destroy_process_group()


def test_two_rank_gloo_aggregation():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, b0, c0, v0, u0, s0), (r1, b1, c1, v1, u1, s1) = res
    assert (b0, c0, b1, c1) == (0, 501, 501, 500)
    assert u0 == u1 == 10010.0 and s0 == s1 == 2.0      # sum of units, MAX of time
    assert v0 == v1 == 10010.0 / 2.0


def test_header_is_plain_c(tmp_path):
    """include/mgb200.h is the drop-in boundary: it must compile as C (no C++, no torch/CUDA types in the signatures)."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "consumer.c"
    src.write_text('#include "mgb200.h"\n'
                   'int use(void) { mgb_config c = {0}; mgb_handle *h = 0; c.gen = MGB_GEN_EMPTY;\n'
                   '  return mgb_create(&c, 1, 0, 0, 0, &h) + mgb_step(h, 0, 0, 0, 0, 0, 0) + (int)mgb_num_envs(h); }\n')
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_two_draws_per_philox_word_are_jointly_uniform():
    """RNG design (DESIGN.md "RNG"): draw 2i is Philox word i, draw 2i+1 is the same word times DRAW_ODD_MULT mod 2^32.
    The x and y of a placement try are such a pair: their joint histogram over span_x x span_y cells must be flat
    (chi-square against the uniform distribution), for the 3x3 window of an obstacle move and for a 16x16 / 19x19 grid."""
    import numpy as np
    from oracle.ref_shim import PhiloxShim
    M = PhiloxShim.DRAW_ODD_MULT
    assert M % 2 == 1
    rs = np.random.RandomState(7)
    w = rs.randint(0, 1 << 32, size=1 << 20, dtype=np.uint64)           # stand-in for Philox words (uniform 32-bit)
    wm = (w * M) & 0xFFFFFFFF
    for sx, sy in ((3, 3), (2, 3), (16, 16), (19, 19), (64, 64)):
        x = (w * sx) >> 32
        y = (wm * sy) >> 32
        h = np.bincount((x * sy + y).astype(np.int64), minlength=sx * sy).astype(np.float64)
        e = len(w) / (sx * sy)
        chi2 = ((h - e) ** 2 / e).sum()
        dof = sx * sy - 1
        assert chi2 < dof + 6 * (2 * dof) ** 0.5, (sx, sy, chi2, dof)        # 6 sigma
    # and exhaustively on the lattice itself for the 3x3 window: every cell gets 2^32/9 words to within 1e-4
    u = np.arange(0, 1 << 32, 4099, dtype=np.uint64)                       # a coprime stride: 1M lattice points
    h = np.bincount((((u * 3) >> 32) * 3 + ((((u * M) & 0xFFFFFFFF) * 3) >> 32)).astype(np.int64), minlength=9)
    assert abs(h / h.sum() - 1 / 9).max() < 2e-3
    # the shim itself: consecutive randint(0,3) pairs
    sh = PhiloxShim(123, 5, 0)
    pairs = np.array([[sh.randint(0, 3), sh.randint(0, 3)] for _ in range(9000)])
    hh = np.bincount(pairs[:, 0] * 3 + pairs[:, 1], minlength=9)
    assert ((hh - 1000.0) ** 2 / 1000.0).sum() < 8 + 6 * 4
