"""The visibility flood of the occluded kernels, checked as pure integer algebra on the CPU.

`observe()` in gym_minigrid_b200/csrc/mgb_kernels.cuh computes both sweeps of one view row of Grid.process_vis
(minigrid.py:617-648) with a single carry chain on a word that holds the row in bits 0..V-1 and, bit-reversed, in bits
31..32-V.  This file restates (a) the reference's two sweeps literally and (b) the kernel's formula, and compares them
for every (visible, see-through) pair of a row -- exhaustively for V <= 7, on a seeded sample above -- including the seeds
handed to the row above.  The CUDA code itself is compared with the oracle in the GPU suite; this pins the algebra.
"""
import random

import pytest


def reference_row(v, t, V):
    """one row j of process_vis: returns (visible mask of the row, cells made visible in row j-1)"""
    vis = [(v >> i) & 1 for i in range(V)]
    see = [(t >> i) & 1 for i in range(V)]
    up = [0] * V
    for i in range(0, V - 1):                       # minigrid.py:624-635
        if not vis[i] or not see[i]:
            continue
        vis[i + 1] = 1
        up[i + 1] = 1
        up[i] = 1
    for i in reversed(range(1, V)):                 # minigrid.py:637-648
        if not vis[i] or not see[i]:
            continue
        vis[i - 1] = 1
        up[i - 1] = 1
        up[i] = 1
    pack = lambda bits: sum(b << i for i, b in enumerate(bits))
    return pack(vis), pack(up)


def brev32(x):
    return int("{:032b}".format(x & 0xFFFFFFFF)[::-1], 2)


def kernel_row(vw, t, V):
    """the kernel's formula on the doubled word; returns (vis word, next seed word)"""
    M = 0xFFFFFFFF
    tw = (t | brev32(t)) & M
    fw = (((((vw & tw) + tw) & M) ^ tw) | vw) & M
    vis = (fw | brev32(fw)) & M
    sw = vis & tw
    return vis, (sw | (sw << 1) | (sw >> 1)) & M


def doubled(v):
    return (v | brev32(v)) & 0xFFFFFFFF


@pytest.mark.parametrize("V", [3, 5, 7, 9, 11])
def test_single_carry_chain_equals_the_two_sweeps(V):
    mask = (1 << V) - 1
    if V <= 7:
        pairs = [(v, t) for v in range(1 << V) for t in range(1 << V)]
    else:
        rs = random.Random(V)
        pairs = [(rs.randrange(1 << V), rs.randrange(1 << V)) for _ in range(20000)]
    for v, t in pairs:
        want_vis, want_up = reference_row(v, t, V)
        vis, nxt = kernel_row(doubled(v), t, V)
        assert vis & mask == want_vis, (V, v, t)
        assert nxt & mask == want_up, (V, v, t)
        # the high field mirrors the low one, so the next row can use the word as it is
        assert brev32(vis) & mask == want_vis and brev32(nxt) & mask == want_up, (V, v, t)


def test_rows_chain_like_process_vis():
    """seven rows in sequence, seeds carried in the doubled form (garbage bits 7 and 24 included), against the
    reference's row-by-row propagation"""
    V, mask = 7, 0x7F
    rs = random.Random(1)
    for _ in range(3000):
        ts = [rs.randrange(128) for _ in range(V)]
        v_ref, vw = 1 << (V // 2), doubled(1 << (V // 2))
        for j in reversed(range(V)):
            want_vis, want_up = reference_row(v_ref, ts[j], V)
            vis, vw = kernel_row(vw, ts[j], V)
            assert vis & mask == want_vis
            v_ref = want_up                       # vis_mask[., j-1] starts as what row j set (minigrid.py:631-633)


def test_packed_action_nibbles_round_trip():
    """the occluded kernels hold 32 steps of actions as 4-bit fields, clamped to 15 (still invalid: n_actions <= 9),
    and shift them out one per step with three funnel shifts"""
    rs = random.Random(3)
    for _ in range(200):
        acts = [rs.choice([0, 1, 2, 3, 4, 5, 6, 7, 15, 200, 255]) for _ in range(32)]
        w = [0, 0, 0, 0]
        for k, a in enumerate(acts):
            w[k >> 3] |= min(a, 15) << ((k & 7) * 4)
        out = []
        for _ in range(32):
            out.append(w[0] & 15)
            w = [((w[0] >> 4) | (w[1] << 28)) & 0xFFFFFFFF, ((w[1] >> 4) | (w[2] << 28)) & 0xFFFFFFFF,
                 ((w[2] >> 4) | (w[3] << 28)) & 0xFFFFFFFF, w[3] >> 4]
        assert out == [min(a, 15) for a in acts]
        assert all((o >= 9) == (a >= 9) for o, a in zip(out, acts))      # invalid stays invalid
