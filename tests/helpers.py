"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import glob
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load(path):
    z = np.load(path, allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["cfg"] = {str(k): int(v) for k, v in zip(d["cfg_keys"], d["cfg_vals"])}
    d["env_id"] = str(d["env_id"])
    return d


def bits(x):
    """float64 -> uint64 bit pattern (rewards are compared bit for bit)."""
    return np.ascontiguousarray(x, np.float64).view(np.uint64)


def assert_same(name, got, want):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, "%s: shape %s vs %s" % (name, got.shape, want.shape)
    if not np.array_equal(got, want):
        bad = np.argwhere(got != want)
        raise AssertionError("%s: %d mismatches, first at %s: got %s want %s" % (
            name, len(bad), tuple(bad[0]), got[tuple(bad[0])], want[tuple(bad[0])]))
