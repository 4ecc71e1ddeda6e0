"""Build libmgb200.so (CUDA kernels + C-ABI) in-tree for sm_100a with nvcc.

    python -m gym_minigrid_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libmgb200.so")
SOURCES = [os.path.join(CSRC, "mgb_api.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "mgb_kernels.cuh"), os.path.join(ROOT, "include", "mgb200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmgb200.so cannot be built (there is no CPU fallback)")


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(d) > t for d in DEPS)


def source_stamp():
    """git revision of the sources the library is built from ("+dirty" when csrc/ or include/ differ from it), so that
    every profile taken with a library can name the exact code (mgb_version() carries it)."""
    try:
        sha = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short=12", "HEAD"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                             text=True, timeout=10).stdout.strip()
        if not sha:
            return "unknown"
        dirty = subprocess.run(["git", "-C", ROOT, "status", "--porcelain", "--", "gym_minigrid_b200/csrc", "include"],
                               stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=10).stdout.strip()
        return sha + ("+dirty" if dirty else "")
    except Exception:
        return "unknown"


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: build an experimental variant (e.g. defines=["MGB_DYN_BLOCKS=6"]) next to the product .so"""
    if not force and not needs_build() and not defines:
        return SO
    target = out or SO
    stamp = ["-DMGB_BUILD_SHA=\"%s\"" % source_stamp(), "-DMGB_BUILD_DEFINES=\"%s\"" % (" ".join("-D" + d for d in defines) or "none")]
    cmd = [find_nvcc()] + NVCC_FLAGS + stamp + ["-D" + d for d in defines] + ["-I", os.path.join(ROOT, "include"), "-o", target] + SOURCES
    env = dict(os.environ)
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed (see %s)" % log)
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs, out=outs[0] if outs else None))
