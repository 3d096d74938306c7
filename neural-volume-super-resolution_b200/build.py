"""Build recipe for libnvsr_b200.so (sm_100a only, in-tree so the .so travels to the GPU box).

`python -m ...build` is not needed: `__graft_entry__.build()` and `_lib.load()` call `build_library()`.
"""
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
# NVSR_B200_LIB selects a prebuilt library (A/B timing of kernel variants); it is never rebuilt
LIB_OVERRIDE = os.environ.get("NVSR_B200_LIB")
LIB_PATH = LIB_OVERRIDE or os.path.join(PKG_DIR, "libnvsr_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "--threads", "0",          # the .cu files compile in parallel (one nvcc invocation, all host cores)
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(REPO_DIR, "include", "nvsr.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library with nvcc (cross-compiles without a GPU).

    Safe under several processes (torchrun ranks, pytest-xdist): the build is serialised with an exclusive file lock,
    a process that waited re-checks staleness (the winner has built it), and the library appears atomically
    (compiled to a temporary name, then renamed) — a reader never maps a half-written file."""
    if LIB_OVERRIDE or (not force and not _stale()):
        return LIB_PATH
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return LIB_PATH
            nvcc = os.environ.get("NVCC", "nvcc")
            tmp = "%s.tmp.%d" % (LIB_PATH, os.getpid())
            cmd = [nvcc] + NVCC_FLAGS + ["-I", os.path.join(REPO_DIR, "include"), "-I", CSRC, "-o", tmp] + sources()
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
                print(" ".join(cmd), file=sys.stderr)
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB_PATH)
            if verbose:
                print(res.stderr, file=sys.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
