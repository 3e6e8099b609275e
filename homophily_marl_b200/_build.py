"""In-tree build of the CUDA library (``libssd_b200.so``) for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the CPU-only dev container;
the built ``.so`` is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
SRC = os.path.join(PKG_DIR, "csrc", "ssd_b200.cu")
SRCS = [SRC, os.path.join(PKG_DIR, "csrc", "policy_b200.cu"), os.path.join(PKG_DIR, "csrc", "frontend_b200.cu")]
HDR = os.path.join(ROOT, "include", "ssd_b200.h")
LIB = os.path.join(PKG_DIR, "libssd_b200.so")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SRCS + [HDR])


LIB_CHECK = os.path.join(PKG_DIR, "libssd_b200_check.so")


def build_checked(force: bool = False) -> str:
    """Variant with -DSSD_BOUNDS_CHECK (own shared-memory index checks; see ssd_debug_oob_count)."""
    if not force and os.path.exists(LIB_CHECK) and os.path.getmtime(LIB_CHECK) >= max(os.path.getmtime(f) for f in SRCS + [HDR]):
        return LIB_CHECK
    return build(force=True, extra_flags=["-DSSD_BOUNDS_CHECK"], out=LIB_CHECK)


def build(force: bool = False, extra_flags=(), verbose: bool = False, out: str = LIB) -> str:
    if not force and out == LIB and not needs_build():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra_flags) + ["-I", os.path.join(ROOT, "include"), "-o", out] + SRCS
    env = dict(os.environ)
    if os.path.exists("/usr/bin/gcc"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose or r.returncode != 0:
        print(" ".join(cmd))
        print(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stderr[-4000:])
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print(build_checked(force=True))                       # keep the checked variant in step with the source
