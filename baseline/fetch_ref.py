#!/usr/bin/env python
"""Copies the UNMODIFIED reference sources to the git-ignored ``baseline/_ref/`` so they travel to the GPU box.

``/root/reference`` exists only in the dev container; ``gpurun`` ships ``/root/repo`` (git-ignored files included,
``.gpurunignore`` files excluded).  ``baseline/_ref/`` is listed in ``.gitignore`` and is NEVER committed: it is the
reference itself, used (a) as the CPU baseline that ``bench.py`` times on the box's host cores, (b) as the unchanged
runner / MAC / learner that the ``-m gpu`` training-loop tests drive on top of the CUDA env.

    python baseline/fetch_ref.py            # idempotent; prints the tree hash
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SSD_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")


def tree_hash(root: str) -> str:
    h = hashlib.sha256()
    for d, dirs, files in sorted(os.walk(root)):
        dirs[:] = sorted(x for x in dirs if x != "__pycache__")
        for f in sorted(files):
            if f.endswith((".py", ".yaml")):
                p = os.path.join(d, f)
                h.update(os.path.relpath(p, root).encode())
                with open(p, "rb") as fh:
                    h.update(fh.read())
    return h.hexdigest()


def fetch(force: bool = False) -> str | None:
    """Returns the hash of the copied tree, or None when no reference is mounted (the GPU box)."""
    src = os.path.join(SRC, "src")
    if not os.path.isdir(src):
        return None
    dst = os.path.join(DST, "src")
    if os.path.isdir(dst) and not force and tree_hash(dst) == tree_hash(src):
        return tree_hash(dst)
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    got, want = tree_hash(dst), tree_hash(src)
    if got != want:
        raise RuntimeError("baseline/_ref differs from the reference after the copy")
    return got


if __name__ == "__main__":
    h = fetch(force="--force" in sys.argv)
    print("no reference mounted; baseline/_ref left as is" if h is None else f"baseline/_ref/src == reference src ({h[:16]})")
