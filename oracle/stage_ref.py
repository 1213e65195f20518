"""Stage the reference's own Python sources for the GPU box.  Test infrastructure.

The reference is pure Python, so "running the real reference" needs nothing but its ``src/`` tree and the
external shims of ``oracle/refimport.py``.  ``gpurun`` and the round-end driver ship the working tree of this repo only, so
``python -m oracle.stage_ref`` (also called by ``__graft_entry__.build()`` wherever ``/root/reference`` exists) copies
``/root/reference/src`` to the GIT-IGNORED ``oracle/_ref/src`` -- it travels like a built ``.so``, and it is never committed
(no reference source enters the history).  ``oracle/refimport.py`` falls back to that copy when ``/root/reference`` is absent;
``tests/test_gpu_models.py::test_patched_reference_*`` and ``bench.py --impl reference`` use it there.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("VITGAN_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")


def stage(verbose: bool = True) -> bool:
    src = os.path.join(SRC, "src")
    if not os.path.isdir(os.path.join(src, "v2")):
        if verbose:
            print(f"[stage_ref] {src} not found: nothing staged", file=sys.stderr)
        return False
    dst = os.path.join(DST, "src")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    n = sum(len(f) for _, _, f in os.walk(dst))
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write(f"{SRC}/src ({n} files); git-ignored, regenerate with `python -m oracle.stage_ref`\n")
    if verbose:
        print(f"[stage_ref] {n} files -> {dst}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
