"""
Builds libmrc_b200.so in-tree with nvcc for sm_100a (B200).  No JIT cache: the .so sits next to
this file so it travels with the repo snapshot to the GPU box.

    python -m min_ratio_cycle_b200.build [--force] [--verbose]
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libmrc_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "mrc_b200.cu")]
DEPS = SOURCES + [os.path.join(HERE, "csrc", f) for f in ("mrc_kernels.cuh", "mrc_batch.cuh", "mrc_search.h")] + [
    os.path.join(ROOT, "include", "mrc_b200.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def command(verbose: bool = False, defines: list[str] | None = None, out: str | None = None) -> list[str]:
    cmd = [
        nvcc_path(), "-O3", "-std=c++17",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-lineinfo", "--shared", "-Xcompiler", "-fPIC,-O2",
        "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "csrc"),
        "-o", out or LIB,
    ] + [f"-D{d}" for d in (defines or [])] + SOURCES
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    return cmd


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or not up_to_date():
        cmd = command(verbose)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    return LIB


def build_variant(tag: str, defines: list[str]) -> str:
    """Tuning builds: extra -D defines, written next to the main library as libmrc_b200.<tag>.so"""
    out = os.path.join(HERE, f"libmrc_b200.{tag}.so")
    r = subprocess.run(command(False, defines, out), capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed for variant " + tag)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
