"""
Builds libmrc_b200.so in-tree with nvcc for sm_100a (B200).  No JIT cache: the .so sits next to
this file so it travels with the repo snapshot to the GPU box.

    python -m min_ratio_cycle_b200.build [--force] [--verbose]

Two translation units (the CUB-based array build compiles slowly and rarely changes); objects are cached
under min_ratio_cycle_b200/_obj and rebuilt only when a source or header they include is newer.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libmrc_b200.so")
HEADER = os.path.join(ROOT, "include", "mrc_b200.h")
UNITS = {   # source -> headers it depends on
    "mrc_b200.cu": ["mrc_kernels.cuh", "mrc_probe.cuh", "mrc_batch.cuh", "mrc_search.h", "mrc_host.h", "mrc_comm.cuh"],
    "mrc_build.cu": ["mrc_build.cuh", "mrc_host.h"],
}
SOURCES = [os.path.join(CSRC, s) for s in UNITS]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _base(verbose: bool, defines) -> list[str]:
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC,-O2", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    return cmd + [f"-D{d}" for d in (defines or [])]


def command(verbose: bool = False, defines: list[str] | None = None, out: str | None = None) -> list[str]:
    """Single-step command that compiles every unit (used for -Xptxas -v listings and tuning variants)."""
    return _base(verbose, defines) + ["--shared", "-o", out or LIB] + SOURCES


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))


def up_to_date() -> bool:
    deps = SOURCES + [os.path.join(CSRC, h) for hs in UNITS.values() for h in hs] + [HEADER]
    return not _stale(LIB, deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    for src, hdrs in UNITS.items():
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        deps = [os.path.join(CSRC, src), HEADER] + [os.path.join(CSRC, h) for h in hdrs]
        if _stale(obj, deps) or (force and src != "mrc_build.cu"):
            _run(_base(verbose, None) + ["-c", os.path.join(CSRC, src), "-o", obj], verbose)
        objs.append(obj)
    _run(_base(False, None) + ["--shared", "-o", LIB] + objs, verbose)
    return LIB


def build_variant(tag: str, defines: list[str]) -> str:
    """Tuning builds: extra -D defines for the kernel unit, written as libmrc_b200.<tag>.so"""
    out = os.path.join(HERE, f"libmrc_b200.{tag}.so")
    os.makedirs(OBJ, exist_ok=True)
    obj = os.path.join(OBJ, f"mrc_b200.{tag}.o")
    _run(_base(False, defines) + ["-c", os.path.join(CSRC, "mrc_b200.cu"), "-o", obj], False)
    build()
    _run(_base(False, None) + ["--shared", "-o", out, obj, os.path.join(OBJ, "mrc_build.o")], False)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
