"""In-tree build of libatmonr_b200.so (sm_100a) and the host self-check library.

    python atmospheric-neural-rendering_b200/build.py [--force] [--verbose]

The CUDA source is compiled in parts in parallel (ATM_PART) and linked into one shared
library under lib/. Objects are cached under build/ keyed by a hash of all sources + flags.

Flags: -fmad=false (nvcc) / -ffp-contract=off (g++) so that every a*b+c written as two
operations rounds twice, like the eager torch ops of the reference; fused multiply-adds in
the kernels are explicit (fmaf, tcgen05.mma).
"""

from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "lib"
BUILD = HERE / "build"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo", "-fmad=false",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
] + os.environ.get("ATMONR_NVCC_EXTRA", "").split()   # tuning experiments (-DATM_FWD_CTAS=8 ...); part of the cache key
# (object name, source, extra defines)
UNITS = [
    ("basic", "atmonr_b200.cu", ["-DATM_PART=0"]),
    ("mlp", "atmonr_b200.cu", ["-DATM_PART=1"]),
    ("field", "atmonr_b200.cu", ["-DATM_PART=2"]),
    ("surf", "atmonr_b200.cu", ["-DATM_PART=3"]),
    ("fused", "ngp_fused.cu", []),
    ("rays", "rays.cu", []),
    ("linear", "linear_tc.cu", []),
    ("nerfpts", "nerf_points.cu", []),
]


def _digest(src: str, extra: list[str]) -> str:
    """Hash of one source file + every header + the flags."""
    h = hashlib.sha256()
    deps = [CSRC / src] + sorted(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "atmonr_b200.h"]
    for p in deps:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS + extra).encode())
    return h.hexdigest()[:16]


def _run(cmd: list[str], verbose: bool) -> None:
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr)


def build(force: bool = False, verbose: bool = False) -> Path:
    LIB.mkdir(exist_ok=True)
    BUILD.mkdir(exist_ok=True)
    so = LIB / "libatmonr_b200.so"
    host_so = LIB / "libatmonr_hostcheck.so"
    units = [u for u in UNITS if (CSRC / u[1]).exists()]
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_unit(u):
        name, src, defs = u
        obj = BUILD / f"{name}.o"
        tag = BUILD / f"{name}.hash"
        digest = _digest(src, defs)
        if not force and obj.exists() and tag.exists() and tag.read_text() == digest:
            return obj, False
        _run([NVCC, *NVCC_FLAGS, *extra, *defs, "-c", str(CSRC / src), "-o", str(obj)], verbose)
        tag.write_text(digest)
        return obj, True

    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_unit, units))
    objs = [o for o, _ in results]
    host_tag = BUILD / "hostcheck.hash"
    host_digest = _digest("hostcheck.cpp", ["g++"])
    if force or not host_so.exists() or not host_tag.exists() or host_tag.read_text() != host_digest:
        _run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
              "-o", str(host_so), str(CSRC / "hostcheck.cpp"), "-lm"], verbose)
        host_tag.write_text(host_digest)
    if so.exists() and not any(changed for _, changed in results):
        return so
    _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(so), *map(str, objs)], verbose)
    return so


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(f"built {path}")
