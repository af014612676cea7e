"""Build libcpetb200.so in-tree with nvcc for sm_100a (no GPU needed to compile).

    python -m pycpet_b200.build [--force] [--verbose]

The shared object lands next to this file (pycpet_b200/libcpetb200.so); it is git-ignored but
travels with the working tree, which is how it reaches the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libcpetb200.so")
SOURCES = ["capi.cu", "field.cu", "topo.cu", "topo8.cu", "hist.cu", "legacy.cu", "textio.cu"]
HEADERS = ["common.cuh", "cpet_internal.h", os.path.join("..", "..", "include", "cpet_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libcpetb200.so cannot be built")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS] + [os.path.abspath(__file__)]
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a wrapper gcc; nvcc should use the system host compiler
    host_cxx = shutil.which("g++") or "g++"
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + hdrs):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        extra = os.environ.get("CPET_NVCC_EXTRA", "").split()     # e.g. -DCPET_K2_MAXT=768 for experiments
        cmd = [nvcc, "-ccbin", host_cxx] + NVCC_FLAGS + extra + ["-c", s, "-o", o]
        p = subprocess.run(cmd, capture_output=True, text=True, env=env)
        return s, p.returncode, p.stdout + p.stderr

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        for s, rc, log in ex.map(compile_one, jobs):
            if verbose or rc != 0:
                sys.stderr.write(f"--- nvcc {os.path.basename(s)} ---\n{log}\n")
            if rc != 0:
                raise RuntimeError(f"nvcc failed on {s}")
    objs = [os.path.join(OBJ, src.replace(".cu", ".o")) for src in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-ccbin", host_cxx, "-shared", "-cudart", "static",
               "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        p = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if p.returncode != 0:
            sys.stderr.write(p.stdout + p.stderr)
            raise RuntimeError("link of libcpetb200.so failed")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(path)
