"""Build libdrag_b200.so (sm_100a only) in-tree with nvcc.

    python ai-dial-rag_b200/csrc/build.py [--force] [--verbose]

Every ``*.cu`` in this directory is compiled to an object (in parallel, rebuilt
only when the source, a header or the flags changed) and linked into
``ai-dial-rag_b200/dial_rag_b200/_lib/libdrag_b200.so``.  The .so is git-ignored but
travels to the GPU box with the repo snapshot.
"""

from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(os.path.dirname(HERE), "dial_rag_b200")
OUT_DIR = os.path.join(PKG, "_lib")
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(OUT_DIR, "libdrag_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-DDRAG_BUILD",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _digest(paths, extra: str) -> str:
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    sources = sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))
    headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "drag_b200.h"))
    flags = list(NVCC_FLAGS)
    if verbose:
        flags += ["-Xptxas", "-v"]
    jobs = []
    objs = []
    for src in sources:
        path = os.path.join(HERE, src)
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        stamp = obj + ".sha"
        want = _digest([path] + headers, " ".join(NVCC_FLAGS))
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == want:
            continue
        jobs.append((path, obj, stamp, want))

    def compile_one(job):
        path, obj, stamp, want = job
        cmd = [nvcc()] + flags + ["-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {path}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        with open(stamp, "w") as f:
            f.write(want)
        return path

    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for done in ex.map(compile_one, jobs):
                if verbose:
                    print("compiled", done, file=sys.stderr)
    if jobs or not os.path.exists(LIB):
        # the driver API (cuTensorMapEncodeTiled) is resolved at run time through
        # cudaGetDriverEntryPoint, so libcuda is not a link-time dependency
        tmp = LIB + ".tmp"   # link next to the target, then rename: a repo snapshot never sees a half-written library
        cmd = [nvcc(), "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
