"""Build libclift_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m contrastive_lift_b200.build [--force] [--verbose]

The shared library is a plain C-ABI object (include/clift_b200.h); it is git-ignored but travels
to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libclift_b200.so")
STAMP = os.path.join(HERE, "csrc", ".build_stamp")
SOURCES = ["api.cu", "march.cu", "heads.cu", "heads_tc.cu", "heads_tc16.cu", "heads_x16.cu", "wgrad_tc.cu", "backward.cu", "pack.cu", "rays.cu", "loss.cu", "epoch.cu", "collective.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _headers_digest() -> "hashlib._Hash":
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cuh", ".h")):
            h.update(f.encode())
            h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(HERE, "..", "include", "clift_b200.h"), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h


def _source_digest(src: str, headers) -> str:
    h = headers.copy()
    h.update(open(os.path.join(CSRC, src), "rb").read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Recompiles only the sources whose text (or any header, or the flags) changed since their object was built."""
    headers = _headers_digest()
    want = {src: _source_digest(src, headers) for src in SOURCES}
    stamp = {}
    if os.path.exists(STAMP) and not force:
        for line in open(STAMP).read().splitlines():
            k, _, v = line.partition(" ")
            stamp[k] = v
    obj_of = lambda src: os.path.join(CSRC, src.replace(".cu", ".o"))
    todo = [src for src in SOURCES if stamp.get(src) != want[src] or not os.path.exists(obj_of(src))]
    if not todo and os.path.exists(LIB):
        return LIB
    nvcc = _nvcc()
    procs = []
    for src in todo:
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj_of(src)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f"--- {src}\n{out}", flush=True)
        if p.returncode != 0:
            failed = True
        else:
            stamp[src] = want[src]
    open(STAMP, "w").write("".join(f"{k} {v}\n" for k, v in stamp.items() if k in want))
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *[obj_of(src) for src in SOURCES], "-ldl"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
