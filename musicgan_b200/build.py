"""Build libmusicgan_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m musicgan_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repo snapshot.  No JIT, no torch
extension machinery: plain `nvcc -shared`, one translation unit per .cu, linked into one library.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmusicgan_b200.so")
DEBUG_LIB = os.path.join(HERE, "libmusicgan_b200_debug.so")      # test-only probes (debug_*.cu), never loaded by the product

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_digest() -> str:
    h = hashlib.sha1()
    roots = [CSRC, os.path.join(os.path.dirname(HERE), "include")]
    for root in roots:
        for f in sorted(os.listdir(root)):
            if f.endswith((".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(ARCH_FLAGS + COMMON).encode())
    return h.hexdigest()


def _compile_one(src: str, hdr_digest: str, force: bool) -> str:
    obj = os.path.join(BUILD, src[:-3] + ".o")
    stamp = obj + ".stamp"
    with open(os.path.join(CSRC, src), "rb") as fh:
        digest = hashlib.sha1(fh.read() + hdr_digest.encode()).hexdigest()
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
        return obj
    cmd = [NVCC, *ARCH_FLAGS, *COMMON, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(digest)
    return obj


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(BUILD, exist_ok=True)
    srcs = _sources()
    hdr = _headers_digest()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, hdr, force), srcs))
    is_debug = lambda o: os.path.basename(o).startswith("debug_")
    product = [o for o in objs if not is_debug(o)]
    # the probes need the library's error / profiling helpers: lib.o is linked into both
    debug = [o for o in objs if is_debug(o) or os.path.basename(o) == "lib.o"]
    for lib, members in ((LIB, product), (DEBUG_LIB, debug)):
        newest = max(os.path.getmtime(o) for o in members)
        if force or not os.path.exists(lib) or os.path.getmtime(lib) < newest:
            cmd = [NVCC, *ARCH_FLAGS, "-shared", "-o", lib, *members]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
            if verbose:
                print(f"[musicgan_b200.build] linked {lib} from {len(members)} objects")
        elif verbose:
            print(f"[musicgan_b200.build] {lib} up to date")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
