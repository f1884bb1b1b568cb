"""In-tree build of libpnb200.so (hand-written CUDA for sm_100a, C ABI in include/pnb200.h).

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the resulting
.so sits next to this file and travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "csrc")
REPO = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libpnb200.so")
OBJDIR = os.path.join(CSRC, "build")

SOURCES = ["grid.cu", "sweep.cu", "nlist.cu", "hashgrid.cu", "hoststep.cu", "link.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # bit parity with the Julia reference: no FMA contraction, IEEE division and square root,
    # denormals kept.  The parity-critical code additionally uses
    # explicit *_rn intrinsics, so these flags are a second line of defence.
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    # the image exports CC=/opt/gcc/bin/gcc; pin the host compiler explicitly
    "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


HASHFILE = LIB + ".srchash"


def _source_hash() -> str:
    """Content hash of everything the binary is built from (file mtimes do not survive the copy
    to the GPU box, contents do)."""
    import hashlib
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    files.append(os.path.join(REPO, "include", "pnb200.h"))
    files.append(os.path.abspath(__file__))
    h = hashlib.sha256()
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def is_stale() -> bool:
    """True when there is no binary or it was built from other sources than the ones present."""
    if not os.path.exists(LIB) or not os.path.exists(HASHFILE):
        return True
    with open(HASHFILE) as fh:
        return fh.read().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link libpnb200.so.  Returns the library path."""
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []
    if os.environ.get("PNB_DIAG") == "1":      # measurement-only kernel variants (tools/)
        extra.append("-DPNB_DIAG")

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        cmd = [nvcc(), *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    cmd = [nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-ccbin", NVCC_FLAGS[NVCC_FLAGS.index("-ccbin") + 1]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(HASHFILE, "w") as fh:
        fh.write(_source_hash() + "\n")
    return LIB


def build_locked(force: bool = False, verbose: bool = False) -> str:
    """build() serialised across processes (ranks of one node share the source tree): the first
    process compiles, the others wait on the lock and then find a fresh binary."""
    import fcntl
    os.makedirs(OBJDIR, exist_ok=True)
    with open(os.path.join(OBJDIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return build(force=force, verbose=verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
