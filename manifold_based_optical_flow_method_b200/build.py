"""Builds csrc/libmof_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m manifold_based_optical_flow_method_b200.build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(CSRC, "libmof_b200.so")

SOURCES = ["error.cpp", "pattern.cpp", "csvio.cpp", "geom.cu", "assemble.cu", "pcg.cu", "detect.cu", "wave.cu", "winding.cu", "rbf.cu"]
HEADERS = ["mof_error.h", "mof_bodies.h", "mof_common.cuh", os.path.join(INCLUDE, "mof_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-pthread", "-shared", "--cudart", "static",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libmof_b200.so")


STAMP = LIB + ".srchash"


def _dep_paths():
    return [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]


def source_hash():
    """sha256 over the compiler flags and every source / header the library is built from."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for path in _dep_paths():
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def needs_build():
    """True when the library is missing or was built from other sources.  Decided on content, not on
    modification times: a copied tree (the GPU box receives a snapshot) keeps no reliable mtimes."""
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != source_hash()


def build(force=False, verbose=False):
    """Compile every CUDA source of the package for sm_100a -> csrc/libmof_b200.so (written under a
    temporary name and renamed into place, so a concurrent reader never maps a half-written file)."""
    if not force and not needs_build():
        return LIB
    tmp = f"{LIB}.{os.getpid()}.tmp"
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
        "-I", INCLUDE, "-I", CSRC, "-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
    try:
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed building libmof_b200.so")
        os.replace(tmp, LIB)
        with open(STAMP, "w") as fh:
            fh.write(source_hash() + "\n")
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    return LIB


def build_locked(force=False, verbose=False):
    """build() under an exclusive file lock: of several processes that find the library stale at the same
    moment (one rank per GPU under torchrun) one compiles, the others wait and then find it fresh."""
    import fcntl
    with open(os.path.join(CSRC, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return build(force=force, verbose=verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


if __name__ == "__main__":
    print(build_locked(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
