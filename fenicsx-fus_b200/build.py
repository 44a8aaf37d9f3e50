"""Build the CUDA extension in-tree for sm_100a (and nothing else).

    python fenicsx-fus_b200/build.py [--force] [--verbose]

Output: fenicsx-fus_b200/lib/libfus_b200.so (git-ignored; travels to the GPU box with the tree).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libfus_b200.so")
SOURCES = ["fus_capi.cu", "fus_halo.cu", "fus_host.cpp", "fus_partition.cpp"]
HEADERS = ["fus_kernels.cuh", "fus_internal.hpp", "fus_halo.hpp", "fus_trilinear.hpp",
           "fus_halo_kernels.cuh", "fus_cell_kernel.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xcompiler", "-fopenmp",          # host set-up loops (fus_partition.cpp)
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _source_hash():
    """Content hash of everything the library is built from (mtimes do not survive a snapshot)."""
    import hashlib
    h = hashlib.sha256()
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(ROOT, "include", "fus_b200.h"))
    for d in deps:
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


HASHFILE = os.path.join(LIBDIR, "libfus_b200.hash")


def _stale():
    if not os.path.exists(LIB) or not os.path.exists(HASHFILE):
        return True
    with open(HASHFILE) as f:
        return f.read().strip() != _source_hash()


def build_library(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + [
        "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-o", LIB,
    ] + [os.path.join(CSRC, f) for f in SOURCES] + ["-ldl", "-lgomp"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    with open(HASHFILE, "w") as f:
        f.write(_source_hash())
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
