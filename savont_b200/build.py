"""Builds libsavont_b200.so (sm_100a only) in-tree with nvcc.  No JIT cache, no other arch."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libsavont_b200.so")
SOURCES = ["sketch.cu", "sketch_fast.cu", "prefilter.cu", "kmercount.cu", "filter.cu", "align.cu", "assign.cu", "pileup.cu", "cluster.cu", "reassign.cu", "snpcluster.cu", "api.cu", "api_kmer.cu", "api_pileup.cu", "api_cluster.cu", "api_reassign.cu", "api_multi.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not os.path.isdir(CSRC):
        if os.path.exists(LIB_PATH):
            return LIB_PATH
        raise RuntimeError("sources and library both missing")
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "sketch_dev.cuh"), os.path.join(CSRC, "api_internal.cuh"), os.path.join(HERE, "..", "include", "savont_b200.h")]
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a gcc without a full install; let nvcc find /usr/bin/g++
    ccbin = shutil.which("g++") or "/usr/bin/g++"

    def compile_one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            cmd = [nvcc, "-ccbin", ccbin] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True, env=env)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                sys.stderr.write(r.stderr)
        return o

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-ccbin", ccbin, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


HOST_LIB = os.path.join(LIB_DIR, "libsavont_host.so")


def build_host(force: bool = False) -> str:
    """the C++ host layer above the C ABI (savont_b200/host); links against libsavont_b200.so"""
    src = os.path.join(HERE, "host", "savont_host.cpp")
    hdr = os.path.join(HERE, "host", "savont_host.hpp")
    if not os.path.exists(src):
        return HOST_LIB
    if force or _stale(HOST_LIB, [src, hdr, LIB_PATH]):
        cxx = shutil.which("g++") or "/usr/bin/g++"
        cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", HOST_LIB, src, "-L" + LIB_DIR, "-lsavont_b200",
               "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"host build failed:\n{r.stdout}\n{r.stderr}")
    return HOST_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host(force="--force" in sys.argv))
