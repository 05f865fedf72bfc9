"""Builds libmarinenav_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot).

    python -m distributional_rl_navigation_b200.build [--force]
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
SO = os.path.join(LIBDIR, "libmarinenav_b200.so")
OBJDIR = os.path.join(PKG, "build")
HOST_SRC = os.path.join(PKG, "csrc_host", "mnv_host.c")
HOST_SO = os.path.join(LIBDIR, "libmnv_host.so")      # native host-side helper of step_host (plain C, pthreads)

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
# per-file extra flags: the reset kernel reproduces numpy's separately rounded arithmetic bit-for-bit
EXTRA = {"mnv_reset.cu": ["-fmad=false"]}


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    objs, rebuilt = [], False
    for src in sources():
        obj = os.path.join(OBJDIR, src[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [os.path.join(CSRC, src)] + headers):
            cmd = [_nvcc()] + ARCH + COMMON + EXTRA.get(src, []) + ["-c", os.path.join(CSRC, src), "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
            rebuilt = True
    if rebuilt or force or _stale(SO, objs):
        cmd = [_nvcc()] + ARCH + ["-shared", "-o", SO] + objs
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    if force or _stale(HOST_SO, [HOST_SRC, os.path.join(ROOT, "include", "mnv_host.h")]):
        cmd = [shutil.which("gcc") or "gcc", "-O3", "-Wall", "-shared", "-fPIC", "-pthread", "-I", os.path.join(ROOT, "include"),
               "-o", HOST_SO, HOST_SRC]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
