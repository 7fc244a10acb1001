"""Build libetpnav_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m etpnav_b200.build [--force]

Each .cu is compiled to an object file under etpnav_b200/csrc/build/ (skipped when up to date) and
linked into etpnav_b200/libetpnav_b200.so with a static cudart, so the library has no link-time
dependency on libcuda.so and can be dlopen'ed on a box without a driver.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libetpnav_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "etpnav_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(BUILD, exist_ok=True)
    hm = _headers_mtime()
    jobs = []
    objs = []
    for src in _sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hm):
            jobs.append([NVCC, *FLAGS, "-c", s, "-o", o])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB):
        run([NVCC, "-shared", "-o", LIB, *objs, "-cudart", "static", "-Xlinker", "--no-undefined"])
    build_host_helper(force, verbose)
    return LIB


PY_HELPER_SRC = os.path.join(HERE, "csrc_py", "gmap_mirror.c")
PY_HELPER = os.path.join(HERE, "_gmap_mirror.so")


def build_host_helper(force: bool = False, verbose: bool = True) -> str:
    """etpnav_b200/_gmap_mirror.so: the map packer's host half against the CPython API (csrc_py/gmap_mirror.c), loaded with
    ctypes.PyDLL.  Plain gcc; the packer falls back to its pure-Python twin when this file is absent."""
    import sysconfig
    if force or not os.path.exists(PY_HELPER) or os.path.getmtime(PY_HELPER) < os.path.getmtime(PY_HELPER_SRC):
        cmd = [os.environ.get("CC", "gcc"), "-O2", "-Wall", "-shared", "-fPIC", "-I" + sysconfig.get_paths()["include"],
               PY_HELPER_SRC, "-o", PY_HELPER]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("gcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return PY_HELPER


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
