"""In-tree build of libcp360.so (nvcc, sm_100a only). Used by __graft_entry__.build().

The shared object is written next to the package (lib/libcp360.so): it is git-ignored but
travels to the GPU box with the repo snapshot. There is no JIT and no other architecture.
"""
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.environ.get("CP360_LIB") or os.path.join(LIB_DIR, "libcp360.so")   # CP360_LIB: dev override
OBJ_DIR = os.path.join(PKG_DIR, "build")

CU_SOURCES = ["common.cu", "cubepad.cu", "e2c.cu", "c2e.cu", "hostmem.cu"]
CPP_SOURCES = ["maps.cpp", "npy.cpp"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                     "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]
# map builders must follow numpy's float64 operation order literally: no FMA contraction
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fvisibility=hidden"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libcp360.so cannot be built on this machine")
    return exe


def _newest_source_mtime():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    paths += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    paths.append(os.path.abspath(__file__))
    return max(os.path.getmtime(p) for p in paths)


def is_fresh():
    return os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_source_mtime()


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build step failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu + maps.cpp for sm_100a and link lib/libcp360.so. Returns its path."""
    if not force and is_fresh():
        return LIB_PATH
    nvcc = _nvcc()
    extra = os.environ.get("CP360_NVCC_EXTRA", "").split()
    obj_dir = OBJ_DIR if not extra else OBJ_DIR + "_" + os.path.splitext(os.path.basename(LIB_PATH))[0]
    os.makedirs(obj_dir, exist_ok=True)
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    jobs = []
    for src in CU_SOURCES:
        obj = os.path.join(obj_dir, src + ".o")
        jobs.append(([nvcc] + NVCC_FLAGS + extra + ["-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj], obj))
    for src in CPP_SOURCES:
        obj = os.path.join(obj_dir, src + ".o")
        jobs.append((["g++"] + CXX_FLAGS + ["-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj], obj))
    with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
        logs = list(ex.map(lambda j: _run(j[0]), jobs))
    if verbose:
        print("".join(logs))
    tmp = LIB_PATH + ".tmp"
    _run([nvcc] + ARCH + ["-shared", "-o", tmp] + [j[1] for j in jobs])
    os.replace(tmp, LIB_PATH)
    return LIB_PATH
