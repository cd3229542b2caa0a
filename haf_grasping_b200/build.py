"""Builds libhafgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIB_DIR, "libhafgpu.so")
SOURCES = ["hafgpu.cu"]


def _lib_deps():
    """every file libhafgpu.so is compiled from: all of csrc/*.{cu,cuh,hpp} plus the public header"""
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".hpp", ".h"))]
    deps.append(os.path.join(PKG, "..", "include", "hafgpu.h"))
    return deps

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false",                    # no implicit FMA contraction: reference-exact FP32/FP64; FMAs are explicit
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = _lib_deps() + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-ccbin", "/usr/bin/g++"] * int(os.path.exists("/usr/bin/g++")) + \
          ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libhafgpu.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB


CLI = os.path.join(LIB_DIR, "haf_cli")
HOST_DIR = os.path.join(CSRC, "host")


def build_cli(force: bool = False) -> str:
    """haf_cli: the ROS-free C++ host (csrc/host) on top of the C ABI."""
    build_lib()
    srcs = [os.path.join(HOST_DIR, f) for f in ("haf_cli.cpp", "calc_grasppoints_b200.hpp", "pcd_io.hpp")]
    if not force and os.path.exists(CLI) and all(os.path.getmtime(s) <= os.path.getmtime(CLI) for s in srcs + [LIB]):
        return CLI
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-o", CLI, srcs[0], "-L" + LIB_DIR, "-lhafgpu", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building haf_cli")
    return CLI


SVM_PREDICT = os.path.join(LIB_DIR, "svm-predict-b200")
SVM_SCALE = os.path.join(LIB_DIR, "svm-scale-b200")


def build_svm_tools(force: bool = False):
    """svm-predict-b200 / svm-scale-b200: command-line compatible front ends of libsvm's two programs."""
    build_lib()
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    outs = []
    for src, out in (("svm_predict_b200.cpp", SVM_PREDICT), ("svm_scale_b200.cpp", SVM_SCALE)):
        srcs = [os.path.join(HOST_DIR, src), os.path.join(HOST_DIR, "libsvm_text.hpp")]
        if force or not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs + [LIB]):
            cmd = [cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-o", out, srcs[0], "-L" + LIB_DIR, "-lhafgpu", "-Wl,-rpath,$ORIGIN"]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                raise RuntimeError("g++ failed building " + os.path.basename(out))
        outs.append(out)
    return outs


if __name__ == "__main__":
    print(build_lib(force=True, verbose="-v" in sys.argv))
    print(build_cli(force=True))
    print(build_svm_tools(force=True))
