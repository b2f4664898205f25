"""Build the native libraries in-tree (outputs are git-ignored but travel with gpurun snapshots).

  cnosdb_b200/libtskv_gpu.so      CUDA kernels + C ABI (include/tskv_gpu.h), sm_100a only
  cnosdb_b200/libtskv_hostgen.so  host-side TSM page writer + synthetic data generator
(The CPU checker under oracle/ is test infrastructure with its own Makefile; __graft_entry__.build()
and tests/conftest.py build it, this package never does.)
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cnosdb_b200")
CSRC = os.path.join(PKG, "csrc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, cwd=None):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd, cwd=cwd)


def _glob_sources(*dirs, exts=(".cu", ".cuh", ".cc", ".h")):
    out = []
    for d in dirs:
        for f in sorted(os.listdir(d)):
            if f.endswith(exts):
                out.append(os.path.join(d, f))
    return out


def build_gpu(force=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    target = os.path.join(PKG, "libtskv_gpu.so")
    deps = _glob_sources(CSRC) + [os.path.join(ROOT, "include", "tskv_gpu.h")]
    if force or _newer(target, deps):
        _run([nvcc] + NVCC_FLAGS + ["-o", target, os.path.join(CSRC, "tskv_gpu.cu"),
                                    os.path.join(CSRC, "host_util.cc"), "-ldl"])  # libnccl is dlopen-ed on first use
    return target


def build_hostgen(force=False):
    target = os.path.join(PKG, "libtskv_hostgen.so")
    host = os.path.join(CSRC, "host")
    deps = _glob_sources(host) + [os.path.join(CSRC, "host_util.cc"), os.path.join(CSRC, "host_util.h")]
    if force or _newer(target, deps):
        _run(["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wall", "-o", target,
              os.path.join(host, "tsm_writer.cc"), os.path.join(host, "datagen.cc"), os.path.join(host, "tsm_file.cc"),
              os.path.join(CSRC, "host_util.cc"), "-lz"])
    return target


def build_all(force=False):
    return build_gpu(force), build_hostgen(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
