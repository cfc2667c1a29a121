"""Builds ctcdecode_b200/_lib/libctcdecode_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python -m ctcdecode_b200.build [--force]

The library is a plain C-ABI shared object (include/ctcdecode_b200.h): CUDA runtime linked statically,
libstdc++ dynamically, no torch / pybind dependency.
"""
import os
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
_OUT_DIR = os.path.join(_PKG, "_lib")
LIB_PATH = os.path.join(_OUT_DIR, "libctcdecode_b200.so")
_NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "1886"]


def _sources():
    inc = os.path.join(_PKG, "..", "include", "ctcdecode_b200.h")
    return [os.path.join(_CSRC, f) for f in sorted(os.listdir(_CSRC))] + [inc]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(_OUT_DIR, exist_ok=True)
    obj = os.path.join(_OUT_DIR, "ctc_api.%d.o" % os.getpid())
    cmd = [_NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-c", os.path.join(_CSRC, "ctc_api.cu"), "-o", obj]
    subprocess.run(cmd, check=True)
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(_NVCC)), "lib64")
    # link with the host compiler so that libstdc++ stays a shared dependency (this image's g++ would
    # otherwise pull in a static libstdc++ that clashes with the one python already loaded)
    tmp = LIB_PATH + ".tmp%d" % os.getpid()  # link beside the target and rename: readers never see a partial file
    link = ["g++", "-shared", "-nostdlib++", "-o", tmp, obj, "-L" + cuda_lib, "-lcudart_static",
            "-l:libstdc++.so.6", "-lm", "-lrt", "-lpthread", "-ldl"]
    subprocess.run(link, check=True)
    os.replace(tmp, LIB_PATH)
    os.remove(obj)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
