"""In-tree builds of the native libraries.

  lib/libh264r.so      CUDA engine + C ABI (include/h264r.h); nvcc, sm_100a only
  lib/libh264host.so   host side: stream synthesizer + syntax parser (include/h264_syntax.h); g++
  oracle/_ref/*        test oracles (built by oracle/Makefile; not part of the product)

The .so files are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "videodecoder_b200")
LIB = os.path.join(PKG, "lib")
CSRC = os.path.join(PKG, "csrc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--use_fast_math", "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, cwd=None, log=None):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout


def _glob(d, exts):
    out = []
    for base, _, files in os.walk(d):
        for f in files:
            if f.endswith(exts):
                out.append(os.path.join(base, f))
    return sorted(out)


def build_host(force=False):
    os.makedirs(LIB, exist_ok=True)
    target = os.path.join(LIB, "libh264host.so")
    srcs = _glob(os.path.join(CSRC, "host"), (".cpp",))
    deps = srcs + _glob(os.path.join(CSRC, "host"), (".h",)) + _glob(os.path.join(ROOT, "include"), (".h",))
    if force or _newer(target, deps):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I", os.path.join(ROOT, "include")]
             + srcs + ["-o", target])
    return target


def build_cuda(force=False):
    os.makedirs(LIB, exist_ok=True)
    target = os.path.join(LIB, "libh264r.so")
    srcs = _glob(os.path.join(CSRC, "cuda"), (".cu",))
    deps = srcs + _glob(os.path.join(CSRC, "cuda"), (".cuh", ".h")) + _glob(os.path.join(ROOT, "include"), (".h",))
    if not srcs:
        raise RuntimeError("no CUDA sources")
    if force or _newer(target, deps):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        if not os.path.exists(nvcc):
            if os.path.exists(target):
                return target          # GPU box without a toolkit: use the prebuilt library
            raise RuntimeError("nvcc not found and no prebuilt libh264r.so")
        _run([nvcc] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include")] + srcs + ["-o", target, "-lcudart"],
             log=os.path.join(LIB, "nvcc_build.log"))
    return target


def build_oracle():
    """Builds the test oracles (the checker, never the product)."""
    _run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])


def build_all(force=False):
    build_host(force)
    build_cuda(force)
    build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("ok")
