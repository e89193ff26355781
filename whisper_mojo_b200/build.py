"""Build whisper_mojo_b200/libwhisper_b200.so in-tree with nvcc for sm_100a.

    python -m whisper_mojo_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the
snapshot.  No torch / pybind dependency: the library is a plain C-ABI shared object.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libwhisper_b200.so")  # fp16 operands (default; csrc/dtype.h)
LIB_BF16 = os.path.join(HERE, "libwhisper_b200_bf16.so")  # -DWB_BF16: the bf16 variant, for A/B numbers
SOURCES = ["api.cu", "ops.cu", "gemm.cu", "kernels.cu", "frontend.cu", "frontend_tc.cu", "model.cu", "attn_tc.cu", "cross_attn_tc.cu", "decode_chain.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler",
              "-fPIC,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _host_compiler_flags():
    # the image's default CC/CXX wrappers lack some specs; prefer the distro g++ when present
    return ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(force: bool = False, verbose: bool = False, bf16: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "whisper_b200.h")]
    LIB = LIB_BF16 if bf16 else globals()["LIB"]
    OBJ = globals()["OBJ"] + ("_bf16" if bf16 else "")
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest(deps):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        cmd = [nvcc] + _host_compiler_flags() + NVCC_FLAGS + (["-DWB_BF16"] if bf16 else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, (r.stdout + r.stderr)[-6000:]))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc] + _host_compiler_flags() + ["-shared", "-o", LIB] + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


def build_example(force: bool = False) -> str:
    """examples/main.cpp (the reference's main.mojo over the C ABI, plain C++) -> examples/main_b200."""
    root = os.path.dirname(HERE)
    src, out = os.path.join(root, "examples", "main.cpp"), os.path.join(root, "examples", "main_b200")
    hdr = os.path.join(root, "include", "whisper_b200.h")
    if not force and os.path.exists(out) and os.path.getmtime(out) >= _newest([src, hdr, build()]):
        return out
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(root, "include"), src, "-L" + HERE, "-lwhisper_b200",
           "-Wl,-rpath,$ORIGIN/../whisper_mojo_b200", "-Wl,--allow-shlib-undefined", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("example build failed:\n" + r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, bf16=True))
    print(build_example(force="--force" in sys.argv))
