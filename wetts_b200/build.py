"""Builds wetts_b200/csrc/libwetts_b200.so in-tree with nvcc for sm_100a.

    python -m wetts_b200.build [--force]

nvcc cross-compiles without a GPU.  The shared library is git-ignored but travels to the
GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libwetts_b200.so")
SOURCES = ["engine.cu", "conv_kernels.cu", "misc_kernels.cu", "tc_conv_kernel.cu", "tc16_conv_kernel.cu", "tc16p_conv.cu", "tc16r_conv.cu", "fused_rb.cu", "fused_mrf16.cu", "attn_tc.cu"]
HEADERS = ["kernels.cuh", "conv_args.h", "tc_epilogue.cuh", "epilogue.cuh", "tc_prims.cuh", "fused_rb_args.h", "fused_rb_kernel.cuh", "fused_mrf16_args.h", "fused_mrf16_kernel.cuh", "attn_tc_kernel.cuh", "tc16p_conv_kernel.cuh", "tc16r_conv_kernel.cuh", os.path.join("..", "..", "include", "wetts_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    link = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    return LIB


RUNTIME = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "runtime")
RUNTIME_BIN = os.path.join(RUNTIME, "vits_main")


def build_runtime(force=False):
    """C++ `VitsModel` shim + CLI (runtime/): host code only (g++), linked against libwetts_b200.so and cudart."""
    srcs = [os.path.join(RUNTIME, f) for f in ("vits_model.cc", "vits_main.cc")]
    deps = srcs + [os.path.join(RUNTIME, "vits_model.h"), LIB]
    if not force and os.path.exists(RUNTIME_BIN) and all(os.path.getmtime(d) <= os.path.getmtime(RUNTIME_BIN) for d in deps):
        return RUNTIME_BIN
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    inc = os.path.join(os.path.dirname(RUNTIME), "include")
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-I", inc, "-I", RUNTIME, "-I", os.path.join(cuda, "include")] + srcs + [
        "-o", RUNTIME_BIN, "-L", CSRC, "-lwetts_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart",
        "-Wl,-rpath," + CSRC, "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout)
        raise RuntimeError("runtime build failed")
    return RUNTIME_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_runtime(force="--force" in sys.argv))
