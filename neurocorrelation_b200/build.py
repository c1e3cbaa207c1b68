"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo snapshot).

  csrc/libneucor_b200.so   CUDA engine + C ABI (include/neucor_b200.h), sm_100a only
  host/libneucor_host.so   host-side NeuCor class + flat C wrapper, linked against the engine
  csrc/libnc_mathhost.so   host build of the device math replicas, for CPU-side verification vs libm
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
HOST = os.path.join(ROOT, "host")
ENGINE_SO = os.path.join(CSRC, "libneucor_b200.so")
HOST_SO = os.path.join(HOST, "libneucor_host.so")
MATH_SO = os.path.join(CSRC, "libnc_mathhost.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false",  # no FMA contraction anywhere: the reference's float/double typing is kept operator by operator
              "-Xcompiler", "-fPIC", "-shared"]
HOST_FLAGS = ["-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-pthread"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd[:3]))
    return r


def build_engine(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in ("engine.cu", "step_logic.cuh", "glibc_math.cuh", "glibc_tables.h", "rand_stream.cuh", "peer_exchange.cuh")]
    srcs.append(os.path.join(ROOT, "..", "include", "neucor_b200.h"))
    if force or _newer(ENGINE_SO, srcs):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", ENGINE_SO, srcs[0], "-ldl"]
        r = _run(cmd)
        if verbose:
            print(r.stderr)
    return ENGINE_SO


def build_host(force=False):
    srcs = [os.path.join(HOST, f) for f in ("NeuCor.cpp", "capi.cpp", "checkpoint.cpp", "NeuCor.h")]
    if force or _newer(HOST_SO, srcs + [ENGINE_SO]):
        _run(["g++"] + HOST_FLAGS + srcs[:3] + ["-o", HOST_SO, "-L" + CSRC, "-lneucor_b200", "-Wl,-rpath,$ORIGIN/../csrc"])
    return HOST_SO


def build_mathhost(force=False):
    srcs = [os.path.join(CSRC, f) for f in ("glibc_math_host.cpp", "glibc_math.cuh", "glibc_tables.h")]
    if force or _newer(MATH_SO, srcs):
        _run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-fPIC", "-shared", srcs[0], "-o", MATH_SO])
    return MATH_SO


def build_all(force=False, verbose=False):
    build_engine(force, verbose)
    build_host(force)
    build_mathhost(force)
    return ENGINE_SO, HOST_SO, MATH_SO


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
