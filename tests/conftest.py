import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a machine without a CUDA device skips the gpu-marked tests instead of failing them."""
    try:
        import ctypes
        from neurocorrelation_b200 import build
        have = os.path.exists(build.ENGINE_SO) and ctypes.CDLL(build.ENGINE_SO).nc_device_count() > 0
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device (the engine has no CPU path); run with -m gpu on a B200")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def native_libs():
    """Builds (in-tree, if stale) the product libraries; returns their paths."""
    from neurocorrelation_b200 import build
    if os.path.exists("/usr/local/cuda/bin/nvcc"):
        return build.build_all()
    return build.ENGINE_SO, build.HOST_SO, build.MATH_SO


@pytest.fixture(scope="session")
def mock_host_lib():
    """Host NeuCor class linked against the single-threaded CPU model of the engine ABI (a TEST DOUBLE that lives
    in tests/native and is never loaded by the product)."""
    out_dir = os.path.join(ROOT, "tests", "native", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libneucor_host_mock.so")
    srcs = [os.path.join(ROOT, "tests", "native", "mock_ncabi.cpp"),
            os.path.join(ROOT, "neurocorrelation_b200", "host", "NeuCor.cpp"),
            os.path.join(ROOT, "neurocorrelation_b200", "host", "capi.cpp"),
            os.path.join(ROOT, "neurocorrelation_b200", "host", "checkpoint.cpp")]
    deps = srcs + [os.path.join(ROOT, "neurocorrelation_b200", "csrc", f) for f in ("step_logic.cuh", "glibc_math.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-mfma", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I" + ROOT] + srcs + ["-o", out])
    return out


@pytest.fixture(scope="session")
def have_ref():
    from oracle import refbind
    return refbind.available("ref") and refbind.available("ref_canon")
