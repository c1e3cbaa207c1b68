"""The drop-in boundary at source level: ONE client translation unit written against the reference's
NeuCor class surface (tests/native/client_presets.cpp: headless restatements of main.cpp's STANDARD / ONE_INPUT /
FEW_NEURONS presets plus the members NeuCor_Renderer reads through friendship) is compiled against the reference's own
NeuCor.h + NeuCor.cpp and against neurocorrelation_b200/host/NeuCor.h, and both binaries are run in lock-step: same lines.
  * CPU (not gpu): reference build (when /root/reference is present) vs the drop-in linked against the CPU test double;
    the reference's output is also kept as tests/golden/client_presets.txt.
  * GPU (-m gpu): the drop-in linked against the real libraries, against that committed output — in tests/test_zy_client_gpu.py,
    so that it runs after the parity tests proper (its renderer-frame walk was added after the round's GPU budget was spent)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NATIVE = os.path.join(ROOT, "tests", "native")
BUILD = os.path.join(NATIVE, "_build")
HOST = os.path.join(ROOT, "neurocorrelation_b200", "host")
GOLDEN = os.path.join(ROOT, "tests", "golden", "client_presets.txt")
REF = "/root/reference/src"
RUNS = [("standard", "1", "400"), ("standard", "4", "300"), ("one_input", "2", "400"), ("few_neurons", "3", "8000")]


def _run_all(exe, env=None):
    out = []
    for args in RUNS:
        r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=900, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        out.append(r.stdout)
    return "".join(out)


def _build_reference_client():
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "client_ref")
    subprocess.check_call(["g++", "-O3", "-std=c++17", "-DCLIENT_REFERENCE_BUILD", "-I" + REF, os.path.join(NATIVE, "client_presets.cpp"),
                           os.path.join(REF, "NeuCor.cpp"), os.path.join(NATIVE, "zero_heap.cpp"), "-o", exe])
    return exe


def _build_dropin_client(mock):
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "client_mock" if mock else "client_b200")
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-pthread", "-I" + HOST, "-I" + ROOT, os.path.join(NATIVE, "client_presets.cpp")]
    if mock:
        cmd += ["-mfma", os.path.join(HOST, "NeuCor.cpp"), os.path.join(HOST, "checkpoint.cpp"), os.path.join(NATIVE, "mock_ncabi.cpp")]
    else:
        cmd += ["-L" + HOST, "-lneucor_host", "-Wl,-rpath," + HOST, "-Wl,-rpath," + os.path.join(ROOT, "neurocorrelation_b200", "csrc")]
    subprocess.check_call(cmd + ["-o", exe])
    return exe


def test_client_compiles_against_both_headers_and_runs_in_lockstep():
    got = _run_all(_build_dropin_client(mock=True))
    assert "DIFFERENT" not in got
    assert "w(0->1) > 0.9, w(0->2) < 0.1" in got  # the essay's known answer (section 2.5.1), FEW_NEURONS
    if os.path.exists(os.path.join(REF, "NeuCor.cpp")):
        want = _run_all(_build_reference_client())
        if not os.path.exists(GOLDEN) or open(GOLDEN).read() != want:
            open(GOLDEN, "w").write(want)
        assert got == want, "the drop-in and the reference print different lines"
    assert got == open(GOLDEN).read()


def test_client_against_the_emulated_engine():
    """The same client linked against the engine's own kernels on the CPU emulator (tests/emu_build.py): the renderer-frame walk
    over the object graph — every synapse's end potentials from k_synapse_pots, weights, potentials, activities, the raster
    rule — detector reads, input toggles and resetActivities print the same lines as the build on the CPU test double, which
    the test above holds to the reference's own output.  A shortened first preset (the emulator is slow)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emu_build
    lib = emu_build.build()
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "client_emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-pthread", "-I" + HOST, "-I" + ROOT, os.path.join(NATIVE, "client_presets.cpp"),
                           lib, "-Wl,-rpath," + os.path.dirname(lib), "-o", exe])
    run = ("standard", "1", "150")
    got = subprocess.run([exe, *run], capture_output=True, text=True, timeout=900)
    assert got.returncode == 0, got.stderr[-2000:]
    want = subprocess.run([_build_dropin_client(mock=True), *run], capture_output=True, text=True, timeout=900)
    assert want.returncode == 0, want.stderr[-2000:]
    assert got.stdout.count("\n") >= 4 and "DIFFERENT" not in got.stdout and "carrying a spike" in got.stdout
    assert got.stdout == want.stdout, "the drop-in prints different lines on the emulated engine and on the test double"
