"""Offline (not part of the test-suite: ~10 minutes per fixture): a whole 10 000-step C1 fixture recorded from the reference,
run through the engine's own kernels on the CPU emulator (tests/emu_build.py) — state signature, detector voltage and explicit
spike raster compared at every step, final potentials / weights / lastFire bit-identical.
usage: python tools/emu_long_golden.py c1_long_seed1_normalised.npz [steps|all] [driver_draws]
       (c1_control_h.npz, the horizon control, was recorded without driver draws: pass `all 0`)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emu_build  # noqa: E402
import scenarios  # noqa: E402

lib = emu_build.build()
t0 = time.time()
name = sys.argv[1]
steps = None if len(sys.argv) <= 2 or sys.argv[2] == "all" else int(sys.argv[2])
st, z = scenarios.c1_long_golden(lib, name, steps=steps, driver_draws=int(sys.argv[3]) if len(sys.argv) > 3 else 3)
print(name, "OK", st, "horizon", int(z["horizon"]), "time %.0f s" % (time.time() - t0))
