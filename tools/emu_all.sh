#!/usr/bin/env bash
# Everything that checks the kernels' source WITHOUT a GPU, in one place (each line can be run on its own):
set -eu
cd "$(dirname "$0")/.."
python -m pytest tests/test_engine_emulated.py tests/test_kernel_emulation.py tests/test_boundary_client.py tests/test_sharded.py -q -m "not gpu"
NC_EMU_ORDER=1 python -m pytest tests/test_engine_emulated.py -q            # threads and blocks resumed in descending order
NC_EMU_ORDER=2 python -m pytest tests/test_engine_emulated.py -q            # ... in random order
NC_EMU_THREADS=4 NC_EMU_SMS=4 python -m pytest tests/test_engine_emulated.py -q   # blocks of a launch on four OS threads
python tools/emu_fuzz.py 1 60                                                # random configurations against the oracle
python tools/emu_fuzz_sharded.py 1 6                                         # world 2 / 3 over gloo
python tools/emu_long_golden.py c1_long_seed1_normalised.npz                 # a whole 10 000-step fixture (~10 min)
tools/emu_tsan.sh standard 1 60                                              # ThreadSanitizer over concurrently running blocks
python tools/sass_diff.py "${1:-neurocorrelation_b200/csrc/libneucor_b200.so}" neurocorrelation_b200/csrc/libneucor_b200.so   # which kernels differ from a validated build
