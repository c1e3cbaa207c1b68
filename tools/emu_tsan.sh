#!/usr/bin/env bash
# ThreadSanitizer over the engine's kernels on the CPU emulator: the boundary client (C++) linked against csrc/engine.cu compiled
# for the CPU with -fsanitize=thread, blocks of every launch on several OS threads (NC_EMU_THREADS).  Reports conflicting
# non-atomic accesses of different BLOCKS to the same memory (threads of one block are fibres of one OS thread: not seen).
# usage: tools/emu_tsan.sh [preset=standard] [seed=1] [steps=60]      (offline tool; needs g++ with libtsan)
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
B=$ROOT/tests/native/_build
python "$ROOT/tests/emu_build.py" > /dev/null      # generates $B/engine_emu.cpp
g++ -O1 -g -std=c++17 -ffp-contract=off -mfma -fsanitize=thread -pthread -w -fpermissive \
    -I"$ROOT/tests/native" -I"$ROOT/neurocorrelation_b200/csrc" -I"$ROOT/neurocorrelation_b200/host" -I"$ROOT" \
    "$B/engine_emu.cpp" "$ROOT/neurocorrelation_b200/host/NeuCor.cpp" "$ROOT/neurocorrelation_b200/host/checkpoint.cpp" \
    "$ROOT/tests/native/client_presets.cpp" -o "$B/client_tsan" -ldl
NC_EMU_THREADS=${NC_EMU_THREADS:-4} NC_EMU_SMS=${NC_EMU_SMS:-4} TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=4" \
    "$B/client_tsan" "${1:-standard}" "${2:-1}" "${3:-60}"
