#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY — builds the reference's own CPU implementation of the hot path
# (NeuCor::run and everything it dispatches to, /root/reference/src/NeuCor.{h,cpp}) into
# oracle/_ref/, from the sources where they lie.  No reference source is copied into this repo:
# the tie-canonicalised variant is produced from a throw-away copy in a temp dir that is deleted
# again; only the two shared objects land in oracle/_ref/ (git-ignored, NOT gpurun-ignored).
#
# Flags are the reference's own (CMakeLists.txt:22-25): -O3, C++17, default arch, no fast-math.
#   libneucor_ref.so        unmodified NeuCor.cpp + oracle/ref_harness.cpp
#   libneucor_ref_canon.so  same, with simulation::operator> (NeuCor.h:145) extended to break
#                           equal-time ties by (rank, a, b): InputFirer (0, index, 0) <
#                           Synapse (1, target, parent) < Neuron (2, id, 0)  — SURVEY.md App. C.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${NEUCOR_REFERENCE:-/root/reference}/src"
OUT="$HERE/_ref"
if [ ! -f "$REF/NeuCor.cpp" ]; then
    echo "build_ref: $REF/NeuCor.cpp not present; keeping prebuilt oracle/_ref" >&2
    exit 0
fi
mkdir -p "$OUT"
CXXFLAGS="-O3 -std=c++17 -fPIC -shared -Wl,-Bsymbolic"

g++ $CXXFLAGS -I"$REF" "$REF/NeuCor.cpp" "$HERE/ref_harness.cpp" -o "$OUT/libneucor_ref.so"

TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
# 1) comparator becomes a declaration; 2) let it read Synapse::pN/tN through a friend declaration
sed -e 's|bool operator>(const simulation &otherSim) const {return stime > otherSim.stime;};|bool operator>(const simulation \&otherSim) const;|' \
    -e 's|friend class NeuCor_Renderer;|friend class NeuCor_Renderer; friend struct simulation;|' \
    "$REF/NeuCor.h" > "$TMP/NeuCor.h"
grep -q 'bool operator>(const simulation &otherSim) const;' "$TMP/NeuCor.h" || { echo "build_ref: comparator patch did not apply" >&2; exit 1; }
cp "$REF/NeuCor.cpp" "$TMP/NeuCor.cpp"
cat "$HERE/canon_comparator.inc" >> "$TMP/NeuCor.cpp"
g++ $CXXFLAGS -I"$TMP" "$TMP/NeuCor.cpp" "$HERE/ref_harness.cpp" -o "$OUT/libneucor_ref_canon.so"
echo "build_ref: built $OUT/libneucor_ref.so and $OUT/libneucor_ref_canon.so"
