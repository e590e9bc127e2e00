#!/bin/bash
# Compiles the reference's OWN gtest source (read where it lies under /root/reference, never
# copied) against this repository's drop-in header and libraries, with tests/gtest_shim in
# place of GoogleTest. Only possible where /root/reference exists (the build container); the
# resulting binary tests/_build/ref_tests_rehosted is shipped to the GPU box by gpurun.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${REF:-/root/reference}"
[ -f "$REF/tests/src/long_term_planner_tests.cc" ] || { echo "no reference tree at $REF - keeping prebuilt binary"; exit 0; }
mkdir -p "$HERE/_build"
g++ -std=c++17 -O1 -I "$HERE/gtest_shim" -I "$ROOT/include" -I "$REF/tests/include" \
    "$REF/tests/src/long_term_planner_tests.cc" -o "$HERE/_build/ref_tests_rehosted" \
    -L "$ROOT/longtermplanner_b200/lib" -llong_term_planner -lltp_b200 \
    -Wl,-rpath,'$ORIGIN/../../longtermplanner_b200/lib'
echo "built $HERE/_build/ref_tests_rehosted"
