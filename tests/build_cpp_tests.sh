#!/bin/bash
# Builds tests/cpp/*.cc against the drop-in header and libraries (g++ + the CUDA runtime);
# the binaries land in tests/_build/ and travel to the GPU box with the snapshot.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
CUDA="${CUDA_HOME:-/usr/local/cuda}"
mkdir -p "$HERE/_build"
g++ -std=c++17 -O1 -I "$ROOT/include" -I "$CUDA/include" "$HERE/cpp/batched_dropin_test.cc" \
    -o "$HERE/_build/batched_dropin_test" -L "$ROOT/longtermplanner_b200/lib" -llong_term_planner -lltp_b200 \
    -L "$CUDA/lib64" -lcudart -Wl,-rpath,'$ORIGIN/../../longtermplanner_b200/lib' -Wl,-rpath,"$CUDA/lib64"
g++ -std=c++17 -O2 -I "$ROOT/include" "$HERE/cpp/single_plan_bench.cc" \
    -o "$HERE/_build/single_plan_bench" -L "$ROOT/longtermplanner_b200/lib" -llong_term_planner -lltp_b200 \
    -Wl,-rpath,'$ORIGIN/../../longtermplanner_b200/lib'
echo "built $HERE/_build/batched_dropin_test $HERE/_build/single_plan_bench"
