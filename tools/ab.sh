#!/bin/bash
# A/B timing of library variants built by tools/build_variants.py (run on the GPU box)
mkdir -p gpurun_out
for v in "$@"; do
  LTP_B200_LIB=$PWD/tools/_bin/libltp_$v.so python tools/solve_timing.py 2>&1 | grep -v sampler | sed "s#.*_bin/##"
done
for v in "$@"; do
  LTP_B200_LIB=$PWD/tools/_bin/libltp_$v.so python tools/solve_timing.py 2>&1 | grep -v sampler | sed "s#.*_bin/##"
done
