#!/bin/bash
# A/B of the time-major sampler (configs[2] shape) between library variants
for r in 1 2; do for v in "$@"; do
  LTP_B200_LIB=$PWD/tools/_bin/libltp_$v.so python tools/solve_timing.py 2>&1 | grep sampler | sed "s#^#$v #"
done; done
