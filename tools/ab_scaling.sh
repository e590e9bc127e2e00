#!/bin/bash
# tools/ab_scaling.sh "n1 n2 ..." variant...   (time-major sampler GB/s per variant)
ns=$1; shift
for v in "$@"; do echo "== $v"; LTP_B200_LIB=$PWD/tools/_bin/libltp_$v.so python tools/sampler_scaling_probe.py $ns 2>&1 | grep "n ="; done
