#!/bin/bash
# A/B of the builds named on the command line (ab/*.so), twice each, bench tiles, device-resident
set -u
mkdir -p gpurun_out
T=$1; shift
timeout 1200 python tools/gpu_lib_variants.py "$@" "$@" 2>&1 | grep -v "^$" | tail -12 | tee gpurun_out/${T}_variants.txt
