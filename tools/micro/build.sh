#!/bin/bash
# Builds the stand-alone micro-benchmarks (not part of libwm_b200).
set -eo pipefail
cd "$(dirname "$0")"
mkdir -p _bin
for f in *.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I../../world_modelz_b200/csrc -o _bin/${f%.cu} $f -lcuda
done
