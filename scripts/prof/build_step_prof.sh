#!/bin/bash
# builds build/step_prof from the library sources with in-kernel timestamps enabled
set -e
cd "$(dirname "$0")/../.."
mkdir -p build
C=tacorl_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DTACORL_STEP_PROFILE $EXTRA -o build/step_prof$SUFFIX \
  scripts/prof/step_prof.cu $C/api.cu $C/gemm_f32.cu -lcuda
