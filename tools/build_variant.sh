#!/bin/bash
# tools/build_variant.sh NAME [extra nvcc flags...]  -> gpurun_out/variants/libsbx_NAME.so (kernel tuning experiments)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off -shared "$@" \
  -o variants/libsbx_$name.so spectrogram-rs_b200/csrc/sbx_plan.cu spectrogram-rs_b200/csrc/sbx_live.cu
echo variants/libsbx_$name.so
