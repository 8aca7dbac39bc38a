#!/bin/bash
# Build a kernel variant for A/B timing: scripts/build_var.sh <name> "<extra nvcc flags>"  ->  gst_b200/lib/var/<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p gst_b200/lib/var
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -shared -x cu \
  $2 -I include -o gst_b200/lib/var/$1.so gst_b200/csrc/gst_kernels.cu gst_b200/csrc/gst_capi.cu
echo built gst_b200/lib/var/$1.so
