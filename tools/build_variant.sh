#!/bin/bash
# Build a tuning variant of libicpb.so next to the product library (developer tool for A/B runs):
#   bash tools/build_variant.sh <name> [extra nvcc flags, e.g. -DICPB_G=2 -DICPB_MIN_CTAS=4]
# -> icp-slam-with-loop-closure_b200/libicpb_<name>.so   (select with ICPB_SO=<path>)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xcompiler -fPIC -shared "$@" \
     -I include -I icp-slam-with-loop-closure_b200/csrc \
     -o icp-slam-with-loop-closure_b200/libicpb_${name}.so icp-slam-with-loop-closure_b200/csrc/icpb_api.cu
echo built libicpb_${name}.so
