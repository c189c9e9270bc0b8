#!/bin/sh
# Experiment builds of the library next to the product build (never loaded unless CNEUS_LIB points at them):
#   sh tools/build_variant.sh myexp -DMY_EXPERIMENT_FLAG   ->  tools/libcneus_myexp.so
#   CNEUS_LIB=$PWD/tools/libcneus_myexp.so python bench.py --extras none --no-cpu-baseline
#   sh tools/build_variant.sh single -DCNEUS_TC_SINGLE       (the one-CTA kernel, for A/B against the CTA-pair product build)
#   sh tools/build_variant.sh shfl -DCNEUS_TC_SHFL_CONSTS     (narrow-layer rows through shuffles instead of shared memory)
# (round 2 used it for the single-accumulator A/B: the product build against a -DCNEUS_TC_SINGLE_ACC variant of the then
#  two-accumulator kernel, profiles/r2b_parity_errors*.json; that scheme is the product path now)
name="$1"; shift
cd "$(dirname "$0")/.." || exit 1
exec /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC \
  "$@" -Iinclude -Icolor_neus_b200/csrc -o "tools/libcneus_${name}.so" color_neus_b200/csrc/*.cu
