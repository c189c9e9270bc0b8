#!/bin/sh
# Experiment builds of the library next to the product build (never loaded unless CNEUS_LIB points at them):
#   sh tools/build_variant.sh single_acc -DCNEUS_TC_SINGLE_ACC   ->  tools/libcneus_single_acc.so
#   CNEUS_LIB=$PWD/tools/libcneus_single_acc.so python -m pytest tests -m gpu ...
name="$1"; shift
cd "$(dirname "$0")/.." || exit 1
exec /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC \
  "$@" -Iinclude -Icolor_neus_b200/csrc -o "tools/libcneus_${name}.so" color_neus_b200/csrc/*.cu
