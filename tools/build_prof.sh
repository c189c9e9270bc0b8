#!/bin/sh
# Profiling build of the library (fine-grained epilogue timeline, see EpiProf in mlp_tc_kernel.cu):
#   sh tools/build_prof.sh && CNEUS_LIB=$PWD/tools/libcneus_prof.so python tools/tc_role_prof.py
cd "$(dirname "$0")/.." || exit 1
exec /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC \
  -DCNEUS_TC_EPI_PROF -Iinclude -Icolor_neus_b200/csrc -o tools/libcneus_prof.so color_neus_b200/csrc/*.cu
