#!/bin/bash
# build a variant of the CUDA library into gpurun_ab/lib_<name>.so:  tools/build_variant.sh <name> [-DFLAG=..]...
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --cudart=shared \
  -Xlinker -rpath,/usr/local/cuda/lib64 "$@" -shared -o gpurun_ab/lib_$name.so cuda-grmonty_b200/csrc/gm_api.cu -ldl 2>&1 | grep -v "warning #177\|SD = R\|^ *\^\|^$\|Remark"
