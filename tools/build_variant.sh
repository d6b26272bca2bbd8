#!/bin/bash
# usage: tools/build_variant.sh NAME -DVV_X=..   -> vvflow_b200/lib/variants/libvvgpu_NAME.so
name=$1; shift
mkdir -p vvflow_b200/lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" -o vvflow_b200/lib/variants/libvvgpu_$name.so vvflow_b200/csrc/vvgpu.cu 2>&1 | grep -v "warning\|kMaxDepth\|\^\|^$" 
