#!/bin/bash
# developer shortcut: rebuild csrc/libvbfem.so with ptxas statistics (the official build is __graft_entry__.build())
cd "$(dirname "$0")/../variational-bayesian-inference-for-computational-mechanics_b200/csrc" || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -diag-suppress 550 -Xptxas -v -I ../../include -o libvbfem.so vbfem.cu 2>&1 | grep -E "error|warning|fem_front|spill|Used" | grep -A2 -E "error|fem_front" | head -${1:-12}
