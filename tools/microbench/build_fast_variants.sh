#!/bin/bash
# builds tools/microbench/fast_variants from the product's kernel header
cd "$(dirname "$0")/../.." && nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -I ofdm-based-systems_b200/csrc -I include tools/microbench/fast_variants.cu -o tools/microbench/fast_variants "$@"
