#!/bin/bash
# standalone probe of the tensor-core match kernel (scripts/mma_probe.cu); the binary travels to the GPU box in build/
set -e
cd "$(dirname "$0")/.."
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o build/mma_probe scripts/mma_probe.cu
