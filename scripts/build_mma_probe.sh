#!/bin/bash
# standalone probes of the tensor-core match kernels; the binaries travel to the GPU box in build/ (git-ignored)
#   mma_probe        bit-exactness of every match-kernel variant on 14 shapes (PROBE_VARIANT=0 IMAD epilogue, 2 keys from the MMA,
#                    3 CTA pairs) and their times against the integer-pipe kernel:  build/mma_probe check | time P N [pool]
#   mma_probe_prof   the same with -DUZ_MMA_PROF: clocks the MMA issuer spends waiting for tiles / accumulators
#   mma2_probe       issue rate of tcgen05.mma kind::i8, cta_group::1 and ::2, operands resident and fed by a live load pipeline
#   tmem_pack_probe  what tcgen05.ld ... .pack::16b returns
#   tmem_probe, pipe_probe, pcie_probe: TMEM read bandwidth, integer-pipe rates, PCIe gather rates
set -e
cd "$(dirname "$0")/.."
mkdir -p build
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo"
nvcc $F -o build/mma_probe scripts/mma_probe.cu
nvcc $F -DUZ_MMA_PROF=1 -o build/mma_probe_prof scripts/mma_probe.cu
nvcc $F -o build/mma2_probe scripts/mma2_probe.cu
nvcc $F -o build/tmem_pack_probe scripts/tmem_pack_probe.cu
nvcc $F -o build/tmem_probe scripts/tmem_probe.cu
nvcc $F -o build/pipe_probe scripts/pipe_probe.cu
nvcc $F -o build/pcie_probe scripts/pcie_probe.cu
