#!/bin/bash
# standalone probes of the tensor-core match kernels; the binaries travel to the GPU box in build/ (git-ignored)
#   mma_probe        bit-exactness of every match-kernel variant on 14 shapes (PROBE_VARIANT=0 IMAD epilogue, 2 keys from the MMA,
#                    3 CTA pairs, 5 four-bit operands) and their times against the integer-pipe kernel:
#                    build/mma_probe check | time P N [pool]
#   mma_probe_prof   the same with -DUZ_MMA_PROF: clocks the MMA issuer spends waiting for tiles / accumulators
#   mma_probe_trace  -DUZ_F4_TRACE: clock stamps of the issuer(s) and one epilogue warp per accumulator of knn2_mmaf_kernel
#                    (-DF4_ISSUERS=1, -DF4_EPI_GROUPS=1 build the measured alternatives of that kernel)
#   mxf4_probe       tcgen05.mma kind::mxf4.block_scale: exactness of +-4 x +-4 x 2 x 2 sums on an fp32 accumulator that starts
#                    at 2^23 + 16384 + 127 - column (set by one kind::f8f6f4 instruction), and clocks per instruction
#   minmax_pipe_probe  packed 16-bit min/max as u16x2 / f16x2 / bf16x2 alone and mixed: one pipe, 2 clocks each
#   mma2_probe       issue rate of tcgen05.mma kind::i8, cta_group::1 and ::2, operands resident and fed by a live load pipeline
#   tmem_pack_probe  what tcgen05.ld ... .pack::16b returns
#   tmem_probe, pipe_probe, pcie_probe: TMEM read bandwidth, integer-pipe rates, PCIe gather rates
set -e
cd "$(dirname "$0")/.."
mkdir -p build
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo"
nvcc $F -o build/mma_probe scripts/mma_probe.cu
nvcc $F -DUZ_MMA_PROF=1 -o build/mma_probe_prof scripts/mma_probe.cu
nvcc $F -DUZ_F4_TRACE=1 -o build/mma_probe_trace scripts/mma_probe.cu
nvcc $F -o build/mxf4_probe scripts/mxf4_probe.cu
nvcc $F -o build/minmax_pipe_probe scripts/minmax_pipe_probe.cu
nvcc $F -o build/mma2_probe scripts/mma2_probe.cu
nvcc $F -o build/tmem_pack_probe scripts/tmem_pack_probe.cu
nvcc $F -o build/tmem_probe scripts/tmem_probe.cu
nvcc $F -o build/pipe_probe scripts/pipe_probe.cu
nvcc $F -o build/pcie_probe scripts/pcie_probe.cu
