"""Instruction mix of the match kernels per descriptor compare, as the roofline arithmetic of bench.py uses it.
tests/test_sass.py disassembles the shipped library (cuobjdump) and checks these numbers against the hot loops, so a
compiler change cannot move the real ceiling silently."""

# knn2_kernel<256, 2, CSA, PACK16>: integer pipes, per 256-bit compare
KNN2_POPC = 4          # XU pipe
KNN2_LOP3 = 13         # ALU pipe
KNN2_IMAD = 4          # FMA pipe (weighted popcount chain); +1 per iteration of 8 compares for the row index
KNN2_MINMAX = 1.25     # VIMNMX(.3).U16x2 per compare (5 per 4 compares), ALU pipe
KNN2_COMPARES_PER_ITERATION = 8      # 4 train rows x 2 queries per thread

# knn2_mmaf_kernel (default): tensor cores on 4-bit operands, per 128 x 240 tile of compares
F4_INSTRUCTIONS_PER_TILE = 4         # tcgen05.mma kind::mxf4.block_scale, K = 64 each, K = 256 in all
F4_START_INSTRUCTIONS = 1            # + the kind::f8f6f4 instruction that starts the accumulator at 2^23 + 16384 + 127 - column (overhead)
F4_TILE_ROWS = 240                   # train rows per accumulator (two accumulators + the block scales fill the 512 TMEM columns)
F4_EPILOGUE_MINMAX3 = 0.5            # per compare: two sweeps of three-input packed max, each over two compares per register
F4_EPILOGUE_IMAD = 0.5               # per compare: the 32-bit multiply-add of the second sweep (two compares per register)

F4_WIDE_INSTRUCTIONS_PER_TILE = 8    # knn2_mmaf_kernel<true>: 64-byte rows, K = 512

# knn2_mmak_kernel (UZ_MATCH_MMA=4): tensor cores on int8 operands, per 128 x 256 tile of compares
MMA_INSTRUCTIONS_PER_TILE = 8        # tcgen05.mma kind::i8, K = 32 each, K = 256 in all
MMA_KEY_SLICE_INSTRUCTIONS = 1       # + the constant K-slice that turns the accumulator into the packed key (overhead, not counted as work)
MMA_OPS_PER_COMPARE = 512            # 256 int8 multiply-adds
MMA_EPILOGUE_IMAD = 0.0              # per compare: none (knn2_mma_kernel, UZ_MATCH_MMA=7, needed 1: key16 = dot * (-64) + (16384 + column))
MMA_EPILOGUE_MINMAX = 1.25
