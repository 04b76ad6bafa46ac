#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list, ncu --set full of the kernels of the step.
# Usage (from the repo root, via gpurun): bash scripts/gpu_round.sh <tag> [quick]
TAG=${1:-r02}
QUICK=${2:-}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
cat gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_ref_$TAG.json
timeout 600 adapter/bench_adapter 10000 25000 3 > gpurun_out/bench_adapter_$TAG.json 2>> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_adapter_$TAG.json
[ -n "$QUICK" ] && exit 0
# launch list of the default step: the tensor-core match kernel, then the solve kernel behind it on the same stream, so
# the per-kernel shares of this (serialised, cold-cache) list are comparable with a live step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn2_mmaf_kernel -s 3 -c 1 -f -o gpurun_out/knn2_mmaf_full_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-places --no-extras > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 3 -c 1 -f -o gpurun_out/solve_full_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-places --no-extras > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:derive_layouts_kernel -c 1 -f -o gpurun_out/derive_layouts_full_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-places --no-extras > /dev/null 2>&1
# 512-bit rows: the 4-bit tensor-core kernel (default), the int8 one and the integer-pipe one; then the 256-bit integer-pipe fallback
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn2_mmaf_kernel -s 2 -c 1 -f -o gpurun_out/knn2_mmaf_wide_full_$TAG \
    python scripts/gpu_wide_probe.py 200 > /dev/null 2>&1
UZ_MATCH_MMA_WIDE=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn2_mmaw_kernel -s 2 -c 1 -f -o gpurun_out/knn2_mmaw_full_$TAG \
    python scripts/gpu_wide_probe.py 200 > /dev/null 2>&1
UZ_MATCH_MMA_WIDE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn2_wide_kernel -s 2 -c 1 -f -o gpurun_out/knn2_wide_full_$TAG \
    python scripts/gpu_wide_probe.py 200 > /dev/null 2>&1
UZ_MATCH_MMA=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn2_kernel -s 3 -c 1 -f -o gpurun_out/knn2_full_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-places --no-extras > /dev/null 2>&1
for k in places_insert_kernel places_vote_kernel places_select_kernel; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/${k}_full_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
done
ls -la gpurun_out
# summarise on the box (gpurun copies back at most 64 MiB): one small JSON per capture, the reports themselves stay behind
# except the match kernel's
for rep in gpurun_out/*_full_$TAG.ncu-rep; do
  name=$(basename $rep .ncu-rep); name=${name%_full_$TAG}
  python scripts/ncu_summary.py $rep gpurun_out/${name}_${TAG}_ncu_full.json
  case $name in knn2_mmaf) ;; *) rm -f $rep ;; esac
done
du -sh gpurun_out
