#!/bin/bash
# Runs ON the GPU box (gpurun): ncu launch list of a short bench run, then ncu --set full of the step's kernels.
# usage: tools/profile_gpu.sh <tag>      outputs under gpurun_out/
set -u
TAG=${1:-r03}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side-legs --e2e-steps 1 --in-flight 1"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
# one step = 13 launches (begin, mask, scan, emit, plan, hist, calibrate, scatter, long_reads, count, reduce, fold, image); skip the generator + warm-up steps
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on \
    -k regex:"countt_kernel|count_kernel|image_kernel|long_reads|parse_emit|parse_mask|bucket_scatter|fold_kernel|reduce_slabs|plan_kernel|parse_scan|step_begin|prio_hist|thr_calibrate" \
    -s 52 -c 13 -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full rc=$?"
tail -2 gpurun_out/ncu_full_${TAG}.log
ls -la gpurun_out/prof_${TAG}.ncu-rep
