#!/bin/bash
# Runs ON the GPU box (gpurun): plain bench, ncu launch list, ncu --set full of the main kernels.
# usage: tools/profile_gpu.sh <tag>      outputs under gpurun_out/
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --in-flight 1"
mkdir -p gpurun_out
python bench.py --steps 500 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || exit 1
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
# one step = 12 launches; skip synth + 3 warm-up steps
ncu --set full --clock-control none --import-source on \
    -k regex:"count_kernel|image_kernel|parse_emit|parse_mask|bucket_count|bucket_scatter|fold_kernel|reduce_slabs|segment" \
    -s 24 -c 10 -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log
cat gpurun_out/bench_${TAG}.json
