#!/bin/bash
# usage (under gpurun): bash tools/profile.sh <tag> [blocks]
# 1. launch list of one bench step (every kernel with its device time; cold-cache, serialised)
# 2. full ncu capture of the dense column-walk kernel (the one launch of a shared-walk step)
TAG=${1:-r02}; BLOCKS=${2:-1000000}
mkdir -p gpurun_out
COMMON="--blocks $BLOCKS --steps 1 --warmup 3 --cpu-sample-mb 1 --no-e2e --no-indep --no-side --no-sharded"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py $COMMON > gpurun_out/launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_walk_tma -s 4 -c 1 -f -o gpurun_out/prof_walk_${TAG} \
    python bench.py $COMMON > gpurun_out/prof_walk_${TAG}.log 2>&1
ls -la gpurun_out/ | tail -6
