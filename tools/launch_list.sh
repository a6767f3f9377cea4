#!/bin/bash
# Per-launch device times of a few eager training steps (ncu serialises launches: compare shares, not absolutes).
set -e
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-6000} -c ${COUNT:-3000} --csv \
    --log-file gpurun_out/launches_${TAG:-train}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras ${EXTRA} > gpurun_out/launches_${TAG:-train}.log 2>&1 || true
python tools/summarize_launches.py gpurun_out/launches_${TAG:-train}.csv | tee gpurun_out/launches_${TAG:-train}_summary.txt
