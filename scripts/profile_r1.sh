#!/bin/bash
# ncu evidence for round 1 (run under gpurun, 1 GPU). Outputs under gpurun_out/.
mkdir -p gpurun_out
# every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/bench_under_ncu.log 2>&1
# the SBM vertical kernel, full set
ncu --set full --clock-control none --import-source on -k regex:land_hydrology -s 2 -c 2 \
    -o gpurun_out/prof_v1_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/bench_under_ncu2.log 2>&1
ls -la gpurun_out
