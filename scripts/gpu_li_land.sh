#!/bin/bash
# GPU box session for the 2-D local-inertial overland flow: its parity tests (with the river-only
# local-inertial tests as regression), then a short bench line. usage: gpu_li_land.sh [size]
SIZE=${1:-1000}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_lil.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "local_inertial" \
    > gpurun_out/gputests_lil.log 2>&1
tail -25 gpurun_out/gputests_lil.log
timeout 400 python bench.py --local-inertial-land --size $SIZE --steps 5 --warmup 3 --no-cpu-baseline \
    > gpurun_out/bench_lil_$SIZE.json 2> gpurun_out/bench_lil_$SIZE.err
tail -c 1800 gpurun_out/bench_lil_$SIZE.json; tail -5 gpurun_out/bench_lil_$SIZE.err
