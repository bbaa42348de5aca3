#!/bin/bash
# GPU box session: parity tests, bench line, ncu launch list, ncu full captures of the vertical
# kernels and of the overland wave kernel. usage: gpu_round.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_$TAG.txt 2>&1
nproc >> gpurun_out/smi_$TAG.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_$TAG.log 2>&1
tail -3 gpurun_out/gputests_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 600 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
# every launch with its device time (cold-cache, serialised: compare SHARES); WFB_NO_GRAPH so
# that ncu sees the kernels of the vertical update one by one
WFB_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 10 --no-cpu-baseline \
    > gpurun_out/bench_under_ncu_$TAG.log 2>&1
WFB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'land_surface|soil_column|unsat_loop|unsat_resume' -s 40 -c 6 \
    -o gpurun_out/prof_v1_$TAG python bench.py --steps 2 --warmup 10 --no-cpu-baseline \
    > gpurun_out/bench_under_ncu2_$TAG.log 2>&1
WFB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'surface_wave|subsurface_wave' -s 20 -c 2 \
    -o gpurun_out/prof_wave_$TAG python bench.py --steps 2 --warmup 10 --no-cpu-baseline \
    > gpurun_out/bench_under_ncu3_$TAG.log 2>&1
ls -la gpurun_out | tail -12
