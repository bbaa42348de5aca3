#!/bin/bash
# Final GPU session of round 2: all parity tests, smoke, the default bench line (+ reference arm),
# ncu launch list, ncu --set full of the vertical kernels and of the land local-inertial kernel,
# compute-sanitizer on the new kernel.
TAG=r2final
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_$TAG.txt 2>&1
nproc >> gpurun_out/smi_$TAG.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/gputests_$TAG.log 2>&1
tail -3 gpurun_out/gputests_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 600 gpurun_out/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_$TAG.err
tail -c 300 gpurun_out/bench_${TAG}_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    --option vertical_graph=0 > gpurun_out/bench_under_ncu_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'land_hydrology|unsat_engine|soil_column' -s 30 -c 3 \
    -o gpurun_out/prof_v1_$TAG python bench.py --steps 2 --warmup 10 --no-cpu-baseline \
    --option vertical_graph=0 > gpurun_out/bench_under_ncu2_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'local_inertial_land_river' -s 3 -c 1 \
    -o gpurun_out/prof_lil_$TAG python bench.py --local-inertial-land --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/bench_under_ncu3_$TAG.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
    -k "land_edge_cases or land_and_river" > gpurun_out/sanitizer_memcheck_lil_$TAG.log 2>&1
tail -4 gpurun_out/sanitizer_memcheck_lil_$TAG.log
ls -la gpurun_out | tail -14
