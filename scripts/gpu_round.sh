#!/bin/bash
# GPU box session: parity tests, smoke, bench line (+ tuning variants), ncu launch list and one
# ncu --set full capture of the vertical kernels. usage: gpu_round.sh <tag> [quick]
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_$TAG.txt 2>&1
nproc >> gpurun_out/smi_$TAG.txt
timeout 1800 python -m pytest tests -m gpu -x -q -s > gpurun_out/gputests_$TAG.log 2>&1
tail -3 gpurun_out/gputests_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 1500 gpurun_out/bench_$TAG.json
[ "$2" = quick ] && exit 0
for v in ${VARIANTS:-}; do
  WFB_LIB=wflow.jl_b200/csrc/_obj/variants/lib_$v.so timeout 300 python bench.py --no-cpu-baseline \
      > gpurun_out/bench_${TAG}_$v.json 2>> gpurun_out/bench_$TAG.err
done
for c in ${CFGS:-}; do
  timeout 300 python bench.py --no-cpu-baseline --cfg $c > gpurun_out/bench_${TAG}_$c.json 2>> gpurun_out/bench_$TAG.err
done
# every launch with its device time (cold-cache, serialised: compare SHARES); vertical_graph=0 so
# that ncu lists the kernels of the vertical update one by one
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    --option vertical_graph=0 > gpurun_out/bench_under_ncu_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'land_hydrology|unsat_engine|soil_column' -s 30 -c 3 \
    -o gpurun_out/prof_v1_$TAG python bench.py --steps 2 --warmup 10 --no-cpu-baseline \
    --option vertical_graph=0 > gpurun_out/bench_under_ncu2_$TAG.log 2>&1
ls -la gpurun_out | tail -12
