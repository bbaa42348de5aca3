mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_r2l.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    --option vertical_graph=0 > gpurun_out/bench_under_ncu_r2l.log 2>&1
wc -l gpurun_out/launches_r2l.csv
