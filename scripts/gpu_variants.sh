mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/vertical_timeline.py 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 10 > gpurun_out/bench_r2l.json 2> gpurun_out/bench_r2l.err
tail -1 gpurun_out/bench_r2l.json | cut -c1-1800
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
    --log-file gpurun_out/launches_r2l.csv python bench.py --steps 2 --warmup 10 --no-cpu-baseline \
    --option vertical_graph=0 > gpurun_out/bench_under_ncu_r2l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'land_hydrology|unsat_engine|soil_column' -s 30 -c 3 \
    -o gpurun_out/prof_v1_r2l python bench.py --steps 2 --warmup 10 --no-cpu-baseline \
    --option vertical_graph=0 > gpurun_out/bench_under_ncu2_r2l.log 2>&1
ls -la gpurun_out/*r2l*
