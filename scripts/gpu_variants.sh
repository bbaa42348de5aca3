mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "update_model_matches or selftest" 2>&1 | grep -E "Error|assert|passed|failed|selftest" | head -20
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'unsat_loop' -s 16 -c 2 \
    -o gpurun_out/prof_loop_r2e python bench.py --steps 2 --warmup 10 --no-cpu-baseline \
    --option vertical_graph=0 --cfg vertical_slices=1 > gpurun_out/bench_under_ncu_r2e.log 2>&1
ls -la gpurun_out/*.ncu-rep
