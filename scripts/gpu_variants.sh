mkdir -p gpurun_out
run() { # label, extra args
  timeout 600 python bench.py --no-cpu-baseline $2 > gpurun_out/bench_var.json 2>> gpurun_out/bench_var.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/bench_var.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'ms/step %.3f'%d['ms_per_step'], 'V1 %.3f'%d['stage_ms_per_step']['land_hydrology'], 'e2e ms %.3f'%d['e2e']['ms_per_step'])
PY
}
run inline2_50 "--cfg unsat_inline_iters=2"
run inline4_50 "--cfg unsat_inline_iters=4"
run inline8_50 "--cfg unsat_inline_iters=8"
run inline2_20 "--cfg unsat_inline_iters=2 --steps 20 --warmup 10"
run inline8_20 "--cfg unsat_inline_iters=8 --steps 20 --warmup 10"
timeout 600 python -m pytest tests/test_gpu_cut_exchange.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
