mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "Error|assert|passed|failed" | head -20
run() { # label, extra args
  timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 10 $2 > gpurun_out/bench_var.json 2>> gpurun_out/bench_var.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/bench_var.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'ms/step %.3f'%d['ms_per_step'], 'V1 %.3f'%d['stage_ms_per_step']['land_hydrology'])
PY
}
run default ""
run slices_1 "--cfg vertical_slices=1"
run slices_2 "--cfg vertical_slices=2"
run slices_3 "--cfg vertical_slices=3"
run noengine_1 "--option vertical_engine=0 --cfg vertical_slices=1"
