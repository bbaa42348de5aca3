mkdir -p gpurun_out
run() { # label, env-lib, extra args
  WFB_LIB=$2 timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 10 $3 > gpurun_out/bench_var.json 2>> gpurun_out/bench_var.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/bench_var.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'ms/step %.3f'%d['ms_per_step'], 'V1 %.3f'%d['stage_ms_per_step']['land_hydrology'])
PY
}
V=wflow.jl_b200/csrc/_obj/variants
run default wflow.jl_b200/libwflow_b200.so ""
for v in ${VARIANTS}; do run $v $V/lib_$v.so ""; done
run default_1slice wflow.jl_b200/libwflow_b200.so "--cfg vertical_slices=1"
timeout 1200 python -m pytest tests -m gpu -x -q -k "reservoir or update_model_matches or alternative or layered" 2>&1 | tail -5
