mkdir -p gpurun_out
run() { # label, env-lib, extra args
  WFB_LIB=$2 timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 10 $3 > gpurun_out/bench_var.json 2>> gpurun_out/bench_var.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/bench_var.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'ms/step %.3f'%d['ms_per_step'], 'V1 %.3f'%d['stage_ms_per_step']['land_hydrology'], 'e2e ms %.3f'%d['e2e']['ms_per_step'])
PY
}
V=wflow.jl_b200/csrc/_obj/variants
D=wflow.jl_b200/libwflow_b200.so
run default $D ""
for v in ${VARIANTS}; do run $v $V/lib_$v.so ""; done
