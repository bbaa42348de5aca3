mkdir -p gpurun_out
run() { # label, extra args
  timeout 900 python bench.py --no-cpu-baseline --steps 10 --warmup 5 $2 > gpurun_out/bench_var.json 2>> gpurun_out/bench_var.err
  python - "$1" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/bench_var.json').read().strip().splitlines()[-1])
s=d['stage_ms_per_step']
print(sys.argv[1], 'ms/step %.3f'%d['ms_per_step'], 'value %.3g'%d['value'], {k: round(v,2) for k,v in s.items()}, d['details'].get('substeps'))
open('gpurun_out/bench_li_'+sys.argv[1]+'.json','w').write(json.dumps(d))
PY
}
run li_1000 "--local-inertial"
run li_2000 "--local-inertial --size 2000"
tail -3 gpurun_out/bench_var.err
