#!/bin/bash
# bench lines of V1 variants: VARIANTS="opt=val,opt=val opt=val ..."
mkdir -p gpurun_out
for v in $VARIANTS; do
  tag=$(echo $v | tr '=,' '__')
  timeout 300 python bench.py --steps ${STEPS:-30} --warmup 10 --no-cpu-baseline $(for o in $(echo $v | tr ',' ' '); do echo --option $o; done) \
      > gpurun_out/bench_ov_$tag.json 2> gpurun_out/bench_ov_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_ov_$tag.json"))
    print("$v", "ms/step", round(d["ms_per_step"], 4), "V1 ms", round(d["stage_ms_per_step"]["land_hydrology"], 4), "frac", round(d["roofline"]["frac"], 4))
except Exception as e:
    print("$v", "FAILED", e)
PY
done
