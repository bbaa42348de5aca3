#!/bin/bash
# BASELINE configs[4] per-GPU unit: hourly model step (modified Rutter interception), one B200
mkdir -p gpurun_out
for S in "$@"; do
  steps=30; [ $S -gt 2000 ] && steps=8
  timeout 900 python bench.py --hourly --size $S --steps $steps --warmup 5 --no-cpu-baseline \
      > gpurun_out/bench_hourly_$S.json 2> gpurun_out/bench_hourly_$S.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_hourly_$S.json"))
    print("hourly $S", "ms/step", round(d["ms_per_step"], 3), "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "substeps", d["details"]["substeps"], "V1 frac", round(d["roofline"]["frac"], 3), {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
except Exception as e:
    print("$S", "FAILED", e)
PY
done
