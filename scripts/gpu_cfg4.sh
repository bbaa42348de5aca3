#!/bin/bash
# BASELINE configs[3]: local-inertial routing on a 4000 x 4000 basin, one B200: river + 1-D floodplain
# and 2-D overland flow + river.
mkdir -p gpurun_out
S=${1:-4000}
for mode in local-inertial local-inertial-land; do
  timeout 900 python bench.py --$mode --size $S --steps 3 --warmup 2 --no-cpu-baseline \
      > gpurun_out/bench_cfg4_${mode}_$S.json 2> gpurun_out/bench_cfg4_${mode}_$S.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_cfg4_${mode}_$S.json"))
    print("$mode $S", "ms/step", round(d["ms_per_step"], 3), "value", "%.4g" % d["value"], "substeps", d["details"]["substeps"], "stages", {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
except Exception as e:
    print("$mode", "FAILED", e)
PY
done
