#!/bin/bash
# local-inertial river + 1-D floodplain: parity tests, then bench lines at the given sizes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "local_inertial" 2>&1 | tail -3
for S in "$@"; do
  timeout 900 python bench.py --local-inertial --size $S --steps 3 --warmup 2 --no-cpu-baseline \
      > gpurun_out/bench_li_$S.json 2> gpurun_out/bench_li_$S.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_li_$S.json"))
    print("river LI + floodplain $S", "ms/step", round(d["ms_per_step"], 3), "substeps", d["details"]["substeps"], "river ms", round(d["stage_ms_per_step"]["river"], 3))
except Exception as e:
    print("$S", "FAILED", e)
PY
done
