#!/bin/bash
# V1 with the loop engine beside soil_column_kernel (vertical_overlap = 1, default) against the
# plain sequence (0): parity tests of the vertical update, then bench lines of both.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q \
    > gpurun_out/gputests_ov.log 2>&1
tail -5 gpurun_out/gputests_ov.log
for v in ${VARIANTS:-vertical_overlap=1 vertical_overlap=0}; do
  tag=$(echo $v | tr '= ,' '___')
  timeout 300 python bench.py --steps ${STEPS:-30} --warmup 10 --no-cpu-baseline $(for o in $(echo $v | tr ',' ' '); do echo --option $o; done) \
      > gpurun_out/bench_ov_$tag.json 2> gpurun_out/bench_ov_$tag.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_ov_$tag.json"))
print("$v", "ms/step", round(d["ms_per_step"], 4), "V1 ms", round(d["stage_ms_per_step"]["land_hydrology"], 4), "frac", round(d["roofline"]["frac"], 4))
PY
done
