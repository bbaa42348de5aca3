#!/bin/bash
# BASELINE configs[0]: Moselle-shape basin, fixed and adaptive internal time steps, with the CPU
# port timed beside it and the parity check at this size
mkdir -p gpurun_out
for mode in "" "--adaptive"; do
  tag=moselle$(echo $mode | tr -d ' -')
  timeout 600 python bench.py --moselle $mode --steps 30 --warmup 10 --cpu-steps 5 \
      > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$tag.json"))
    print("$tag", "ms/step", round(d["ms_per_step"], 3), "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "cpu %.4g (%d cores)" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]), "substeps", d["details"]["substeps"], "parity", d["parity_checked"].get("ok"), d["parity_checked"].get("worst_rel", d["parity_checked"].get("error")))
except Exception as e:
    print("$tag", "FAILED", e)
PY
  tail -2 gpurun_out/bench_$tag.err
done
