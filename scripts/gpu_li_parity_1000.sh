#!/bin/bash
# parity at the benchmarked size for the local-inertial schemes: bench.py's parity_checked leg
# (the CPU port starts from the GPU's spun-up state and every field is compared)
mkdir -p gpurun_out
for mode in local-inertial local-inertial-land; do
  timeout 600 python bench.py --$mode --steps 5 --warmup 3 --cpu-steps 1 \
      > gpurun_out/bench_parity_${mode}_1000.json 2> gpurun_out/bench_parity_${mode}_1000.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_parity_${mode}_1000.json"))
    print("$mode", "ms/step", round(d["ms_per_step"], 3), "parity", d["parity_checked"], "cpu %.4g (%d cores)" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]))
except Exception as e:
    print("$mode", "FAILED", e)
PY
  tail -2 gpurun_out/bench_parity_${mode}_1000.err
done
