#!/bin/bash
# adaptive internal time steps: parity tests, Moselle-shape and 1000^2 bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "adaptive or moselle or shards or reservoirs or floodplain" 2>&1 | tail -3
timeout 600 python bench.py --moselle --adaptive --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_moselleadaptive2.json 2> gpurun_out/bench_moselleadaptive2.err
timeout 600 python bench.py --adaptive --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/bench_adaptive_1000.json 2> gpurun_out/bench_adaptive_1000.err
python - <<PY
import json
for t in ("moselleadaptive2", "adaptive_1000"):
    try:
        d = json.load(open(f"gpurun_out/bench_{t}.json"))
        print(t, "ms/step", round(d["ms_per_step"], 3), "substeps", d["details"]["substeps"], "launches", d["gpu_launches"], {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()})
    except Exception as e:
        print(t, "FAILED", e)
PY
