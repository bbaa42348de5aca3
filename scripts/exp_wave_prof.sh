mkdir -p gpurun_out
WFB_WAVE_PROF=${1:-1} timeout 300 python bench.py --steps 1 --warmup 6 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
ls -la gpurun_out/
