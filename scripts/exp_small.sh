mkdir -p gpurun_out
for sz in 64 128 256 512; do
  WFB_WAVE_PROF=1 timeout 300 python bench.py --size $sz --steps 1 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('size', $sz, d['ms_per_step'], d['stage_ms_per_step'])"
  mv gpurun_out/wave_prof.csv gpurun_out/wave_prof_olf_$sz.csv
done
