# developer aid: stage timings at larger tiles (the 1e8-cell / 8 GPU configuration is 3536^2 per GPU)
mkdir -p gpurun_out
for sz in 2000 3536; do
  timeout 900 python bench.py --size $sz --steps 5 --warmup 6 --no-cpu-baseline > gpurun_out/bench_size_$sz.json 2> gpurun_out/bench_size_$sz.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/bench_size_$sz.json').read().strip().splitlines()[-1]); print('size', $sz, d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['e2e'])"
done
