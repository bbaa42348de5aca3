# developer aid: vertical-update slices (WFB_V_SLICES)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/gputests.log; tail -2 gpurun_out/gputests.log
for sl in ${SLICES:-1 2 4 8}; do
  for sz in ${SIZES:-1000}; do
    WFB_V_SLICES=$sl timeout 600 python bench.py --size $sz --steps 10 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('slices $sl', $sz, 'total %.3f'%d['ms_per_step'], ' '.join('%s=%.3f'%(k[:5],v) for k,v in s.items()))" || echo "$sl $sz failed"
  done
done
