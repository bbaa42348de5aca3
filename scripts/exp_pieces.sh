# developer aid: stage timings for chunk piece depths (WFB_PIECE_LAND / WFB_PIECE_RIVER)
mkdir -p gpurun_out
run() { # label, env...
  label=$1; shift
  for sz in ${SIZES:-1000 3536}; do
    env "$@" timeout 600 python bench.py --size $sz --steps 5 --warmup 8 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('$label', $sz, 'total %.3f'%d['ms_per_step'], ' '.join('%s=%.3f'%(k[:5],v) for k,v in s.items()))" || echo "$label $sz failed"
  done
}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run L4R0
run L0R0 WFB_PIECE_LAND=0
run L6R0 WFB_PIECE_LAND=6
run L4R4 WFB_PIECE_RIVER=4
run L4R8 WFB_PIECE_RIVER=8
