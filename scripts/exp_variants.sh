# developer aid: stage timings of the kernel-tuning variants (scripts/build_variants.sh)
mkdir -p gpurun_out
run() { # label, env...
  label=$1; shift
  for sz in ${SIZES:-1000 3536}; do
    env "$@" timeout 600 python bench.py --size $sz --steps 5 --warmup 8 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('$label', $sz, 'total %.3f'%d['ms_per_step'], ' '.join('%s=%.3f'%(k[:5],v) for k,v in s.items()))" || echo "$label $sz failed"
  done
}
V=wflow.jl_b200/csrc/_obj/variants
run base WFB_SSF_BANDS=0
run ssf2 WFB_SSF_BANDS=0 WFB_LIB=$V/lib_ssf2.so
run olf3 WFB_SSF_BANDS=0 WFB_LIB=$V/lib_olf3.so
run v128 WFB_SSF_BANDS=0 WFB_LIB=$V/lib_v128.so
run v128b WFB_SSF_BANDS=0 WFB_LIB=$V/lib_v128b.so
run inl4 WFB_SSF_BANDS=0 WFB_INLINE_ITERS=4
run inl16 WFB_SSF_BANDS=0 WFB_INLINE_ITERS=16
run inl32 WFB_SSF_BANDS=0 WFB_INLINE_ITERS=32
