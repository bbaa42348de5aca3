# developer aid: per-bundle timing of the single-sub-step subsurface kernel
mkdir -p gpurun_out
for w in ${WARPS:-7 1}; do
WFB_SSF_BANDS=1 WFB_BAND_WARPS=$w WFB_BAND_PROF=1 timeout 300 python bench.py --size ${1:-1000} --steps 1 --warmup 6 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('warps $w', d['stage_ms_per_step']['subsurface'])"
mv gpurun_out/band_prof.csv gpurun_out/band_prof_w$w.csv
done
