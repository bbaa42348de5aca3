mkdir -p gpurun_out
for k in 1 2 3; do
  WFB_WAVE_PROF=$k timeout 300 python bench.py --steps 1 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
  mv gpurun_out/wave_prof.csv gpurun_out/wave_prof_$k.csv
done
ls -la gpurun_out/
