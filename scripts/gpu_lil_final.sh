mkdir -p gpurun_out
bash scripts/gpu_lil_bench.sh 4000
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'local_inertial_land_river' -s 3 -c 1 \
    -o gpurun_out/prof_lil_r2final2 python bench.py --local-inertial-land --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/bench_under_ncu3_r2final2.log 2>&1
ls -la gpurun_out/prof_lil_r2final2.ncu-rep
