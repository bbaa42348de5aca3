# developer aid: GPU parity tests + a short bench + per-chunk wave profiles (run under gpurun)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/gputests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.log 2>&1
tail -12 gpurun_out/gputests.log
tail -1 gpurun_out/bench_b.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms_per_step'])"
bash scripts/exp_wave_prof.sh > /dev/null 2>&1
