# developer aid: GPU parity tests + a short bench (run under gpurun)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/gputests.log
tail -12 gpurun_out/gputests.log
for sz in ${SIZES:-1000}; do
timeout 600 python bench.py --size $sz --steps ${STEPS:-10} --warmup 10 --no-cpu-baseline > gpurun_out/bench_b_$sz.log 2>&1
tail -1 gpurun_out/bench_b_$sz.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('size $sz', d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'])" || tail -5 gpurun_out/bench_b_$sz.log
done
