# 8 x B200: the weak-scaling line of the default workload and the 1e8-cell configuration
# (3536^2 cells per GPU; smaller if the box has little host memory)
mkdir -p gpurun_out
G=${1:-8}
free -g | head -2 > gpurun_out/n${G}_box.txt; nproc >> gpurun_out/n${G}_box.txt
MEM=$(awk '/MemTotal/ {print int($2/1048576)}' /proc/meminfo)
SIZE=3536; if [ "$MEM" -lt $((G * 30)) ]; then SIZE=2500; fi
echo "host memory ${MEM} GB -> size ${SIZE}"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29521 bench.py --gpus $G --steps 20 --warmup 10 > gpurun_out/bench_n${G}.json 2> gpurun_out/bench_n${G}.err
tail -c 400 gpurun_out/bench_n${G}.json; echo
if [ -z "$SKIP_BIG" ]; then
timeout 1200 $RUN --master-port 29522 bench.py --gpus $G --size $SIZE --steps 5 --warmup 6 --no-cpu-baseline > gpurun_out/bench_n${G}_big.json 2> gpurun_out/bench_n${G}_big.err
fi
python -c "
import json
for f in ('gpurun_out/bench_n${G}.json','gpurun_out/bench_n${G}_big.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['n_gpus'], d['value'], d['ms_per_step'], d['config']['cells_per_gpu'], d['stage_ms_per_step'], d['e2e']['value'])
    except Exception as e: print(f, 'failed', e)
"
tail -3 gpurun_out/bench_n${G}_big.err
