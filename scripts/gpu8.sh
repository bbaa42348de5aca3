mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu_r2.json 2> gpurun_out/bench_8gpu_r2.err
tail -1 gpurun_out/bench_8gpu_r2.json | cut -c1-3000
tail -5 gpurun_out/bench_8gpu_r2.err
