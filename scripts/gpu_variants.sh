mkdir -p gpurun_out
rm -f gpurun_out/cut_basin_r2.jsonl
for cfgline in "2 300 300" "2 1000 1000" ; do
set -- $cfgline
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29533 scripts/cut_basin_multi_gpu.py --d1 $2 --d2 $3 --steps 10 2>&1 | grep "^{" | tee -a gpurun_out/cut_basin_r2.jsonl | cut -c1-400
done
