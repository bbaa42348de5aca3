mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cut_exchange.py -m gpu -x -q -s 2>&1 | tail -30
