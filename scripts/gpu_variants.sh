timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "floodplain" 2>&1 | tail -25
