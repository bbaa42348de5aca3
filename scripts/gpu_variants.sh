timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "local_inertial" 2>&1 | tail -15
