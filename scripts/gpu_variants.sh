mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_memcheck_smoke.log 2>&1
tail -3 gpurun_out/r2_memcheck_smoke.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_cut_exchange.py -m gpu -x -q -k "floodplain or cut_basin or snow_transport or layered or two_shards or reservoirs or degenerate" > gpurun_out/r2_memcheck_tests.log 2>&1
tail -4 gpurun_out/r2_memcheck_tests.log
timeout 600 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_racecheck_smoke.log 2>&1
tail -3 gpurun_out/r2_racecheck_smoke.log
