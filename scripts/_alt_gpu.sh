timeout 300 python -m pytest tests/test_gpu_alt.py tests/test_gpu_followers.py tests/test_gpu_robustness.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2w_alt_pytest.log
cat gpurun_out/r2w_alt_pytest.log
timeout 120 python scripts/alt_bench.py > gpurun_out/r2w_alt_bench.json 2> gpurun_out/r2w_alt_bench.err; tail -3 gpurun_out/r2w_alt_bench.err; cut -c1-400 gpurun_out/r2w_alt_bench.json
