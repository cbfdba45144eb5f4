timeout 600 python -m pytest tests/test_gpu_followers.py tests/test_gpu_alt.py tests/test_gpu_robustness.py -x -q -m gpu 2>&1 | tail -4
timeout 120 python scripts/alt_bench.py 2>/dev/null | cut -c150-330
