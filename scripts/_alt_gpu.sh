timeout 600 python -m pytest tests/test_gpu_followers.py tests/test_gpu_alt.py tests/test_gpu_robustness.py -x -q -m gpu 2>&1 | tail -4
timeout 120 python scripts/alt_bench.py 2>/dev/null | cut -c150-330
timeout 300 python bench.py --steps 300 --configs none --no-parity > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; tail -3 gpurun_out/r2w_bench.err
python - <<'P'
import json; d=json.load(open('gpurun_out/r2w_bench.json')); print(d['value'], d['ms_per_step']); print(json.dumps(d['leader_chain'])[:700])
P
