timeout 900 python -m pytest tests/test_gpu_full_parity.py tests/test_gpu_parity.py -m gpu -q -x --tb=short -s 2>&1 | grep -v "^\[parity\]" | tail -8
python bench.py --steps 300 --configs none --no-parity > gpurun_out/r2_b10.json 2>gpurun_out/r2_b10.err; python -c "
import json; d=json.load(open('gpurun_out/r2_b10.json')); print(d['value'], d['ms_per_step']); c=d['cpu_baseline']; print(c['value'], c['cores'], c['single_thread_value']); print(json.dumps(c['structured_cpu'])[:300])"
tail -3 gpurun_out/r2_b10.err
