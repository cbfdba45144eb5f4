timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_robustness.py tests/test_gpu_sharded.py tests/test_gpu_full_parity.py -m gpu -q -x --tb=short 2>&1 | tail -4
python bench.py --steps 300 --configs cfg3,cfg5 --no-parity --no-cpu-baseline > gpurun_out/r2_b6.json 2>gpurun_out/r2_b6.err; python -c "
import json; d=json.load(open('gpurun_out/r2_b6.json')); print(d['value'], d['ms_per_step'], d['config']['single_stream_ms_per_step'], d['roofline']['kernels_ms_per_step'], d['roofline']['frac'])
for k,v in d['variants'].items():
    if 'kernels_ms_per_step' in v: print(k, '%.4g'%v['value'], v['ms_per_step'], {a:round(b,3) for a,b in v['kernels_ms_per_step'].items()}, v['rows_per_step'])
    else: print(k, '%.4g'%v['value'], v['ms_per_step'], v.get('single_stream_ms_per_step'))"
tail -3 gpurun_out/r2_b6.err
