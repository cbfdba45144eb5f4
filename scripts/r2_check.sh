timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_robustness.py -m gpu -q -x --tb=short 2>&1 | tail -6
python scripts/phase_clocks.py 4096 16 2>&1 | tail -11
python bench.py --steps 500 --configs cfg1,cfg3 --no-parity --no-cpu-baseline > gpurun_out/r2_b4.json 2>gpurun_out/r2_b4.err; python -c "
import json; d=json.load(open('gpurun_out/r2_b4.json')); print(d['value'], d['ms_per_step'], d['config']['single_stream_ms_per_step'], d['roofline']['kernels_ms_per_step'], d['roofline']['frac']); print(d['variants']['plain']['roofline']['kernels_ms_per_step'])
for k,v in d['variants'].items():
    if 'kernels_ms_per_step' in v: print(k, '%.4g'%v['value'], v['ms_per_step'], v['kernels_ms_per_step'], v.get('single_call_ms'))"
MSNAP_SAMPLER=scan python bench.py --steps 500 --configs none --no-parity --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('scan sampler:', d['value'], d['ms_per_step'], d['roofline']['kernels_ms_per_step'])"
