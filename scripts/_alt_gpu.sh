timeout 600 python bench.py --steps 200 --configs cfg1,cfg4 --no-parity > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; tail -3 gpurun_out/r2w_bench.err
python - <<'P'
import json; d=json.load(open('gpurun_out/r2w_bench.json')); print(d['value'], d['ms_per_step'])
for k,v in d['variants'].items(): print(k, {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','single_call_ms')})
P
