timeout 300 python bench.py --steps 100 --configs cfg1 --no-parity > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; tail -3 gpurun_out/r2w_bench.err
python - <<'P'
import json; d=json.load(open('gpurun_out/r2w_bench.json')); v=d['variants']['cfg1']; print(v['single_call_ms'], v['ms_per_step'], json.dumps(v['dev_call_ms']))
P
