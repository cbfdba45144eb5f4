timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 300 --configs cfg1 --no-parity > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; tail -3 gpurun_out/r2w_bench.err
python - <<'P'
import json; d=json.load(open('gpurun_out/r2w_bench.json')); v=d['variants']['cfg1']; print(d['value'], d['ms_per_step'], 'cfg1', v['single_call_ms'], v['ms_per_step'], json.dumps(v['dev_call_ms'])[:120]); print(d['leader_chain']['kernels_ms_per_step'])
P
