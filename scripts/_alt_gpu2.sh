timeout 300 python bench.py --steps 300 --configs none --no-parity > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; tail -3 gpurun_out/r2w_bench.err
python - <<'P'
import json; d=json.load(open('gpurun_out/r2w_bench.json')); print(d['value'], d['ms_per_step']); print(d['leader_chain']['value'], d['leader_chain']['ms_per_step'], json.dumps(d['leader_chain']['kernels_ms_per_step']))
P
