#!/bin/bash
# bench.py on the 8 GPUs of one box (gpurun --gpus 8 -- bash scripts/gpu_bench_8gpu.sh)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 500 --warmup 10 > gpurun_out/r2v8_bench_8gpu.json 2> gpurun_out/r2v8_bench_8gpu.err
tail -c 300 gpurun_out/r2v8_bench_8gpu.err
python - <<'P'
import json; d=json.load(open('gpurun_out/r2v8_bench_8gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'leader', d['leader_chain']['value']); print({k:(round(v.get('value')), round(v.get('ms_per_step'),3), v.get('imbalance_worst_over_mean')) for k,v in d['variants'].items()})
P
