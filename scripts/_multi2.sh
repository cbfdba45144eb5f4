timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 500 --warmup 10 > gpurun_out/r2v6_bench_2gpu.json 2> gpurun_out/r2v6_bench_2gpu.err
tail -c 300 gpurun_out/r2v6_bench_2gpu.err
python - <<'P'
import json; d=json.load(open('gpurun_out/r2v6_bench_2gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['leader_chain']['value']); print({k:(v.get('value'), v.get('ms_per_step')) for k,v in d['variants'].items()})
P
timeout 300 python -m pytest tests -m gpu -q -x -k "shard or multi or two_gpu" 2>&1 | tail -3
