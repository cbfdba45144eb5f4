for w in 5 3; do
timeout 300 python bench.py --gpus 1 --steps 20 --warmup $w > gpurun_out/r2v8_bench_short.json 2> gpurun_out/r2v8_bench_short.err; echo "short rc=$?"
python - <<'P'
import json; s=json.load(open('gpurun_out/r2v8_bench_short.json')); print('short', s['value'], s['steps'], s['warmup'], s['e2e']['value'], s['leader_chain']['value'], s['config']['streams_per_gpu'])
P
done
