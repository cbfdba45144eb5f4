#!/bin/bash
# 8-GPU session: D2H ceiling at 1/2/4/8 ranks (with / without NUMA placement), then the bench at 8 ranks.
OUT=gpurun_out
for n in 1 2 4 8; do for numa in 0 1; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n scripts/d2h_ceiling.py --pin $numa --mb 20 --out $OUT/r2_d2h_ceiling.jsonl > /dev/null 2>&1
done; done
cat $OUT/r2_d2h_ceiling.jsonl | cut -c1-400
nvidia-smi topo -m > $OUT/r2_topo.txt 2>&1; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" > $OUT/r2_lscpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 500 --warmup 10 > $OUT/r2_bench_8gpu.json 2> $OUT/r2_bench_8gpu.err
tail -c 300 $OUT/r2_bench_8gpu.err; wc -c $OUT/r2_bench_8gpu.json
