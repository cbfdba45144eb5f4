#!/bin/bash
# One GPU-box pass: parity tests, bench, then ncu launch list + full captures of the two dominant kernels.
# Usage (under gpurun): bash scripts/gpu_check.sh <tag>
set -u
# Every step runs under its own `timeout`: a hung kernel must cost minutes, not the whole gpurun limit.
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $OUT/${TAG}_clocks.csv 2>&1 &
SMI=$!
timeout 1200 python -m pytest tests -m gpu -x -q -s > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 400 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
timeout 900 python bench.py --steps 1000 --warmup 20 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "bench ref rc=$?"
timeout 400 python scripts/geo_bench.py > $OUT/${TAG}_geo_bench.jsonl 2> $OUT/${TAG}_geo_bench.err; echo "geo bench rc=$?"
timeout 400 python scripts/alt_bench.py > $OUT/${TAG}_alt_bench.json 2> $OUT/${TAG}_alt_bench.err; echo "alt bench rc=$?"
kill $SMI
timeout 400 python bench.py --impl rows-cpu > $OUT/${TAG}_rows_cpu.json 2> $OUT/${TAG}_rows_cpu.err; echo "rows-cpu rc=$?"
if [ "${NCU:-1}" = "1" ]; then
timeout 400 python scripts/profile_run.py --weights shipped --iters 4 > $OUT/${TAG}_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches.csv \
    python scripts/profile_run.py --weights shipped --iters 4 > $OUT/${TAG}_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_solve -s 2 -c 1 -o $OUT/${TAG}_fused -f \
    python scripts/profile_run.py --weights shipped --iters 4 > $OUT/${TAG}_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sample_scan -s 2 -c 1 -o $OUT/${TAG}_scan -f \
    python scripts/profile_run.py --weights shipped --iters 4 > $OUT/${TAG}_ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_solve -s 2 -c 1 -o $OUT/${TAG}_fused_plain -f \
    python scripts/profile_run.py --weights plain --iters 4 > $OUT/${TAG}_ncu4.log 2>&1
fi
ls -la $OUT
