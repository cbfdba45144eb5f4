#!/bin/bash
# GPU-box pass for the altitude-optimisation row: all GPU tests, the alt benchmark, ncu launch list + full capture of
# k_alt_part.  Usage (under gpurun): bash scripts/gpu_alt.sh <tag>
set -u
# Every step runs under its own `timeout`: a hung kernel must cost minutes, not the whole gpurun limit.
TAG=${1:-alt}
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log
timeout 400 python scripts/alt_bench.py > $OUT/${TAG}_alt_bench.json 2> $OUT/${TAG}_alt_bench.err; echo "alt bench rc=$?"
cat $OUT/${TAG}_alt_bench.json; tail -3 $OUT/${TAG}_alt_bench.err
timeout 400 python scripts/alt_bench.py --policy 0 > $OUT/${TAG}_alt_bench_pairs.json 2>> $OUT/${TAG}_alt_bench.err; cat $OUT/${TAG}_alt_bench_pairs.json
timeout 400 python bench.py --impl rows-cpu > $OUT/${TAG}_rows_cpu.json 2> $OUT/${TAG}_rows_cpu.err; echo "rows-cpu rc=$?"
if [ "${NCU:-1}" = "1" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_alt_launches.csv \
    python scripts/alt_bench.py --iters 2 > $OUT/${TAG}_alt_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_alt_part -s 3 -c 1 -o $OUT/${TAG}_alt_solve -f \
    python scripts/alt_bench.py --iters 2 > $OUT/${TAG}_alt_ncu2.log 2>&1
fi
ls $OUT | grep ${TAG}
