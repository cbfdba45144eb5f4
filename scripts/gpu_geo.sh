#!/bin/bash
# GPU-box pass for the WGS84 <-> ENU row: all GPU tests, the geo micro-benchmark, then an ncu launch list and one full
# capture of k_enu_to_wgs84.  Usage (under gpurun): bash scripts/gpu_geo.sh <tag>
set -u
# Every step runs under its own `timeout`: a hung kernel must cost minutes, not the whole gpurun limit.
TAG=${1:-geo}
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -15 $OUT/${TAG}_pytest.log
timeout 400 python scripts/geo_bench.py > $OUT/${TAG}_geo_bench.jsonl 2> $OUT/${TAG}_geo_bench.err; echo "geo bench rc=$?"
cat $OUT/${TAG}_geo_bench.jsonl; tail -3 $OUT/${TAG}_geo_bench.err
timeout 400 python bench.py --impl rows-cpu > $OUT/${TAG}_rows_cpu.json 2> $OUT/${TAG}_rows_cpu.err; echo "rows-cpu rc=$?"
if [ "${NCU:-1}" = "1" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_geo_launches.csv \
    python scripts/geo_bench.py --iters 2 > $OUT/${TAG}_geo_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_enu_to_wgs84 -s 4 -c 1 -o $OUT/${TAG}_enu_to_wgs84 -f \
    python scripts/geo_bench.py --iters 2 > $OUT/${TAG}_geo_ncu2.log 2>&1
fi
ls -la $OUT | tail -12
