#!/bin/bash
# ncu --set full capture of one kernel of the bench workload.  Usage: KERNEL=regex WEIGHTS=shipped|plain bash scripts/gpu_ncu.sh <tag>
TAG=${1:-n}; OUT=gpurun_out; mkdir -p $OUT
python scripts/profile_run.py --weights ${WEIGHTS:-shipped} --iters 4 > $OUT/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s 2 -c 1 -o $OUT/${TAG}_ncu -f \
    python scripts/profile_run.py --weights ${WEIGHTS:-shipped} --iters 4 > $OUT/${TAG}_ncu.log 2>&1
tail -2 $OUT/${TAG}_ncu.log
