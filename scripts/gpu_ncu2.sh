#!/bin/bash
# ncu --set full of one kernel for an arbitrary uniform batch.  KERNEL=regex ARGS="--batch 1024 --ns 512 --weights plain" bash scripts/gpu_ncu2.sh tag
TAG=${1:-n}; OUT=gpurun_out; mkdir -p $OUT
python scripts/profile_run.py $ARGS --iters 3 > $OUT/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s ${SKIP:-1} -c 1 -o $OUT/${TAG}_ncu -f \
    python scripts/profile_run.py $ARGS --iters 3 > $OUT/${TAG}_ncu.log 2>&1
tail -2 $OUT/${TAG}_ncu.log
