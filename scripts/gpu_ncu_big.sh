#!/bin/bash
# ncu --set full captures of the dominant kernels at the LARGE configurations (cfg3: 1M x 8 uniform; cfg5: ragged CSR, dense
# output).  Usage (under gpurun): bash scripts/gpu_ncu_big.sh <tag>
set -u
TAG=${1:-big}
OUT=gpurun_out
mkdir -p $OUT
cap() {  # name kernel-regex skip config weights
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o $OUT/${TAG}_$1 -f \
      python scripts/profile_run.py --config $4 --weights $5 --iters 3 > $OUT/${TAG}_$1.log 2>&1
  echo "$1 rc=$?"
}
timeout 600 python scripts/profile_run.py --config cfg3 --weights plain --iters 3 > $OUT/${TAG}_plain_run.log 2>&1; echo "run rc=$?"; tail -2 $OUT/${TAG}_plain_run.log
cap scan_cfg3 k_sample_scan 1 cfg3 plain
cap fused_cfg3 k_fused_solve 1 cfg3 shipped
cap write_cfg5 'k_write' 1 cfg5 plain
cap count_cfg5 'k_count' 1 cfg5 plain
cap pair_cfg5 'k_thomas_pair' 1 cfg5 plain
ls -la $OUT | grep $TAG
