#!/bin/bash
# Full ncu captures of the leader chain's two biggest kernels (k_follow_lla, k_alt_part) inside bench.py's chain, and of the
# long-leg sampler on the reference's own mission.  Usage (under gpurun): bash scripts/gpu_ncu_chain.sh <tag>
TAG=${1:-chain}
OUT=gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_follow_lla -s 2 -c 1 -o $OUT/${TAG}_follow -f \
    python bench.py --steps 20 --warmup 3 --configs none --no-parity > $OUT/${TAG}_ncu_follow.log 2>&1; tail -1 $OUT/${TAG}_ncu_follow.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_sample_scan -s 4 -c 1 -o $OUT/${TAG}_longscan -f \
    python scripts/b1_profile.py > $OUT/${TAG}_ncu_longscan.log 2>&1; tail -1 $OUT/${TAG}_ncu_longscan.log
