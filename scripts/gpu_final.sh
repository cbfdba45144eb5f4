#!/bin/bash
# Final pass of a round on one B200: scripts/gpu_check.sh (all GPU tests, smoke, bench, reference arm, launch list, ncu captures of
# the solver and sampler) plus the altitude stage (both execution forms timed, launch list, full ncu capture of k_alt_part).
# Usage (under gpurun): bash scripts/gpu_final.sh   (edit the tag below per version)
bash scripts/gpu_check.sh r2v8 > gpurun_out/r2v8_check.log 2>&1
tail -12 gpurun_out/r2v8_check.log | head -8
NCU=1 TAG=r2v8
timeout 120 python scripts/alt_bench.py --policy 0 > gpurun_out/r2v8_alt_bench_pairs.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2v8_alt_launches.csv python scripts/alt_bench.py --iters 2 > gpurun_out/r2v8_alt_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_alt_part -s 3 -c 1 -o gpurun_out/r2v8_alt_solve -f python scripts/alt_bench.py --iters 2 > gpurun_out/r2v8_alt_ncu2.log 2>&1
tail -2 gpurun_out/r2v8_alt_ncu2.log
