timeout 900 python bench.py --steps 1000 --warmup 20 > gpurun_out/r2v7_bench.json 2> gpurun_out/r2v7_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2v7_bench.err
