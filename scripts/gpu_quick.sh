#!/bin/bash
# Quick GPU pass: parity tests, short bench, phase clocks; optional ncu of one kernel (KERNEL=regex WEIGHTS=shipped|plain).
TAG=${1:-q}; OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
python bench.py --steps 500 --warmup 20 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
j=json.load(open("$OUT/${TAG}_bench.json"))
print("value", j["value"], "ms/step", j["ms_per_step"], "e2e", j["e2e"]["value"], j["roofline"]["kernels_ms_per_step"])
v=j["variants"]["plain"]; print("plain", v["value"], v["roofline"]["kernels_ms_per_step"])
PY
python scripts/phase_clocks.py > $OUT/${TAG}_phases.log 2>&1; tail -12 $OUT/${TAG}_phases.log
if [ -n "${KERNEL:-}" ]; then
python scripts/profile_run.py --weights ${WEIGHTS:-shipped} --iters 4 > $OUT/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s 2 -c 1 -o $OUT/${TAG}_ncu -f \
    python scripts/profile_run.py --weights ${WEIGHTS:-shipped} --iters 4 > $OUT/${TAG}_ncu.log 2>&1
fi
