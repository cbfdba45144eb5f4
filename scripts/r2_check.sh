timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_robustness.py tests/test_gpu_geo.py -m gpu -q -x --tb=short 2>&1 | tail -4
python - <<'PY'
import sys, time; sys.path.insert(0, ".")
import numpy as np, torch
from cs_pathplan_b200 import TrajectoryGeneratorTool, workloads
tool = TrajectoryGeneratorTool(0)
wp, ns = workloads.cfg3(B=1 << 18)
cfg = workloads.synthetic_config(4, "shipped")
free0 = torch.cuda.mem_get_info()[0]
for i in range(2):
    t0 = time.perf_counter(); r = tool.generate_batch(cfg, wp, ns=ns, outputs="samples"); dt = time.perf_counter() - t0
print("host path 262144 x 8: %.1f ms, rows %d, device memory held %.2f GB" % (dt * 1e3, r.samples.shape[0], (free0 - torch.cuda.mem_get_info()[0]) / 1e9))
PY
