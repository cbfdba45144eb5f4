"""Ad-hoc GPU-vs-oracle check used during development (tests/ holds the real parity suite)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from cs_pathplan_b200 import TrajectoryGeneratorTool, MinimumSnapConfig, shipped_config
from oracle import ref, msnap_oracle as mo

def scaled_err(c, cref, T):
    m = c.shape[2]
    pw = T[:, None, None] ** np.arange(m - 1, -1, -1)[None, None, :]
    den = np.max(np.abs(cref) * pw, axis=2, keepdims=True); den[den == 0] = 1
    return np.max(np.abs(c - cref) * pw / den)

def rc_of(cfg):
    return ref.RefConfig(**{k: getattr(cfg, k) for k in ("order","path_weight","vel_zero_weight","V_avg","min_time_s","sample_distance","start_vel","end_vel","start_acc","end_acc")})

tool = TrajectoryGeneratorTool(0)
print("fp64 peak TFLOP/s:", tool.measure_fp64_peak())
bad = 0
def check(path, cfg, sdo=-1.0, vo=-1.0, tag=""):
    global bad
    res = tool.generate_batch(cfg, path, ns=path.shape[0]-1, sample_distance_override=sdo, v_avg_override=vo)
    rw = ref.reweighted_solve(path, rc_of(cfg), vo)
    s_r = ref.generate(path, rc_of(cfg), sdo, vo)
    s_g = res.trajectory(0)
    ce = scaled_err(res.coeff, rw.coeff, rw.time)
    same = s_g.shape == s_r.shape
    se = np.abs(s_g - s_r).max() if same else float("nan")
    okk = same and se < 1e-6 and ce < 1e-8 and res.iters[0] == rw.iters and np.array_equal(res.times, rw.time)
    bad += (not okk)
    print(f"{tag:28s} S {s_g.shape[0]}/{s_r.shape[0]} iters {res.iters[0]}/{rw.iters} md {res.max_dev[0]:.10f}/{rw.max_dev:.10f} vw {res.vw_final[0]:.4g}/{rw.vw_final:.4g} "
          f"Teq {np.array_equal(res.times, rw.time)} coef {ce:.1e} samp {se:.1e} flags {res.flags[0]} stats {res.stats[0]} {'OK' if okk else 'FAIL'}")

check(ref.UAV31_0_ENU, shipped_config(), 300.0, 30.0, "uav31_0 V30")
check(ref.UAV31_0_ENU, shipped_config(), 300.0, 200.0, "uav31_0 V200")
rng = np.random.default_rng(7)
for order in (2, 3, 4, 5):
    for ns in (1, 2, 5, 16, 33):
        P0 = rng.uniform(-100, 100, 3); steps = rng.normal(0, 10, (ns, 3)); path = np.vstack([P0, P0 + np.cumsum(steps, 0)])
        if order == 5 and ns > 5: continue
        for pw, vw in ((0, 0), (1e-7, 0.01), (0.5, 0.3)):
            cfg = MinimumSnapConfig(order=order, path_weight=pw, vel_zero_weight=vw, start_vel=(1, 0.5, -0.2), end_acc=(0.1, 0.2, 0.3))
            check(path, cfg, tag=f"o{order} ns{ns} pw{pw} vw{vw}")
# batch, mixed lengths
lens = [1, 3, 16, 7, 2, 40]
paths = []
for ns in lens:
    P0 = rng.uniform(-100, 100, 3); steps = rng.normal(0, 10, (ns, 3)); paths.append(np.vstack([P0, P0 + np.cumsum(steps, 0)]))
so = np.concatenate([[0], np.cumsum(lens)])
cfg = MinimumSnapConfig(order=4, path_weight=1e-7, vel_zero_weight=0.01)
res = tool.generate_batch(cfg, np.vstack(paths), seg_offset=so)
for b, p in enumerate(paths):
    s_r = ref.generate(p, rc_of(cfg)); rw = ref.reweighted_solve(p, rc_of(cfg))
    s_g = res.trajectory(b); sl = res.segment_slice(b)
    same = s_g.shape == s_r.shape
    print(f"mixed b{b} ns{lens[b]} S {s_g.shape[0]}/{s_r.shape[0]} iters {res.iters[b]}/{rw.iters} coef {scaled_err(res.coeff[sl], rw.coeff, rw.time):.1e} samp {np.abs(s_g - s_r).max() if same else float('nan'):.1e}")
    bad += not (same and np.abs(s_g - s_r).max() < 1e-6)
# single-call API
t0 = time.time(); s = tool.GenerateTrajectoryMatrix(ref.UAV31_0_ENU, shipped_config(), 300.0, 30.0); print("single call", s.shape, f"{(time.time()-t0)*1e3:.2f} ms")
c, md = tool.SolveQPClosedForm(4, paths[2], np.zeros((2, 3)), np.zeros((2, 3)), mo.allocate_time(paths[2], 5.0, 0.1), 1e-7, 0.01, return_max_deviation=True)
cr, mdr = ref.solve_qp(4, paths[2], np.zeros((2, 3)), np.zeros((2, 3)), mo.allocate_time(paths[2], 5.0, 0.1), 1e-7, 0.01)
print("SolveQP", np.abs(c.reshape(cr.shape) - cr).max(), md, mdr)
# quick timing cfg2
B, ns = 4096, 16
rng = np.random.default_rng(1234)
P0 = rng.uniform(-100, 100, (B, 1, 3)); steps = rng.normal(0, 10, (B, ns, 3))
wp = np.concatenate([P0, P0 + np.cumsum(steps, 1)], 1).reshape(-1, 3)
for pw, vw in ((0, 0), (1e-7, 0.01)):
    cfg = MinimumSnapConfig(order=4, path_weight=pw, vel_zero_weight=vw)
    cap = tool.sample_bound(cfg, wp, ns=ns)
    for rep in range(3):
        t0 = time.time(); res = tool.generate_batch(cfg, wp, ns=ns, capacity=cap); dt = time.time() - t0
    print(f"cfg2 pw{pw} vw{vw}: host e2e {dt*1e3:.2f} ms  rows {res.sample_offset[-1]} cap {cap} iters mean {res.iters.mean():.2f} flags {res.flags.max()} launches {tool.launch_count}")
print("FAILURES:", bad)
sys.exit(1 if bad else 0)
