"""Developer tool: where does a fused-kernel CTA spend its time?  (clock64 stamps after every phase barrier)"""
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from cs_pathplan_b200 import TrajectoryGeneratorTool, workloads
names_pw = ["load", "times", "rows1", "thomas1", "search", "probes+rows2", "selinit"]
names_plain = ["load", "times", "rows", "selinit"]
tool = TrajectoryGeneratorTool(0)
for weights in ("shipped", "plain"):
    for B, ns in ((4096, 16), (1 << 15, 8)):
        cfg = workloads.synthetic_config(4, weights)
        wp = workloads.random_walks(B, ns, 1234)
        cap = tool.sample_bound(cfg, wp, ns=ns)
        for _ in range(2): tool.generate_batch(cfg, wp, ns=ns, capacity=cap)
        tool.debug_phase_clocks(True)
        tool.generate_batch(cfg, wp, ns=ns, capacity=cap)
        c = tool.debug_phase_clocks(True, read=True)
        c = c[c[:, 0] > 0]
        names = names_pw if weights == "shipped" else names_plain
        d = np.diff(c[:, : len(names) + 1], axis=1) / 1.9e3   # us at ~1.9 GHz
        print(f"{weights} B={B} ns={ns}: CTAs {len(c)}  span of first tiles {(c[:, len(names)].max() - c[:, 0].min()) / 1.9e3:.1f} us")
        if weights == "shipped":
            for w in range(3):
                f = (c[:, 8 + 3 * w] - c[:, 7 + 3 * w]) / 1.9e3; b = (c[:, 9 + 3 * w] - c[:, 8 + 3 * w]) / 1.9e3
                s0 = (c[:, 7 + 3 * w] - c[:, 6]) / 1.9e3
                print("   spec lane %3d: starts %.1f us after rows2 barrier, forward %.1f us, backward %.1f us" % (64 * w, s0.mean(), f.mean(), b.mean()))
        sc = tool.debug_phase_clocks(True, read=True) if False else c
        full = sc
        rest = np.diff(c, axis=1)[:, len(names):] / 1.9e3
        rest = np.where(rest > 0, rest, 0)
        print("   " + "  ".join(f"{n} {m:.1f}" for n, m in zip(names, d.mean(0))), " | then", np.round(rest.mean(0)[:8], 1))
# sampler kernel phases (stamps in columns 8..13 of rows indexed by tile)
cfg = workloads.synthetic_config(4, "plain"); wp = workloads.random_walks(4096, 16, 1234); cap = tool.sample_bound(cfg, wp, ns=16)
tool.debug_phase_clocks(True); tool.generate_batch(cfg, wp, ns=16, capacity=cap); c = tool.debug_phase_clocks(True, read=True)
c = c[:512, 8:14]; d = np.diff(c, axis=1) / 1.9e3
print("sampler tiles:", len(c), " phases [sort, A count, B traj, C lookback, D write] us mean:", np.round(d.mean(0), 1), " max:", np.round(d.max(0), 1),
      " span:", round((c[:, -1].max() - c[:, 0].min()) / 1.9e3, 1))
tool.debug_phase_clocks(False)
