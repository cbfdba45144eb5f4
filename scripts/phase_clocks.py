"""Developer tool: where do the fused solve kernel and the sampler kernel spend their time?  (clock64 stamps)"""
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from cs_pathplan_b200 import TrajectoryGeneratorTool, workloads
GHZ = 1.9e3  # cycles per microsecond at the clock the stamps were taken (approximate)
names_pw = ["load", "times", "rows1", "thomas1", "search", "probes+rows2", "spec+select", "outputs+coeff"]
names_plain = ["load", "times", "rows", "solve+select", "outputs+coeff"]
scan_names = ["sort+count", "rows", "scan+expand+lookback", "write"]
tool = TrajectoryGeneratorTool(0)
tool.set_host_chunks(1)  # the stamps are taken on this handle's own launches
cases = [(4096, 16), (1 << 15, 8)] if len(sys.argv) < 3 else [(int(sys.argv[1]), int(sys.argv[2]))]
for weights in ("shipped", "plain"):
    for B, ns in cases:
        cfg = workloads.synthetic_config(4, weights)
        wp = workloads.random_walks(B, ns, 1234)
        cap = tool.sample_bound(cfg, wp, ns=ns)
        for _ in range(2): tool.generate_batch(cfg, wp, ns=ns, capacity=cap)
        tool.debug_phase_clocks(True)
        tool.generate_batch(cfg, wp, ns=ns, capacity=cap)
        c = tool.debug_phase_clocks(True, read=True)
        f = c[:4096]; f = f[f[:, 0] > 0]
        names = names_pw if weights == "shipped" else names_plain
        d = np.diff(f[:, : len(names) + 1], axis=1) / GHZ
        print(f"{weights} B={B} ns={ns}: fused CTAs {len(f)}, first tile {((f[:, len(names)] - f[:, 0]) / GHZ).mean():.1f} us mean, "
              f"span {(f[:, len(names)].max() - f[:, 0].min()) / GHZ:.1f} us")
        print("   " + "  ".join(f"{n} {m:.1f}" for n, m in zip(names, d.mean(0))))
        tot = (f[:, len(names)] - f[:, 0]) / GHZ
        print("   per-CTA tile time: min %.1f  p50 %.1f  p90 %.1f  max %.1f us;  phase maxima: " % (tot.min(), np.median(tot), np.quantile(tot, 0.9), tot.max())
              + "  ".join(f"{n} {m:.1f}" for n, m in zip(names, d.max(0))))
        if weights == "shipped":
            pe = f[:, 9]
            okp = pe > 0
            if okp.any():
                print("   last-iteration lane pair (warp 0) ends %.1f us after the rows2 barrier (max %.1f)" %
                      (((pe - f[:, 6])[okp] / GHZ).mean(), ((pe - f[:, 6])[okp] / GHZ).max()))
            if "--warp-ends" in sys.argv:
                for w in range(1, 7):
                    we = f[:, 9 + w]
                    ok = we > 0
                    if ok.any():
                        print("   warp %d leaves the phase %.1f us after the rows2 barrier (p90 %.1f, max %.1f)" %
                              (w, ((we - f[:, 6])[ok] / GHZ).mean(), np.quantile((we - f[:, 6])[ok] / GHZ, 0.9), ((we - f[:, 6])[ok] / GHZ).max()))
            for w in range(0 if "--warp-ends" in sys.argv else 3):
                fe, be = f[:, 10 + 2 * w], f[:, 11 + 2 * w]
                ok = fe > 0
                if ok.any():
                    print("   spec lane %3d: forward ends %.1f us after the rows2 barrier, backward takes %.1f us" %
                          (64 * w, ((fe - f[:, 6])[ok] / GHZ).mean(), ((be - fe)[ok] / GHZ).mean()))
        s = c[4096:]; s = s[s[:, 0] > 0][:, :5]
        ds = np.diff(s, axis=1) / GHZ
        print(f"   sampler tiles {len(s)}: " + "  ".join(f"{n} {m:.1f} (max {x:.1f})" for n, m, x in zip(scan_names, ds.mean(0), ds.max(0))) +
              f"  span {(s[:, -1].max() - s[:, 0].min()) / GHZ:.1f} us")
tool.debug_phase_clocks(False)
