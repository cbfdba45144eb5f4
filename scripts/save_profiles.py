"""Copy one gpu_check.sh pass from gpurun_out/ into profiles/ (bench lines, launch list, ncu summaries) and refresh
profiles/dominant_kernel_traffic.json (DRAM bytes per launch of the dominant kernels, read by bench.py).
Usage: python scripts/save_profiles.py <gpurun tag> <profiles version tag>"""
import csv, io, json, shutil, subprocess, sys
tag, ver = sys.argv[1], sys.argv[2]
G, P = "gpurun_out/", "profiles/"
for src, dst in (("bench.json", "bench.json"), ("bench_ref.json", "bench_reference.json"), ("launches.csv", "launches.csv"),
                 ("clocks.csv", "clocks.csv")):
    shutil.copy(f"{G}{tag}_{src}", f"{P}{ver}_{dst}")
WANT = ("GPU Speed Of Light Throughput", "Launch Statistics", "Occupancy", "Memory Workload Analysis",
        "Compute Workload Analysis", "Warp State Statistics", "Scheduler Statistics")
traffic = {}
with open(f"{P}{ver}_ncu_summary.txt", "w") as out:
    for rep, weights, kernel in (("fused", "shipped", "k_fused_solve"), ("scan", "shipped", "k_sample_scan"),
                                 ("fused_plain", "plain", "k_fused_solve")):
        path = f"{G}{tag}_{rep}.ncu-rep"
        out.write(f"== {kernel} ({weights} weights)  [ncu --set full --clock-control none, {path}]\n")
        det = subprocess.run(["ncu", "-i", path, "--page", "details", "--csv"], capture_output=True, text=True).stdout
        for r in csv.reader(io.StringIO(det)):
            if len(r) > 14 and r[11] in WANT:
                out.write(f"{r[11]:32s} {r[12]:48s} {r[13]:16s} {r[14]}\n")
        raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
        hdr, unit, val = raw[0], raw[1], raw[2]
        d = {h: (u, v) for h, u, v in zip(hdr, unit, val)}
        def nbytes(name):
            u, v = d[name]
            return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        rd, wr = nbytes("dram__bytes_read.sum"), nbytes("dram__bytes_write.sum")
        for name in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
                     "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                     "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
                     "smsp__thread_inst_executed_per_inst_executed.ratio"):
            out.write(f"raw  {name:72s} {d[name][1]} {d[name][0]}\n")
        traffic.setdefault(weights, {})[kernel] = rd + wr
json.dump(traffic, open(f"{P}dominant_kernel_traffic.json", "w"), indent=1)
print(json.dumps(traffic))
