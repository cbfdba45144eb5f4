import csv, io, shutil, subprocess, sys
tag, ver = sys.argv[1], sys.argv[2]
G, P = "gpurun_out/", "profiles/"
shutil.copy(f"{G}{tag}_alt_bench.json", f"{P}{ver}_alt_bench.json")
shutil.copy(f"{G}{tag}_alt_launches.csv", f"{P}{ver}_alt_launches.csv")
try:
    shutil.copy(f"{G}{tag}_alt_bench_pairs.json", f"{P}{ver}_alt_bench_pairs.json")
except FileNotFoundError:
    pass
WANT = ("GPU Speed Of Light Throughput", "Launch Statistics", "Occupancy", "Compute Workload Analysis", "Scheduler Statistics", "Warp State Statistics")
path = f"{G}{tag}_alt_solve.ncu-rep"
with open(f"{P}{ver}_alt_ncu_summary.txt", "w") as out:
    out.write(f"== k_alt_part, 4096 trajectories x 150-259 rows, shipped altitude parameters  [ncu --set full --clock-control none, {path}]\n")
    det = subprocess.run(["ncu", "-i", path, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    for r in csv.reader(io.StringIO(det)):
        if len(r) > 14 and r[11] in WANT:
            out.write(f"{r[11]:32s} {r[12]:48s} {r[13]:16s} {r[14]}\n")
    raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    d = {h: (u, v) for h, u, v in zip(raw[0], raw[1], raw[-1])}
    for name in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
                 "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                 "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__inst_executed.sum"):
        if name in d:
            out.write(f"raw  {name:72s} {d[name][1]} {d[name][0]}\n")
print(open(f"{P}{ver}_alt_ncu_summary.txt").read()[-900:])
