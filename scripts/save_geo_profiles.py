"""Copy one gpu_geo.sh pass from gpurun_out/ into profiles/ (geo bench lines, launch list, ncu summary of k_enu_to_wgs84).
Usage: python scripts/save_geo_profiles.py <gpurun tag> <profiles version tag> [<tag of an exact-trig capture>]"""
import csv, io, shutil, subprocess, sys
tag, ver = sys.argv[1], sys.argv[2]
G, P = "gpurun_out/", "profiles/"
shutil.copy(f"{G}{tag}_geo_bench.jsonl", f"{P}{ver}_geo_bench.jsonl")
shutil.copy(f"{G}{tag}_geo_launches.csv", f"{P}{ver}_geo_launches.csv")
WANT = ("GPU Speed Of Light Throughput", "Launch Statistics", "Occupancy", "Memory Workload Analysis",
        "Compute Workload Analysis", "Warp State Statistics", "Scheduler Statistics")
reps = [(f"{G}{tag}_enu_to_wgs84.ncu-rep", "k_enu_to_wgs84<false> (direction-vector form, default)")]
if len(sys.argv) > 3:
    reps.append((f"{G}{sys.argv[3]}_enu_to_wgs84.ncu-rep", "k_enu_to_wgs84 (the reference's statements, first version)"))
with open(f"{P}{ver}_geo_ncu_summary.txt", "w") as out:
    for path, what in reps:
        out.write(f"== {what}, 16 777 216 rows  [ncu --set full --clock-control none, {path}]\n")
        det = subprocess.run(["ncu", "-i", path, "--page", "details", "--csv"], capture_output=True, text=True).stdout
        for r in csv.reader(io.StringIO(det)):
            if len(r) > 14 and r[11] in WANT:
                out.write(f"{r[11]:32s} {r[12]:48s} {r[13]:16s} {r[14]}\n")
        raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
        d = {h: (u, v) for h, u, v in zip(raw[0], raw[1], raw[-1])}
        for name in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
                     "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                     "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
                     "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"):
            if name in d:
                out.write(f"raw  {name:72s} {d[name][1]} {d[name][0]}\n")
        out.write("\n")
print(open(f"{P}{ver}_geo_ncu_summary.txt").read()[-1500:])
