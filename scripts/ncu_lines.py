"""Aggregate `ncu --page source --print-source cuda,sass --csv` output: stall samples per CUDA source line."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
fname = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        try: samples = int(r[4])
        except ValueError: samples = 0
        try: inst = int(r[7])
        except ValueError: inst = 0
        # stall columns
        st = {h: int(v) for h, v in zip(hdr, r) if h.startswith("stall_") and "Not Issued" not in h and v.isdigit() and int(v) > 0}
        out.append((samples, inst, fname, int(r[0]), r[1].strip()[:90], st))
tot = sum(o[0] for o in out)
print("total samples", tot, "total warp-inst", sum(o[1] for o in out))
byfile = collections.Counter()
for o in out: byfile[o[2]] += o[0]
print(dict(byfile))
for o in sorted(out, key=lambda o: -o[0])[:top]:
    st = sorted(o[5].items(), key=lambda kv: -kv[1])[:3]
    print(f"{o[0]:6d} {100*o[0]/max(tot,1):5.1f}% inst={o[1]:8d} {o[2]}:{o[3]:<4d} {o[4]:90s} {st}")
