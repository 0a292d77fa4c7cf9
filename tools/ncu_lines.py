"""Aggregate warp-stall samples of an .ncu-rep by CUDA source line:  python tools/ncu_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = ""
hdr = None
agg = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        stall_cols = [(h, i) for i, h in enumerate(r) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < 10 or r[2] != "-":
        continue  # only the per-source-line summary rows (Address == '-')
    try:
        n = int(r[hdr["# Samples"]])
    except ValueError:
        continue
    key = (cur_file, int(r[0]), r[1].strip()[:100])
    st = {h: int(r[i]) for h, i in stall_cols if r[i].isdigit() and int(r[i]) > 0}
    a = agg.setdefault(key, [0, 0, {}])
    a[0] += n
    a[1] += int(r[hdr["Instructions Executed"]] or 0)
    for k, v in st.items():
        a[2][k] = a[2].get(k, 0) + v
tot = sum(v[0] for v in agg.values())
print("total samples", tot)
for (f, ln, src), (n, ex, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    s = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{n:6d} {100 * n / tot:5.1f}%  {ex:9d}  {f}:{ln:<5d} {src:<100s} [{s}]")
