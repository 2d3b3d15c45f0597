"""Hottest SASS instructions (by warp-stall samples) of a .ncu-rep with source info."""
import csv, io, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
ia = hdr.index("Source"); isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
tot = sum(int(r[isamp] or 0) for r in data)
print("total samples", tot)
idx = sorted(range(len(data)), key=lambda i: -int(data[i][isamp] or 0))[:top]
for i in sorted(idx):
    r = data[i]
    st = {h.replace("stall_", ""): r[j] for j, h in enumerate(hdr) if h.startswith("stall_") and "(Not" not in h and r[j] not in ("0", "")}
    print(f"{i:5d} {r[isamp]:>6s} {r[iex]:>8s}  {r[ia].strip()[:64]:64s} {st}")
