"""Warp-stall samples of a .ncu-rep (source page) summed by reason, optionally for a range of SASS rows."""
import csv, io, subprocess, sys
path = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
cols = [(j, h.replace("stall_", "")) for j, h in enumerate(hdr) if h.startswith("stall_") and "(Not" not in h]
tot = {}
for i, r in enumerate(data):
    if not (lo <= i < hi):
        continue
    for j, n in cols:
        if r[j] not in ("", "0"):
            tot[n] = tot.get(n, 0) + int(r[j])
s = sum(tot.values())
for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{n:24s} {v:6d} {100*v/s:5.1f}%")
