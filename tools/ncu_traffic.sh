#!/bin/bash
# DRAM traffic of every k_conv_tc launch of one eager forward (yolov3 608 b16) -> profiles/r02_traffic.json
set -e
cd "$(dirname "$0")/.."
export YB_NO_GRAPH=1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_conv_tc \
    --csv --log-file gpurun_out/traffic_conv_tc.csv python tools/run_forward.py --reps 1 > /dev/null 2>&1
python - <<'PY'
import csv, json
rows = list(csv.reader(open("gpurun_out/traffic_conv_tc.csv")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]; mn, mu, mv, idc = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
kn = hdr.index("Kernel Name")
tot, ids = 0.0, set()
for r in rows[hi + 1:]:
    nm = r[kn].replace("(int)", "").replace("(bool)", "").replace(" ", "") if len(r) > kn else ""
    if len(r) <= mv or not r[mn].startswith("dram__bytes") or "k_conv_tc<2" not in nm:
        continue
    tot += float(r[mv].replace(",", "")) * scale.get(r[mu], 1); ids.add(r[idc])
out = {"yolov3-608-fp32-b16": {"dram_bytes_per_launch": tot / max(len(ids), 1), "launches": len(ids), "total_bytes": tot,
                               "how": "ncu dram__bytes_read.sum + dram__bytes_write.sum over every k_conv_tc<2> launch (the dominant kernel) of one eager forward"}}
json.dump(out, open("profiles/r02_traffic.json", "w"), indent=1); json.dump(out, open("gpurun_out/r02_traffic.json", "w"), indent=1)
print(out)
PY
