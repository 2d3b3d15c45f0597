"""Per-layer efficiency table from a run_forward --list dump (op index, layer, kind, ms) for yolov3@608 b16."""
import sys
sys.path.insert(0, '.')
from yolo2_light_b200 import cfgs
model, size, batch, path = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
shapes = cfgs.conv_shapes(cfgs.MODELS[model](size, size))
rows = []
for line in open(path):
    t = line.split()
    if len(t) != 4 or not t[0].isdigit():
        continue
    li, kind, ms = int(t[1]), t[2], float(t[3])
    if li < 0 or shapes[li]["type"] != "convolutional":
        rows.append((li, kind, ms, 0, 0, "")); continue
    L = shapes[li]
    fl = 2 * L["n"] * L["size"] ** 2 * L["c"] * L["out_h"] * L["out_w"] * batch
    by = (L["c"] * L["h"] * L["w"] + L["n"] * L["out_h"] * L["out_w"]) * 2 * batch
    rows.append((li, kind, ms, fl / ms / 1e9, by / ms / 1e6, f"{L['c']}x{L['h']} -> {L['n']} k{L['size']}s{L['stride']}"))
tot = sum(r[2] for r in rows)
print(f"total {tot:.3f} ms")
groups = {}
for li, kind, ms, tf, gbs, d in rows:
    groups.setdefault((kind, d), []).append((ms, tf, gbs))
for (kind, d), v in sorted(groups.items(), key=lambda kv: -sum(x[0] for x in kv[1])):
    ms = sum(x[0] for x in v)
    print(f"{kind:10s} {d:28s} x{len(v):2d}  {ms:7.3f} ms ({100*ms/tot:4.1f}%)  {v[0][1]:7.1f} TFLOP/s  {v[0][2]:7.1f} GB/s(alg)")
