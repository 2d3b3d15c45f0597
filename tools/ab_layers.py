"""Side-by-side per-op times of two run_forward --list dumps (A/B of a switch)."""
import sys
sys.path.insert(0, '.')
from yolo2_light_b200 import cfgs
model, size, a, b = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
shapes = cfgs.conv_shapes(cfgs.MODELS[model](size, size))
def rd(p):
    r = []
    for line in open(p):
        t = line.split()
        if len(t) == 4 and t[0].isdigit():
            r.append((int(t[1]), t[2], float(t[3])))
    return r
A, B = rd(a), rd(b)
ta = tb = 0
for (li, k, ma), (_, k2, mb) in zip(A, B):
    ta += ma; tb += mb
    d = ""
    if li >= 0 and shapes[li]["type"] == "convolutional":
        L = shapes[li]; d = f"{L['c']}x{L['h']} -> {L['n']} k{L['size']}s{L['stride']}"
    flag = "  <--" if abs(ma - mb) > 0.1 * max(ma, mb) and max(ma, mb) > 0.01 else ""
    print(f"{li:4d} {k:10s} {d:26s} {ma:8.4f} {mb:8.4f} {mb/ma if ma else 0:6.2f}{flag}")
print(f"total {ta:.3f} {tb:.3f}")
