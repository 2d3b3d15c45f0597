"""GPU probe: full-size model forward + per-op profile (diagnostic; numbers printed here are not bench values)."""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolo2_light_b200 as yb  # noqa: E402
from yolo2_light_b200 import cfgs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="yolov3")
    ap.add_argument("--size", type=int, default=608)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--quantized", type=int, default=0)
    ap.add_argument("--precision", type=int, default=yb.YB_PREC_BF16_TC)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--top", type=int, default=25)
    a = ap.parse_args()
    wd = tempfile.mkdtemp()
    secs = cfgs.MODELS[a.model](a.size, a.size)
    t = time.time()
    cfg = cfgs.write_cfg(secs, os.path.join(wd, "m.cfg"))
    wts = cfgs.write_weights(secs, os.path.join(wd, "m.weights"), seed=1)
    print(f"files {time.time() - t:.1f}s", flush=True)
    t = time.time()
    net = yb.load_network(cfg, wts, batch=a.batch, quantized=a.quantized)
    net.set_precision(a.precision)
    print(f"load {time.time() - t:.1f}s", flush=True)
    x = cfgs.synthetic_images(a.batch, 3, a.size, a.size)
    t = time.time()
    net.predict(x, quantized=bool(a.quantized))
    print(f"first predict (engine build) {time.time() - t:.2f}s launches={net.last_launches()}", flush=True)
    for r in range(a.reps):
        t = time.time()
        net.predict(x, quantized=bool(a.quantized))
        dt = time.time() - t
        print(f"predict {dt * 1e3:.2f} ms  -> {a.batch / dt:.1f} img/s (host in/out included)", flush=True)
    prof = net.profile(quantized=bool(a.quantized))
    tot = sum(p[2] for p in prof)
    print(f"profile total {tot:.3f} ms over {len(prof)} ops -> {a.batch / tot * 1e3:.1f} img/s device-only")
    bykind = {}
    for li, kind, ms in prof:
        bykind[kind] = bykind.get(kind, 0.0) + ms
    for k, v in sorted(bykind.items(), key=lambda kv: -kv[1]):
        print(f"  {k:12s} {v:9.3f} ms {100 * v / tot:5.1f}%")
    for li, kind, ms in sorted(prof, key=lambda p: -p[2])[:a.top]:
        L = net.layer(li) if li >= 0 else {}
        desc = f"{L.get('c')}x{L.get('h')}x{L.get('w')} -> n{L.get('n')} k{L.get('size')} s{L.get('stride')}" if L else ""
        print(f"  L{li:3d} {kind:10s} {ms:8.3f} ms  {desc}")


if __name__ == "__main__":
    main()
