"""INT8 input calibration on the GPU -- the job of `darknet ... -input_calibration N` in the reference
(network_calibrate_cpu, yolov2_forward_network.c:731): prints the `input_calibration = ...` line for the cfg.

  python tools/calibrate.py net.cfg net.weights [--images imgs.npy] [--n 100] [--batch 8]

imgs.npy: float32 [N, 3, H, W] in [0, 1] (the network's input layout); without it, seeded synthetic images are used.
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import yolo2_light_b200 as yb
from yolo2_light_b200 import cfgs

ap = argparse.ArgumentParser()
ap.add_argument("cfg"); ap.add_argument("weights")
ap.add_argument("--images"); ap.add_argument("--n", type=int, default=100); ap.add_argument("--batch", type=int, default=8)
a = ap.parse_args()
net = yb.load_network(a.cfg, a.weights, batch=a.batch)
net.set_precision(yb.YB_PREC_FP32)            # the reference calibrates on its float path
imgs = np.load(a.images).astype(np.float32) if a.images else cfgs.synthetic_images(a.n, net.c, net.h, net.w, seed=1234)
rows = []
for k in range(0, len(imgs) - a.batch + 1, a.batch):
    rows.append(net.calibrate(imgs[k:k + a.batch]))
    print(f"\r{(k + a.batch)}/{len(imgs)} images", end="", file=sys.stderr)
print(file=sys.stderr)
print(yb.api.format_input_calibration(np.concatenate(rows, 0)))
