"""Small forward passes covering every kernel family, as a compute-sanitizer target."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import yolo2_light_b200 as yb
from yolo2_light_b200 import cfgs
import test_gpu_tc
wd = tempfile.mkdtemp()
jobs = [("tcnet", test_gpu_tc.tcnet(64), 0), ("tiny", cfgs.slim(cfgs.yolov3_tiny, 2, 64, 64), 1),
        ("xnor", cfgs.slim(cfgs.tiny_yolo_obj_xnor, 2, 64, 64), 0), ("spp", cfgs.slim(cfgs.yolov3_spp, 4, 32, 32), 0),
        # full-width tiny models: fused stem + pool, pool-fused integer epilogues, narrow XNOR layers as +-1 on kind::i8
        ("tiny_full", cfgs.yolov3_tiny(64, 64), 1), ("xnor_full", cfgs.tiny_yolo_obj_xnor(64, 64), 0),
        ("v3_full", cfgs.slim(cfgs.yolov3, 2, 64, 64), 0),
        # stride-2 parity-halo tiles (C = 32 with resident filters, C = 64 over two channel blocks), non-square
        ("s2net", test_gpu_tc.s2net(), 0)]
for name, secs, q in jobs:
    cfg = cfgs.write_cfg(secs, os.path.join(wd, name + ".cfg")); wts = cfgs.write_weights(secs, os.path.join(wd, name + ".weights"), seed=3)
    size = int(secs[0][1]["width"])
    net = yb.load_network(cfg, wts, batch=2, quantized=q)
    x = cfgs.synthetic_images(2, 3, size, size)
    net.predict(x, quantized=bool(q))
    t = net.submit(x, quantized=bool(q)); net.collect(t, quantized=bool(q))
    frames = (x.transpose(0, 2, 3, 1) * 255).astype(np.uint8)
    t = net.submit_u8(frames, 0.3, 0.45, max_rows=512, quantized=bool(q)); net.collect_detections(t, quantized=bool(q))
    print(name, "ok", {i: float(np.abs(o).mean()) for i, o in net.detection_outputs().items()})
