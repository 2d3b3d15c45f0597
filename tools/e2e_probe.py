"""Where does the end-to-end path lose time against the device-resident step?  Variants of the pipelined serving loop."""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import yolo2_light_b200 as yb
from yolo2_light_b200 import cfgs

size, batch = 608, 16
secs = cfgs.MODELS["yolov3"](size, size)
wd = tempfile.mkdtemp()
cfg = cfgs.write_cfg(secs, os.path.join(wd, "m.cfg")); wts = cfgs.write_weights(secs, os.path.join(wd, "m.weights"), seed=1)
net = yb.load_network(cfg, wts, batch=batch)
x = torch.from_numpy(cfgs.synthetic_images(batch, 3, size, size)).cuda()
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for _ in range(5): net.forward_device(x.data_ptr(), stream=st.cuda_stream)
torch.cuda.synchronize()
N = 30
t0 = time.perf_counter()
for _ in range(N): net.forward_device(x.data_ptr(), stream=st.cuda_stream)
torch.cuda.synchronize()
print(f"device-resident, wall clock      : {(time.perf_counter()-t0)/N*1e3:.3f} ms/step")
rng = np.random.default_rng(1)
frames = []
for k in range(3):
    pb = yb.PinnedBuffer(batch * size * size * 3, dtype=np.uint8)
    pb.array[:] = rng.integers(0, 256, size=pb.array.size, dtype=np.uint8)
    frames.append(pb.array.reshape(batch, size, size, 3))
def loop(thresh, nms, n):
    infl = []
    for k in range(n):
        if len(infl) == 3: net.collect_detections(infl.pop(0), copy=False)
        infl.append(net.submit_u8(frames[k % 3], thresh, nms, max_rows=4096))
    while infl: net.collect_detections(infl.pop(0), copy=False)
for name, th, nm in (("u8 + detect thresh .56 nms .45", 0.56, 0.45), ("u8 + detect thresh .56 nms 0 ", 0.56, 0.0),
                     ("u8 + detect thresh .999 (no candidates)", 0.999, 0.45)):
    loop(th, nm, 4); torch.cuda.synchronize()
    t0 = time.perf_counter(); loop(th, nm, N); dt = (time.perf_counter() - t0) / N
    print(f"{name:40s}: {dt*1e3:.3f} ms/step")
pinned = [yb.PinnedBuffer(batch * 3 * size * size) for _ in range(3)]
def loop_raw(n):
    infl = []
    for k in range(n):
        if len(infl) == 3: net.collect(infl.pop(0))
        infl.append(net.submit(pinned[k % 3].array))
    while infl: net.collect(infl.pop(0))
loop_raw(4); t0 = time.perf_counter(); loop_raw(N); print(f"{'float in, raw tensors out':40s}: {(time.perf_counter()-t0)/N*1e3:.3f} ms/step")
