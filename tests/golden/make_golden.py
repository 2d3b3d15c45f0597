"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libyolo2ref_scalar.so, built from
/root/reference by oracle/Makefile) on the generated test models.  Run here (the reference tree is absent on the
GPU box):  python tests/golden/make_golden.py

Each file holds, for one (model, rule): the detection-layer outputs (what get_network_boxes reads), every
convolution's output for the first image, and the metadata needed to regenerate the inputs (seeds live in
tests/ybtest_util.ZOO; the .cfg/.weights/images are regenerated deterministically from them).
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ybtest_util as util  # noqa: E402
from oracle import ref  # noqa: E402

CASES = [("tiny64", 0), ("tiny64", 1), ("xnor64", 0), ("v3_32", 0), ("spp32", 0), ("v2voc32", 0), ("tinyvoc64", 1),
         ("v3_32", 1)]


def main():
    wd = tempfile.mkdtemp()
    for name, q in CASES:
        cfg, wts = util.model_files(name, wd)
        batch = 2
        x = util.images(name, batch)
        net = ref.RefNet(cfg, wts, 1, q, 7)   # the reference decodes / quantises batch item 0 only
        arrays = {}
        for b in range(batch):
            net.predict(x[b:b + 1])
            for i, L in enumerate(net.layers):
                if L["type_name"] in ("YOLO", "REGION"):
                    arrays[f"b{b}_out{i}"] = net.output(i).copy()
                elif b == 0 and L["type_name"] in ("CONVOLUTIONAL", "MAXPOOL", "SHORTCUT", "UPSAMPLE", "REORG"):
                    arrays[f"b0_l{i}"] = net.output(i).astype(np.float32).copy()
        out = os.path.join(HERE, f"{name}_q{q}.npz")
        np.savez_compressed(out, **arrays)
        print(out, os.path.getsize(out) // 1024, "KiB", len(arrays), "arrays")


if __name__ == "__main__":
    main()
