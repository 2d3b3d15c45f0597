"""The CPU restatement against golden vectors produced by the reference itself (tests/golden/make_golden.py).
Runs anywhere (no reference tree, no GPU)."""
import os

import numpy as np
import pytest

import ybtest_util as util

CASES = [("tiny64", 0), ("tiny64", 1), ("xnor64", 0), ("v3_32", 0), ("spp32", 0), ("v2voc32", 0), ("tinyvoc64", 1),
         ("v3_32", 1)]


@pytest.mark.parametrize("name,q", CASES)
def test_port_matches_reference_golden(name, q, workdir):
    import yolo2_light_b200 as yb
    from oracle import port
    g = np.load(os.path.join(util.GOLDEN, f"{name}_q{q}.npz"))
    cfg, wts = util.model_files(name, workdir)
    x = util.images(name, 2)
    net = yb.load_network(cfg, wts, batch=1, quantized=q)
    layers = net.layers
    checked = 0
    for b in range(2):
        outs = port.run_network(layers, x[b:b + 1], quantized=bool(q))
        for key in g.files:
            if not key.startswith(f"b{b}_"):
                continue
            i = int(key.split("_")[1].lstrip("outl"))
            ref = g[key]
            assert util.bits_equal(outs[i].reshape(ref.shape), ref), (name, q, key)
            checked += 1
    assert checked == len(g.files)
