"""Host-side model preparation of the product (loader, BN fold, XNOR statistics, INT8 quantisation) against the
reference's own functions, bit-for-bit (main.c:160-171 sequence)."""
import numpy as np
import pytest

import ybtest_util as util

pytestmark = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")


@pytest.mark.parametrize("name,q", [("tiny64", 1), ("xnor64", 0), ("v3_32", 1), ("v2voc32", 1), ("tinyvoc64", 1),
                                    ("spp32", 0)])
def test_prepared_arrays_bit_exact(name, q, workdir):
    import yolo2_light_b200 as yb
    from oracle import ref
    cfg, wts = util.model_files(name, workdir)
    a = yb.load_network(cfg, wts, batch=1, quantized=q)
    b = ref.RefNet(cfg, wts, 1, q, 7)
    nconv = 0
    for i in range(a.n):
        la, lb = a.layer(i), b.layers[i]
        if lb["type_name"] != "CONVOLUTIONAL":
            continue
        nconv += 1
        nw = lb["n"] * lb["c"] * lb["size"] ** 2
        assert la["batch_normalize"] == lb["batch_normalize"] == 0
        assert util.bits_equal(la["weights"], b.array(i, "weights", nw)), (i, "weights")
        assert util.bits_equal(la["biases"], b.array(i, "biases", lb["n"])), (i, "biases")
        if q:
            assert np.array_equal(la["weights_int8"], b.array(i, "weights_int8", nw, np.int8)), (i, "int8")
            assert la["weights_quant_multipler"] == lb["weights_quant_multipler"], i
            assert la["input_quant_multipler"] == lb["input_quant_multipler"], i
        if lb["xnor"]:
            assert util.bits_equal(la["mean_arr"], b.array(i, "mean_arr", lb["n"])), (i, "mean_arr")
    assert nconv > 0


def test_unprepared_weights_equal_file(workdir):
    """load_weights_upto_cpu alone (no fold): arrays are the file contents in cfg order (additionally.c:3459-3468)."""
    import yolo2_light_b200 as yb
    from oracle import ref
    cfg, wts = util.model_files("tiny64", workdir)
    a = yb.parse_network_cfg(cfg, 1, 0)
    yb.load_weights_upto_cpu(a, wts)
    b = ref.RefNet(cfg, wts, 1, 0, 0)
    for i in range(a.n):
        la, lb = a.layer(i), b.layers[i]
        if lb["type_name"] != "CONVOLUTIONAL":
            continue
        for arr in ("weights", "biases", "scales", "rolling_mean", "rolling_variance"):
            cnt = lb["n"] * lb["c"] * lb["size"] ** 2 if arr == "weights" else lb["n"]
            rb = b.array(i, arr, cnt)
            if rb is None:
                assert la[arr] is None
            else:
                assert util.bits_equal(la[arr], rb), (i, arr)


def test_cutoff_and_short_file(workdir):
    """cutoff stops loading after `cutoff` layers; a truncated file is not an error (the reference ignores short
    reads, additionally.c:3459-3468)."""
    import os
    import yolo2_light_b200 as yb
    cfg, wts = util.model_files("tiny64", workdir)
    a = yb.parse_network_cfg(cfg, 1, 0)
    yb.load_weights_upto_cpu(a, wts, cutoff=1)
    assert np.abs(a.layer(0)["weights"]).max() > 0
    assert np.abs(a.layer(2)["weights"]).max() == 0
    short = os.path.join(workdir, "short.weights")
    data = open(wts, "rb").read()
    open(short, "wb").write(data[:len(data) // 2])
    b = yb.parse_network_cfg(cfg, 1, 0)
    yb.load_weights_upto_cpu(b, short)
    assert np.abs(b.layer(0)["weights"]).max() > 0
    with pytest.raises(yb.YbError):
        yb.load_weights_upto_cpu(b, os.path.join(workdir, "missing.weights"))
