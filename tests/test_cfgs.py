"""Generated model definitions == the reference's shipped assets, as seen by the reference's own parser; and the
product parser == the reference parser on every generated model."""
import os

import numpy as np
import pytest

import ybtest_util as util
from yolo2_light_b200 import cfgs

REF_BIN = "/root/reference/bin"
ASSETS = [("yolov3", lambda: cfgs.yolov3(416, 416), "yolov3.cfg"), ("spp", cfgs.yolov3_spp, "yolov3-spp.cfg"),
          ("tiny", cfgs.yolov3_tiny, "yolov3-tiny.cfg"), ("xnor", cfgs.tiny_yolo_obj_xnor, "tiny-yolo-obj_xnor.cfg"),
          ("v2voc", cfgs.yolov2_voc, "yolov2-voc.cfg"), ("tinyvoc", cfgs.tiny_yolo_voc, "tiny-yolo-voc.cfg")]
FIELDS = ["type", "activation", "batch_normalize", "h", "w", "c", "n", "size", "stride", "pad", "out_h", "out_w",
          "out_c", "xnor", "quantized", "index", "classes", "coords", "softmax", "total", "reverse", "outputs"]


# keys of the shipped assets that only matter for training / drawing and are not emitted by the generators
TRAINING_KEYS = {"momentum", "decay", "angle", "saturation", "exposure", "hue", "learning_rate", "burn_in",
                 "max_batches", "policy", "steps", "scales", "jitter", "ignore_thresh", "truth_thresh", "random",
                 "rescore", "object_scale", "noobject_scale", "class_scale", "coord_scale", "absolute", "thresh",
                 "bias_match"}


@pytest.mark.skipif(not os.path.isdir(REF_BIN), reason="reference tree absent")
@pytest.mark.parametrize("name,build,asset", ASSETS)
def test_generated_cfg_text_equals_reference_asset(name, build, asset):
    """Section by section, every option the forward path reads has the same value in the generated model and in
    the shipped asset (text level, with the reference's own line grammar)."""
    gen = cfgs.parse_text(cfgs.to_text(build()))
    ref = cfgs.parse_text(open(os.path.join(REF_BIN, asset)).read())
    assert len(gen) == len(ref)
    for i, ((tg, og), (tr, orf)) in enumerate(zip(gen, ref)):
        assert tg == tr, (i, tg, tr)
        keys = (set(og) | set(orf)) - TRAINING_KEYS
        if i == 0:
            keys -= {"batch", "subdivisions"}   # the app always overrides the batch (main.c:160)
        for k in keys:
            vg, vr = og.get(k), orf.get(k)
            if k in ("anchors", "input_calibration", "layers", "mask"):
                vg = [float(t) for t in vg.split(",")]; vr = [float(t) for t in vr.split(",")]
            elif vg is not None and vr is not None and k != "activation":
                vg, vr = float(vg), float(vr)
            assert vg == vr, (name, i, tg, k, vg, vr)


@pytest.mark.skipif(not (os.path.isdir(REF_BIN) and util.have_ref()), reason="reference tree / oracle build absent")
@pytest.mark.parametrize("name,build,asset,qs", [(a[0], a[1], a[2], (0, 1) if a[0] in ("tiny", "xnor") else (1,))
                                                 for a in ASSETS if a[0] in ("tiny", "xnor", "yolov3")])
def test_generated_cfg_equals_reference_asset(name, build, asset, qs, workdir):
    """...and the reference's own parser builds identical networks from both files."""
    from oracle import ref
    p = cfgs.write_cfg(build(), os.path.join(workdir, "gen_" + name + ".cfg"))
    for q in qs:
        a = ref.RefNet(p, None, 1, q, 0)
        b = ref.RefNet(os.path.join(REF_BIN, asset), None, 1, q, 0)
        assert a.n == b.n
        for i in range(a.n):
            la, lb = a.layers[i], b.layers[i]
            for k in la:
                if k != "bflops":
                    assert la[k] == lb[k], (name, i, k, la[k], lb[k])
            if la["type_name"] in ("YOLO", "REGION"):
                na = 2 * (la["total"] if la["type_name"] == "YOLO" else la["n"])
                assert np.array_equal(a.array(i, "biases", na), b.array(i, "biases", na))
            if la["type_name"] == "YOLO":
                assert np.array_equal(a.array(i, "mask", la["n"], np.int32), b.array(i, "mask", la["n"], np.int32))
            if la["type_name"] == "ROUTE":
                assert np.array_equal(a.array(i, "input_layers", la["n"], np.int32),
                                      b.array(i, "input_layers", la["n"], np.int32))
        assert np.array_equal(a.input_calibration(), b.input_calibration())


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", list(util.ZOO) + ["full_tiny", "full_xnor"])
def test_product_parser_equals_reference_parser(name, workdir):
    import yolo2_light_b200 as yb
    from oracle import ref
    if name.startswith("full_"):
        secs = cfgs.yolov3_tiny() if name == "full_tiny" else cfgs.tiny_yolo_obj_xnor()
        cfg = cfgs.write_cfg(secs, os.path.join(workdir, name + ".cfg"))
    else:
        cfg, _ = util.model_files(name, workdir)
    for q in (0, 1):
        a = yb.parse_network_cfg(cfg, 3, q)
        b = ref.RefNet(cfg, None, 3, q, 0)
        assert a.n == b.n and a.batch == b.batch == 3
        assert (a.h, a.w, a.c, a.inputs) == (b.height, b.width, b.channels, b.inputs)
        for i in range(a.n):
            la, lb = a.layer(i), b.layers[i]
            for k in FIELDS:
                assert la[k] == lb[k], (name, q, i, k, la[k], lb[k])
            if lb["type_name"] == "YOLO":
                assert np.array_equal(la["mask"], b.array(i, "mask", lb["n"], np.int32))
                assert np.array_equal(la["anchors"], b.array(i, "biases", 2 * lb["total"]))
            if lb["type_name"] == "REGION":
                assert np.array_equal(la["anchors"][:2 * lb["n"]], b.array(i, "biases", 2 * lb["n"]))
            if lb["type_name"] == "ROUTE":
                assert np.array_equal(la["input_layers"], b.array(i, "input_layers", lb["n"], np.int32))
        assert np.array_equal(a.input_calibration(), b.input_calibration())


def test_shape_tracer_agrees_with_product_parser(workdir):
    import yolo2_light_b200 as yb
    for name in util.ZOO:
        build = util.ZOO[name][0]
        cfg, _ = util.model_files(name, workdir)
        net = yb.parse_network_cfg(cfg, 1, 0)
        shapes = cfgs.conv_shapes(build())
        assert len(shapes) == net.n
        for i, s in enumerate(shapes):
            l = net.layer(i)
            if l["type_name"] in ("CONVOLUTIONAL", "MAXPOOL", "UPSAMPLE", "REORG", "ROUTE", "SHORTCUT"):
                assert (s["out_h"], s["out_w"], s["out_c"]) == (l["out_h"], l["out_w"], l["out_c"]), (name, i)


def test_parser_rejects_bad_input(workdir):
    import yolo2_light_b200 as yb
    with pytest.raises(yb.YbError):
        yb.parse_network_cfg(os.path.join(workdir, "does_not_exist.cfg"), 1, 0)
    p = os.path.join(workdir, "bad.cfg")
    open(p, "w").write("[net]\nwidth=32\nheight=32\nchannels=3\n[convolutional]\nfilters=8\nsize=1\n"
                       "[yolo]\nmask=0\nnum=1\nclasses=80\n")
    with pytest.raises(yb.YbError):   # filters= does not match classes/mask (additionally.c:3656-3660)
        yb.parse_network_cfg(p, 1, 0)
