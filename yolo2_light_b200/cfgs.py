"""Programmatic generators for the darknet ``.cfg`` model definitions the hot path is measured on,
plus a seeded synthetic ``.weights`` writer.

The reference ships its model definitions as text assets (``bin/yolov3.cfg``, ``bin/yolov3-spp.cfg``,
``bin/yolov3-tiny.cfg``, ``bin/tiny-yolo-obj_xnor.cfg``, ``bin/yolov2-voc.cfg``, ``bin/tiny-yolo-voc.cfg``).
There is no network here and the GPU box has no ``/root/reference``, so the benchmark/test inputs are
*generated* from the compact architecture descriptions below (Darknet-53 = 5 residual stages of
1/2/8/8/4 blocks, three detection heads, ...).  ``tests/test_cfgs.py`` checks -- when the reference tree
is present -- that the reference parser builds layer-for-layer identical networks from its own assets
and from these generated files.

File formats follow the reference loader: ``.cfg`` (additionally.c:3423 read_cfg, :3858 parse_net_options,
:3534 parse_convolutional ...) and ``.weights`` (additionally.c:3459-3529).
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

Section = Tuple[str, Dict[str, str]]

COCO_ANCHORS = "10,13,  16,30,  33,23,  30,61,  62,45,  59,119,  116,90,  156,198,  373,326"
TINY_ANCHORS = "10,14,  23,27,  37,58,  81,82,  135,169,  344,319"
# per-conv input multipliers shipped for the INT8 path (bin/yolov3-tiny.cfg:25, bin/yolov3.cfg:25)
TINY_INPUT_CALIBRATION = [15.7342, 4.41852, 9.17237, 9.70713, 13.1849, 14.9823, 15.1913, 8.62978,
                          15.7353, 15.6297, 15.6939, 15.4093, 15.8055, 16]
V3_INPUT_CALIBRATION = [15.497, 12.537] + [40] * 74


def _net(width: int, height: int, calib: Optional[Sequence[float]] = None) -> Section:
    o = {"batch": "1", "subdivisions": "1", "width": str(width), "height": str(height), "channels": "3"}
    if calib is not None:
        o["input_calibration"] = ", ".join(f"{v:g}" for v in calib)
    return ("net", o)


_WIDTH_DIV = 1   # >1: thin test variants (same topology, filters // div) -- see slim()


def slim(builder, div: int, *args, **kw) -> List[Section]:
    """Build a model with every BN convolution's filter count divided by ``div`` (detection heads keep their
    size).  Same layer graph, ~div^2 fewer weights: used by the CPU tests, where the full-width models only cost
    time (this container page-faults at ~50 MB/s)."""
    global _WIDTH_DIV
    old, _WIDTH_DIV = _WIDTH_DIV, div
    try:
        return builder(*args, **kw)
    finally:
        _WIDTH_DIV = old


def _conv(filters: int, size: int, stride: int = 1, bn: bool = True, act: str = "leaky", **extra) -> Section:
    if bn and _WIDTH_DIV > 1:
        filters = max(filters // _WIDTH_DIV, 4)
    o: Dict[str, str] = {}
    for k, v in extra.items():
        o[k] = str(v)
    if bn:
        o["batch_normalize"] = "1"
    o.update({"filters": str(filters), "size": str(size), "stride": str(stride), "pad": "1", "activation": act})
    return ("convolutional", o)


def _yolo(mask: str, anchors: str, num: int, classes: int = 80) -> Section:
    return ("yolo", {"mask": mask, "anchors": anchors, "classes": str(classes), "num": str(num)})


def _darknet53() -> List[Section]:
    s: List[Section] = [_conv(32, 3)]
    for filters, blocks in ((64, 1), (128, 2), (256, 8), (512, 8), (1024, 4)):
        s.append(_conv(filters, 3, 2))
        for _ in range(blocks):
            s += [_conv(filters // 2, 1), _conv(filters, 3), ("shortcut", {"from": "-3", "activation": "linear"})]
    return s


def _v3_head(filters: int, mask: str, classes: int, first: bool, spp: bool = False) -> List[Section]:
    s: List[Section] = []
    if spp:
        s += [_conv(512, 1), _conv(1024, 3), _conv(512, 1),
              ("maxpool", {"stride": "1", "size": "5"}), ("route", {"layers": "-2"}),
              ("maxpool", {"stride": "1", "size": "9"}), ("route", {"layers": "-4"}),
              ("maxpool", {"stride": "1", "size": "13"}), ("route", {"layers": "-1,-3,-5,-6"}),
              _conv(512, 1), _conv(1024, 3), _conv(512, 1), _conv(1024, 3)]
    else:
        for _ in range(3):
            s += [_conv(filters, 1), _conv(filters * 2, 3)]
    s += [_conv(3 * (5 + classes), 1, bn=False, act="linear"), _yolo(mask, COCO_ANCHORS, 9, classes)]
    return s


def yolov3(width: int = 416, height: int = 416, classes: int = 80, spp: bool = False) -> List[Section]:
    """yolov3.cfg (107 layers) / yolov3-spp.cfg (114 layers)."""
    s = [_net(width, height, V3_INPUT_CALIBRATION)] + _darknet53()
    s += _v3_head(512, "6,7,8", classes, True, spp)
    s += [("route", {"layers": "-4"}), _conv(256, 1), ("upsample", {"stride": "2"}), ("route", {"layers": "-1, 61"})]
    s += _v3_head(256, "3,4,5", classes, False)
    s += [("route", {"layers": "-4"}), _conv(128, 1), ("upsample", {"stride": "2"}), ("route", {"layers": "-1, 36"})]
    s += _v3_head(128, "0,1,2", classes, False)
    return s


def yolov3_spp(width: int = 608, height: int = 608, classes: int = 80) -> List[Section]:
    return yolov3(width, height, classes, spp=True)


def _tiny_backbone(xnor: bool) -> List[Section]:
    s: List[Section] = []
    for i, f in enumerate((16, 32, 64, 128, 256, 512)):
        extra = {"xnor": 1, "bin_output": 1} if (xnor and i > 0) else {}
        s.append(_conv(f, 3, **extra))
        s.append(("maxpool", {"size": "2", "stride": "2" if f < 512 else "1"}))
    return s


def yolov3_tiny(width: int = 416, height: int = 416, classes: int = 80) -> List[Section]:
    """yolov3-tiny.cfg (24 layers)."""
    s = [_net(width, height, TINY_INPUT_CALIBRATION)] + _tiny_backbone(False)
    s += [_conv(1024, 3), _conv(256, 1), _conv(512, 3), _conv(3 * (5 + classes), 1, bn=False, act="linear"),
          _yolo("3,4,5", TINY_ANCHORS, 6, classes),
          ("route", {"layers": "-4"}), _conv(128, 1), ("upsample", {"stride": "2"}), ("route", {"layers": "-1, 8"}),
          _conv(256, 3), _conv(3 * (5 + classes), 1, bn=False, act="linear"), _yolo("1,2,3", TINY_ANCHORS, 6, classes)]
    return s


def _region(anchors: str, classes: int, num: int = 5) -> Section:
    return ("region", {"anchors": anchors, "bias_match": "1", "classes": str(classes), "coords": "4",
                       "num": str(num), "softmax": "1"})


def tiny_yolo_obj_xnor(width: int = 416, height: int = 416) -> List[Section]:
    """tiny-yolo-obj_xnor.cfg (16 layers): FP32 stem, 7 XNOR 3x3 convs, FP32 1x1 head, region."""
    s = [_net(width, height)] + _tiny_backbone(True)
    s += [_conv(1024, 3, xnor=1, bin_output=1), _conv(1024, 3, xnor=1), _conv(55, 1, bn=False, act="linear"),
          _region("5.2367,6.0570, 8.2272,9.1483, 12.4093,10.7904, 9.7655,14.6023, 16.6749,16.0784", 6)]
    return s


def tiny_yolo_voc(width: int = 416, height: int = 416) -> List[Section]:
    """tiny-yolo-voc.cfg (16 layers, YOLO v2)."""
    calib = [127, 3.88677, 10.5828, 10.3276, 14.3403, 15.2774, 15.2242, 8.08196, 15.7327, 16]
    s = [_net(width, height, calib)] + _tiny_backbone(False)
    s += [_conv(1024, 3), _conv(1024, 3), _conv(125, 1, bn=False, act="linear"),
          _region("1.08,1.19,  3.42,4.41,  6.63,11.38,  9.42,5.11,  16.62,10.52", 20)]
    return s


def yolov2_voc(width: int = 416, height: int = 416) -> List[Section]:
    """yolov2-voc.cfg (32 layers, Darknet-19 + reorg passthrough)."""
    calib = [15.8025, 11.6111, 10.9857, 14.9883, 11.6514, 14.9023, 15.4301, 13.8702, 15.3739, 15.584, 15.3044,
             15.4963, 15.4139, 15.398, 15.7311, 15.2932, 15.7355, 15.2879, 5.79389, 15.6349, 15.5533, 15.453,
             15.7935, 16]
    mp: Section = ("maxpool", {"size": "2", "stride": "2"})
    s: List[Section] = [_net(width, height, calib), _conv(32, 3), mp, _conv(64, 3), mp]
    for f, reps in ((128, 1), (256, 1)):
        s += [_conv(f, 3), _conv(f // 2, 1), _conv(f, 3), mp]
    s += [_conv(512, 3), _conv(256, 1), _conv(512, 3), _conv(256, 1), _conv(512, 3), mp]
    s += [_conv(1024, 3), _conv(512, 1), _conv(1024, 3), _conv(512, 1), _conv(1024, 3), _conv(1024, 3), _conv(1024, 3)]
    s += [("route", {"layers": "-9"}), _conv(64, 1), ("reorg", {"stride": "2"}), ("route", {"layers": "-1,-4"}),
          _conv(1024, 3), _conv(125, 1, bn=False, act="linear"),
          _region("1.3221, 1.73145, 3.19275, 4.00944, 5.05587, 8.09892, 9.47112, 4.84053, 11.2364, 10.0071", 20)]
    return s


MODELS = {
    "yolov3": yolov3,
    "yolov3-spp": yolov3_spp,
    "yolov3-tiny": yolov3_tiny,
    "tiny-yolo-obj_xnor": tiny_yolo_obj_xnor,
    "tiny-yolo-voc": tiny_yolo_voc,
    "yolov2-voc": yolov2_voc,
}


def to_text(sections: Sequence[Section]) -> str:
    out = []
    for name, opts in sections:
        out.append(f"[{name}]")
        out += [f"{k}={v}" for k, v in opts.items()]
        out.append("")
    return "\n".join(out)


def write_cfg(sections: Sequence[Section], path: str) -> str:
    with open(path, "w") as f:
        f.write(to_text(sections))
    return path


def parse_text(text: str) -> List[Section]:
    """INI-like reader with the reference's rules (additionally.c:3423-3457): '#', ';' comment lines,
    whitespace stripped everywhere, key=value split at the first '='."""
    secs: List[Section] = []
    for raw in text.splitlines():
        line = "".join(raw.split())
        if not line or line[0] in "#;":
            continue
        if line[0] == "[":
            secs.append((line.strip("[]"), {}))
        else:
            k, _, v = line.partition("=")
            secs[-1][1][k] = v
    return secs


def conv_shapes(sections: Sequence[Section]) -> List[dict]:
    """Trace tensor shapes through a section list; returns one dict per layer (type, c,h,w in, out_c,out_h,out_w,
    and for convs n,size,stride,pad,bn).  Mirrors the size rules of the reference constructors
    (additionally.c:2699-2706 conv, :2611-2612 maxpool, :3766-3810 route, reorg/upsample)."""
    net = sections[0][1]
    h, w, c = int(net["height"]), int(net["width"]), int(net.get("channels", "3"))
    layers: List[dict] = []
    for idx, (name, o) in enumerate(sections[1:]):
        L = {"type": name, "h": h, "w": w, "c": c}
        if name in ("convolutional", "conv"):
            n, size, stride = int(o.get("filters", 1)), int(o.get("size", 1)), int(o.get("stride", 1))
            pad = size // 2 if int(o.get("pad", 0)) else int(o.get("padding", 0))
            L.update(n=n, size=size, stride=stride, pad=pad, bn=int(o.get("batch_normalize", 0)),
                     xnor=int(o.get("xnor", 0)), activation=o.get("activation", "logistic"))
            h, w, c = (h + 2 * pad - size) // stride + 1, (w + 2 * pad - size) // stride + 1, n
        elif name in ("maxpool", "max"):
            stride = int(o.get("stride", 1)); size = int(o.get("size", stride)); pad = int(o.get("padding", size - 1))
            L.update(size=size, stride=stride, pad=pad)
            h, w = (h + pad - size) // stride + 1, (w + pad - size) // stride + 1
        elif name == "route":
            ids = [int(t) for t in o["layers"].split(",")]
            ids = [i + idx if i < 0 else i for i in ids]
            L.update(layers=ids)
            h, w = layers[ids[0]]["out_h"], layers[ids[0]]["out_w"]
            c = sum(layers[i]["out_c"] for i in ids)
        elif name == "upsample":
            s = int(o.get("stride", 2)); L.update(stride=s); h, w = h * s, w * s
        elif name == "reorg":
            s = int(o.get("stride", 1)); L.update(stride=s); h, w, c = h // s, w // s, c * s * s
        elif name == "shortcut":
            f = int(o["from"]); L.update(index=f + idx if f < 0 else f)
        L.update(out_h=h, out_w=w, out_c=c)
        layers.append(L)
    return layers


def write_weights(sections: Sequence[Section], path: str, seed: int = 0) -> str:
    """Seeded synthetic ``.weights`` in the reference's on-disk format (additionally.c:3459-3529):
    header ``int32 major=0, minor=2, revision=0`` + ``uint64 seen``, then per CONVOLUTIONAL layer in cfg order
    ``biases[n]``, (if batch_normalize) ``scales[n], rolling_mean[n], rolling_variance[n]``, ``weights[n*c*k*k]``.

    Distributions per SURVEY 8(d): conv weights ``sqrt(2/(k*k*c)) * U(-1,1)`` (the reference's own init,
    additionally.c:2751), biases ``U(-0.1,0.1)``, scales ``U(0.5,1.5)``, rolling_mean ``U(-0.1,0.1)``,
    rolling_variance ``U(0.5,1.5)`` -- sane BN statistics, because the parser's defaults (variance 0) blow up
    under yolov2_fuse_conv_batchnorm."""
    rng = np.random.default_rng(seed)
    with open(path, "wb") as f:
        f.write(struct.pack("<iiiQ", 0, 2, 0, 0))
        for L in conv_shapes(sections):
            if L["type"] not in ("convolutional", "conv"):
                continue
            n, c, k = L["n"], L["c"], L["size"]
            rng.uniform(-0.1, 0.1, n).astype("<f4").tofile(f)
            if L["bn"]:
                rng.uniform(0.5, 1.5, n).astype("<f4").tofile(f)
                rng.uniform(-0.1, 0.1, n).astype("<f4").tofile(f)
                rng.uniform(0.5, 1.5, n).astype("<f4").tofile(f)
            scale = np.sqrt(2.0 / (k * k * c))
            (scale * rng.uniform(-1.0, 1.0, n * c * k * k)).astype("<f4").tofile(f)
    return path


def synthetic_images(batch: int, c: int, h: int, w: int, seed: int = 1234) -> np.ndarray:
    """``float32[batch, c, h, w]`` i.i.d. U[0,1) -- the reference's input contract is planar RGB in [0,1]
    (additionally.c:3093-3103).  Image ``i`` depends only on ``seed + i`` so shards agree across ranks."""
    out = np.empty((batch, c, h, w), dtype=np.float32)
    for i in range(batch):
        out[i] = np.random.default_rng(seed + i).random((c, h, w), dtype=np.float32)
    return out
