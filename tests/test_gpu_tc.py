"""Tensor-core (tcgen05) FP32-variant convolution: per-layer against the CPU oracle fed the GPU's own bf16 inputs,
then whole networks against the f32 oracle on the activated detection tensors (north_star: <= 1e-3 rel)."""
import os

import numpy as np
import pytest

import ybtest_util as util
from yolo2_light_b200 import cfgs

pytestmark = pytest.mark.gpu


def bf16_round(a):
    """round-to-nearest-even float32 -> bfloat16 -> float32"""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return (r & 0xFFFFFFFF).astype(np.uint32).view(np.float32).reshape(np.shape(a))


def tcnet(size=64):
    """Exercises every tile configuration of k_conv_tc: BK 16/32/64, BN 32/64/128/256, 3x3 s1, 3x3 s2, 1x1, fused
    shortcut, concat slice output, f32 head with 255 filters."""
    c = cfgs._conv
    s = [cfgs._net(size, size),
         c(16, 3),                   # 0 stem (CUDA cores, C=3)
         c(32, 3, 2),                # 1 s2, BK16, BN32
         c(64, 3),                   # 2 s1, BK32, BN64
         c(32, 1),                   # 3 1x1, BK64, BN32
         c(64, 3),                   # 4 3x3 + fused shortcut
         ("shortcut", {"from": "-3", "activation": "linear"}),   # 5
         c(128, 3, 2),               # 6 s2, BK64, BN128
         c(64, 1),                   # 7
         c(128, 3),                  # 8
         ("shortcut", {"from": "-3", "activation": "linear"}),   # 9
         c(256, 3, 2),               # 10 s2 BN256
         c(128, 1),                  # 11
         c(256, 3),                  # 12
         c(320, 1),                  # 13 two filter tiles (256 + 64)
         c(255, 1, bn=False, act="linear"),   # 14 head, f32 out, n=255
         cfgs._yolo("0,1,2", cfgs.COCO_ANCHORS, 9),              # 15
         ("route", {"layers": "-4"}),                            # 16 -> layer 12
         c(64, 1),                   # 17
         ("upsample", {"stride": "2"}),                          # 18 writes a concat slice
         ("route", {"layers": "-1, 9"}),                         # 19 concat(64 + 128) = 192 channels
         c(128, 3),                  # 20 reads the concat (C=192, BK64)
         c(255, 1, bn=False, act="linear"),   # 21
         cfgs._yolo("3,4,5", cfgs.COCO_ANCHORS, 9)]              # 22
    return s


def _files(workdir, name, secs, seed):
    cfg = os.path.join(workdir, name + ".cfg")
    wts = os.path.join(workdir, name + ".weights")
    cfgs.write_cfg(secs, cfg)
    cfgs.write_weights(secs, wts, seed=seed)
    return cfg, wts


@pytest.mark.parametrize("size,batch", [(64, 2), (96, 3), ((64, 160), 2), ((96, 32), 1)])
def test_tc_every_layer_vs_oracle_on_bf16_inputs(size, batch, workdir):
    import yolo2_light_b200 as yb
    from oracle import port
    h, w = size if isinstance(size, tuple) else (size, size)
    secs = tcnet(64)
    secs[0][1]["height"], secs[0][1]["width"] = str(h), str(w)    # non-square variants exercise H != W tiling
    cfg, wts = _files(workdir, f"tcnet{h}x{w}", secs, 21)
    net = yb.load_network(cfg, wts, batch=batch)
    net.set_precision(yb.YB_PREC_BF16_TC)
    net.set_option("fuse", 0)
    x = cfgs.synthetic_images(batch, 3, h, w, seed=5)
    net.predict(x)
    prof = net.profile()
    kinds = {}
    for li, kind, ms in prof:
        kinds.setdefault(li, []).append(kind)
    layers = net.layers
    got = [None] * net.n
    for i in range(net.n):
        got[i] = net.fetch_layer(i)
    n_tc = 0
    for i, l in enumerate(layers):
        if l["type_name"] != "CONVOLUTIONAL":
            continue
        is_tc = bool({"conv_tc", "conv_tc2"} & set(kinds.get(i, [])))
        # the stem reads the caller's f32 NCHW image directly (tensor-core stem: rounds it to bf16 on the fly)
        xin = (bf16_round(x) if is_tc else x) if i == 0 else got[i - 1]
        n_tc += is_tc
        w = bf16_round(l["weights"]) if is_tc else l["weights"]
        exp = port.conv_fp32(xin, w, l["biases"], l["n"], l["size"], l["stride"], l["pad"], l["activation"])
        head = l["activation"] == 3
        if not head:
            exp = bf16_round(exp)
        err = util.rel_l2(got[i], exp)
        assert err <= 5e-4, (size, i, "tc" if is_tc else "simt", err)
    assert n_tc >= 14, n_tc


def s2net():
    """3x3 / stride-2 layers with 32 and 64 input channels: the parity-halo form of k_conv_tc (TcParams::halo == 2), with the
    resident filter matrix (C = 32) and with streamed filter tiles over two channel blocks (C = 64), BN 64 / 128."""
    c = cfgs._conv
    return [cfgs._net(64, 64),
            c(32, 3),                   # 0 stem
            c(64, 3, 2),                # 1 C=32 s2, filters resident
            c(64, 3, 2),                # 2 C=64 s2, BN 64
            c(128, 3, 2),               # 3 C=64 s2, BN 128
            c(255, 1, bn=False, act="linear"),
            cfgs._yolo("0,1,2", cfgs.COCO_ANCHORS, 9)]


@pytest.mark.parametrize("hw,batch,per_tap", [((64, 64), 2, 0), ((96, 160), 3, 0), ((80, 48), 5, 0), ((608, 32), 1, 0), ((96, 160), 3, 1)])
def test_tc_stride2_parity_halo_vs_oracle(hw, batch, per_tap, workdir, monkeypatch):
    import yolo2_light_b200 as yb
    from oracle import port
    h, w = hw
    secs = s2net()
    secs[0][1]["height"], secs[0][1]["width"] = str(h), str(w)
    cfg, wts = _files(workdir, f"s2net{h}x{w}", secs, 31)
    if per_tap:
        monkeypatch.setenv("YB_TC_S2_HALO_MAXC", "0")   # the one-box-per-tap form of the same layers
    net = yb.load_network(cfg, wts, batch=batch)
    net.set_precision(yb.YB_PREC_BF16_TC)
    net.set_option("fuse", 0)
    x = cfgs.synthetic_images(batch, 3, h, w, seed=9)
    net.predict(x)
    layers = net.layers
    got = [net.fetch_layer(i) for i in range(net.n)]
    for i in (1, 2, 3):
        l = layers[i]
        exp = bf16_round(port.conv_fp32(got[i - 1], bf16_round(l["weights"]), l["biases"], l["n"], l["size"], l["stride"], l["pad"],
                                        l["activation"]))
        err = util.rel_l2(got[i], exp)
        assert err <= 5e-4, (hw, i, err)


def test_tc_fused_equals_unfused(workdir):
    import yolo2_light_b200 as yb
    cfg, wts = _files(workdir, "tcnet64f", tcnet(64), 22)
    x = cfgs.synthetic_images(2, 3, 64, 64, seed=6)
    outs = []
    for fuse in (0, 1):
        net = yb.load_network(cfg, wts, batch=2)
        net.set_option("fuse", fuse)
        net.predict(x)
        outs.append({i: o.copy() for i, o in net.detection_outputs().items()})
        launches = net.last_launches()
        outs.append(launches)
    assert outs[3] < outs[1]
    for i in outs[0]:
        assert util.rel_l2(outs[2][i], outs[0][i]) <= 2e-3, i   # residual add before vs after bf16 rounding


@pytest.mark.parametrize("name", ["tcnet", "v3_32", "spp32", "tiny64"])
def test_bf16_network_vs_f32_oracle(name, workdir):
    """Whole network, default precision, against the f32 oracle: <= 1e-3 rel-L2 on the activated yolo tensors... for
    the slim test nets the bar is 3e-3 (few channels -> less averaging of the bf16 rounding noise); the full-size
    bar is asserted in test_gpu_fullsize.py."""
    import yolo2_light_b200 as yb
    from oracle import port
    if name == "tcnet":
        cfg, wts = _files(workdir, "tcnet64w", tcnet(64), 23)
        x = cfgs.synthetic_images(2, 3, 64, 64, seed=7)
    else:
        cfg, wts = util.model_files(name, workdir)
        x = util.images(name, 2)
    net = yb.load_network(cfg, wts, batch=2)
    net.predict(x)
    layers = net.layers
    exp = [port.run_network(layers, x[b:b + 1]) for b in range(2)]
    for i, o in net.detection_outputs().items():
        e = np.concatenate([exp[b][i] for b in range(2)], 0).reshape(o.shape)
        err = util.rel_l2(o, e)
        assert err <= 3e-3, (name, i, err)


def ksplitnet(h, w):
    """Deep-K layers on small grids: every tensor-core layer has far fewer tiles than the GPU has SMs, so the whole
    layer is one (tail) wave and the K-split schedule cuts each work item into several slices."""
    c = cfgs._conv
    return [cfgs._net(h, w),
            c(32, 3),                   # 0 stem
            c(256, 3, 2),               # 1 s2, K = 9*32
            c(256, 3),                  # 2 CTA pairs (BN 256), K = 9*256: 36 stages
            c(128, 1),                  # 3 1x1, K = 256
            c(256, 3),                  # 4 + fused shortcut, K = 9*128
            ("shortcut", {"from": "-3", "activation": "linear"}),   # 5
            c(512, 3),                  # 6 two filter tiles of 256, CTA pairs
            c(128, 3),                  # 7 BN 128 (single CTAs), K = 9*512: resident-B not possible
            c(64, 3),                   # 8 BN 64, K = 9*128
            c(255, 1, bn=False, act="linear"),   # 9 head (fused yolo), K = 64
            cfgs._yolo("0,1,2", cfgs.COCO_ANCHORS, 9),
            ("route", {"layers": "-4"}),                                # 11 -> layer 7
            c(255, 3, bn=False, act="linear"),   # 12 f32 head with deep K (9*128), generic epilogue
            cfgs._yolo("3,4,5", cfgs.COCO_ANCHORS, 9)]


@pytest.mark.parametrize("h,w,batch", [(32, 32, 2), (64, 32, 3), (96, 96, 4)])
def test_tc_ksplit_tail_matches_plain_schedule(h, w, batch, workdir):
    """K-split tail wave (yb_conv_tc.cu): same numbers as the unsplit schedule up to f32 summation order, on every
    layer, and the layers really are split."""
    import yolo2_light_b200 as yb
    cfg, wts = _files(workdir, f"ksplit{h}x{w}", ksplitnet(h, w), 31)
    x = cfgs.synthetic_images(batch, 3, h, w, seed=9)
    got = []
    for ks in (0, 1):
        net = yb.load_network(cfg, wts, batch=batch)
        net.set_option("ksplit", ks)
        net.set_option("fuse", 0)
        for rep in range(3):     # flags must re-arm between launches
            net.predict(x)
        n_split = net.get_info("ksplit_layers")
        assert net.get_info("tc_layers") >= 9
        assert (n_split >= 3) if ks else (n_split == 0), n_split
        got.append([net.fetch_layer(i) for i in range(net.n)])
    for i, (a, b) in enumerate(zip(*got)):
        # identical bf16 inputs per layer only if the previous layer agreed bit for bit; the drift of a few bf16 ulps
        # accumulates down the (unfused, layer by layer) chain
        assert util.rel_l2(b, a) <= 4e-3, (i, util.rel_l2(b, a))
    # fused engine, CUDA graph replays; the split is opt-in
    net = yb.load_network(cfg, wts, batch=batch)
    assert net.get_info("ksplit_layers") == 0
    net.set_option("ksplit", 1)
    assert net.get_info("ksplit_layers") >= 3
    ref = yb.load_network(cfg, wts, batch=batch)
    for rep in range(3):
        net.predict(x); ref.predict(x)
    for i, o in net.detection_outputs().items():
        assert util.rel_l2(o, ref.detection_outputs()[i]) <= 2e-3, i


@pytest.mark.parametrize("name,q", [("tiny64", 1), ("xnor64", 0), ("tiny_w96_h64", 1)])
def test_tf32_heads_of_exact_networks(name, q, workdir):
    """The float detection heads of the INT8 / XNOR networks run on tcgen05 kind::tf32 in the default precision (their
    result feeds no integer layer); everything upstream stays bit-exact, the detections stay within the FP32 bar of
    the all-f32 engine (north_star: <= 1e-3 rel)."""
    import yolo2_light_b200 as yb
    cfg, wts = util.model_files(name, workdir)
    B = 3
    x = util.images(name, B)
    fast = yb.load_network(cfg, wts, batch=B, quantized=q)
    fast.predict(x, quantized=bool(q))
    kinds = [k for _, k, _ in fast.profile(quantized=bool(q))]
    assert "conv_tc_tf32" in kinds, kinds
    exact = yb.load_network(cfg, wts, batch=B, quantized=q)
    exact.set_precision(yb.YB_PREC_FP32)
    exact.predict(x, quantized=bool(q))
    assert "conv_tc_tf32" not in [k for _, k, _ in exact.profile(quantized=bool(q))]
    for i, o in fast.detection_outputs().items():
        err = util.rel_l2(o, exact.detection_outputs()[i])
        assert err <= 1e-3, (name, i, err)
    # the integer layers did not notice: raw accumulators / counts of the last integer layer are identical
    fast.set_option("keep_counts", 1); exact.set_option("keep_counts", 1)
    fast.predict(x, quantized=bool(q)); exact.predict(x, quantized=bool(q))
    ints = [i for i, l in enumerate(fast.layers) if l["type_name"] == "CONVOLUTIONAL" and (l["xnor"] or (q and i >= 1 and l["activation"] != 3))]
    a = fast.fetch_counts(ints[-1], quantized=bool(q)); b = exact.fetch_counts(ints[-1], quantized=bool(q))
    assert a is not None and np.array_equal(a, b)
