"""Full-size BASELINE.json configurations on the GPU against the reference build shipped in oracle/_ref (the GPU box
has no /root/reference; the prebuilt .so travels with the snapshot), plus size-independent properties."""
import os

import numpy as np
import pytest

import ybtest_util as util
from yolo2_light_b200 import cfgs

pytestmark = pytest.mark.gpu


def _files(workdir, name, secs, seed=1):
    cfg = os.path.join(workdir, name + ".cfg")
    wts = os.path.join(workdir, name + ".weights")
    if not os.path.exists(cfg):
        cfgs.write_cfg(secs, cfg)
        cfgs.write_weights(secs, wts, seed=seed)
    return cfg, wts


def _ref_outputs(cfg, wts, x, q, kind):
    from oracle import ref
    os.environ.setdefault("OMP_NUM_THREADS", str(min(os.cpu_count() or 1, 32)))
    rnet = ref.RefNet(cfg, wts, 1, q, 7, kind=kind)
    outs = []
    for b in range(x.shape[0]):
        rnet.predict(x[b:b + 1])
        outs.append({i: rnet.output(i).copy() for i, L in enumerate(rnet.layers) if L["type_name"] in ("YOLO", "REGION")})
    return outs


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_yolov3_608_bf16_tensor_core_vs_reference(workdir):
    """BASELINE configs[1] (batch reduced to 2 for the CPU side): FP32 detections <= 1e-3 rel (rel-L2 on the
    activated yolo tensors) against the reference CPU path on the same weights and images."""
    import yolo2_light_b200 as yb
    secs = cfgs.yolov3(608, 608)
    cfg, wts = _files(workdir, "yolov3_608", secs)
    x = cfgs.synthetic_images(2, 3, 608, 608)
    net = yb.load_network(cfg, wts, batch=2)
    net.predict(x)
    exp = _ref_outputs(cfg, wts, x, 0, "fast")
    for i, o in net.detection_outputs().items():
        for b in range(2):
            err = util.rel_l2(o[b], exp[b][i].reshape(o[b].shape))
            assert err <= 1e-3, (i, b, err)
    prof = net.profile()
    assert sum(1 for _, k, _ in prof if k in ("conv_tc", "conv_tc2")) >= 70
    assert sum(1 for _, k, _ in prof if k == "conv_tc2") >= 30   # CTA-pair kernel carries the wide layers


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_yolov3_tiny_416_fp32_and_int8_vs_reference(workdir):
    import yolo2_light_b200 as yb
    secs = cfgs.yolov3_tiny(416, 416)
    cfg, wts = _files(workdir, "tiny_416", secs)
    x = cfgs.synthetic_images(2, 3, 416, 416)
    net = yb.load_network(cfg, wts, batch=2)
    net.predict(x)
    exp = _ref_outputs(cfg, wts, x, 0, "fast")
    for i, o in net.detection_outputs().items():
        for b in range(2):
            assert util.rel_l2(o[b], exp[b][i].reshape(o[b].shape)) <= 1e-3, (i, b)
    netq = yb.load_network(cfg, wts, batch=2, quantized=1)
    netq.predict(x, quantized=True)
    kinds = [k for _, k, _ in netq.profile(quantized=True)]
    assert kinds.count("conv_tc_i8") >= 8, kinds   # the s8 x s8 -> s32 tcgen05 path carries the INT8 layers
    expq = _ref_outputs(cfg, wts, x, 1, "scalar")
    for i, o in netq.detection_outputs().items():
        for b in range(2):
            assert util.rel_l2(o[b], expq[b][i].reshape(o[b].shape)) <= 2e-3, (i, b)


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_xnor_416_vs_reference(workdir):
    import yolo2_light_b200 as yb
    secs = cfgs.tiny_yolo_obj_xnor(416, 416)
    cfg, wts = _files(workdir, "xnor_416", secs, seed=2)
    x = cfgs.synthetic_images(2, 3, 416, 416)
    net = yb.load_network(cfg, wts, batch=2)
    net.predict(x)
    exp = _ref_outputs(cfg, wts, x, 0, "scalar")
    for i, o in net.detection_outputs().items():
        for b in range(2):
            assert util.rel_l2(o[b], exp[b][i].reshape(o[b].shape)) <= 2e-3, (i, b)


def test_batch_invariance_and_determinism_at_full_size(workdir):
    """Images are independent: image k of a batch of 16 == the same image run in a batch of 1 (bit-for-bit: the
    kernels' reduction order does not depend on the batch), and two runs of the same batch are identical."""
    import yolo2_light_b200 as yb
    secs = cfgs.yolov3(608, 608)
    cfg, wts = _files(workdir, "yolov3_608", secs)
    x = cfgs.synthetic_images(16, 3, 608, 608)
    net = yb.load_network(cfg, wts, batch=16)
    net.predict(x)
    a = {i: o.copy() for i, o in net.detection_outputs().items()}
    net.predict(x)
    for i, o in net.detection_outputs().items():
        assert util.bits_equal(o, a[i])
        assert np.isfinite(o).all()
    one = yb.load_network(cfg, wts, batch=1)
    for k in (0, 7, 15):
        one.predict(x[k:k + 1])
        for i, o in one.detection_outputs().items():
            assert util.bits_equal(o[0], a[i][k]), (k, i)


def test_spp_608_runs_and_matches_f32_cuda_core_path(workdir):
    """yolov3-spp (BASELINE configs[4] model): tensor-core bf16 result vs the engine's own f32 CUDA-core path
    (which the slim-model tests pin to the oracle); covers the 5/9/13 max-pools and the 4-way concat at 19x19."""
    import yolo2_light_b200 as yb
    secs = cfgs.yolov3_spp(608, 608)
    cfg, wts = _files(workdir, "spp_608", secs, seed=3)
    x = cfgs.synthetic_images(2, 3, 608, 608)
    a = yb.load_network(cfg, wts, batch=2)
    a.predict(x)
    b = yb.load_network(cfg, wts, batch=2)
    b.set_precision(yb.YB_PREC_FP32)
    b.predict(x)
    for i, o in a.detection_outputs().items():
        assert util.rel_l2(o, b.layer_output(i)) <= 1e-3, i


def test_pipelined_submit_collect_equals_predict(workdir):
    """yb_network_submit/collect (copies overlapped with compute, 3 batches in flight) returns exactly what the
    synchronous predict returns, batch after batch, including when slots are reused."""
    import yolo2_light_b200 as yb
    secs = cfgs.yolov3_tiny(416, 416)
    cfg, wts = _files(workdir, "tiny_416", secs)
    B = 4
    net = yb.load_network(cfg, wts, batch=B)
    batches = [cfgs.synthetic_images(B, 3, 416, 416, seed=100 + 10 * k) for k in range(7)]
    expect = []
    for x in batches:
        net.predict(x)
        expect.append({i: o.copy() for i, o in net.detection_outputs().items()})
    pinned = [yb.PinnedBuffer(B * 3 * 416 * 416) for _ in range(3)]
    inflight, got = [], []
    for k, x in enumerate(batches):
        if len(inflight) == 3:
            got.append({i: o.copy() for i, o in net.collect(inflight.pop(0)).items()})
        pinned[k % 3].array[:] = x.ravel()
        inflight.append(net.submit(pinned[k % 3].array))
    while inflight:
        got.append({i: o.copy() for i, o in net.collect(inflight.pop(0)).items()})
    assert len(got) == len(expect)
    for g, e in zip(got, expect):
        for i in e:
            assert util.bits_equal(g[i], e[i])
    with pytest.raises(yb.YbError):
        net.collect(0)   # nothing in flight
