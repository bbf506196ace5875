"""GPU (B200): the steps in front of the hot path -- letterbox_image (src/image.c:812-831) and the dynamic
per-image input quantiser (src/blas.c:108-168 via :279) on the device, with PER-IMAGE (s_in, zp_in) inside layer 0 so that
network_predict on float images works at any batch size."""
import numpy as np
import pytest

from oracle import yq_oracle as O
from yolo_quantization_b200 import darknet, synth

pytestmark = pytest.mark.gpu
YOLO_ATOL = 1e-6


@pytest.mark.parametrize("c,ih,iw,h,w,b", [(3, 375, 500, 416, 416, 2), (3, 500, 375, 416, 416, 1), (3, 100, 100, 96, 96, 3), (1, 7, 300, 64, 128, 2),
                                           (3, 64, 64, 64, 64, 1), (3, 5, 9, 32, 32, 1)])
def test_letterbox_vs_oracle(built, c, ih, iw, h, w, b):
    """device letterbox == the oracle's plain-C restatement (itself bit-equal to the compiled reference: test_letterbox_live)"""
    im = np.random.default_rng(ih * 7 + iw).random((b, c, ih, iw), dtype=np.float32)
    got = darknet.letterbox_image(im, h, w)
    for k in range(b):
        assert np.array_equal(got[k], O.letterbox(im[k], h, w))


def _net96(tmp_path, batch, act="relu6"):
    layers = synth.yolov3_tiny_quant(act)
    cfg, wts = str(tmp_path / "t.cfg"), str(tmp_path / "t.weights")
    synth.write_cfg(cfg, layers, batch=batch, width=96, height=96)
    info = synth.write_weights(wts, layers, width=96, height=96, seed=6)
    return cfg, wts, info


def _oracle_heads(info, x_f32):
    u8, s, zp = O.quantize_input(x_f32)
    ref = O.forward_network(info, u8, input_quant=(s, zp))
    return [r["f32"] for r, sl in zip(ref, info) if sl.kind == "yolo"], (s, zp)


@pytest.mark.parametrize("act", ["relu6", "leaky"])
def test_network_predict_f32_per_image_quantisation(built, tmp_path, act):
    """a batch whose images quantise differently (different maxima, one with negative values => zp_in != 0, two equal): each image's
    heads == the oracle's batch-1 walk with that image's own (s_in, zp_in); production plan and debug plan."""
    cfg, wts, info = _net96(tmp_path, 5, act)
    rng = np.random.default_rng(3)
    x = rng.random((5, 3, 96, 96), dtype=np.float32)
    x[1] *= 0.5                         # another scale
    x[2] = x[2] * 1.2 - 0.3             # negative values: zp_in != 0, padding value != 0
    x[3] = x[0]                         # shares image 0's tables
    x[4] *= 0.77
    net = darknet.load_network(cfg, wts, batch=5)
    for debug in (False, True):
        net.set_debug(debug)
        heads = net.split_heads(net.predict_f32(x))
        quants = set()
        for b in range(5):
            want, q = _oracle_heads(info, x[b])
            quants.add(q)
            for h, r in zip(heads, want):
                assert np.allclose(h[b], r, atol=YOLO_ATOL, rtol=0), f"image {b} (debug={debug})"
        assert len(quants) == 4
    # a uniform batch afterwards takes the planned tensor-core layer 0 again
    xu = np.repeat(x[:1], 5, axis=0)
    heads = net.split_heads(net.predict_f32(xu))
    want, _ = _oracle_heads(info, x[0])
    for h, r in zip(heads, want):
        assert np.allclose(h[4], r, atol=YOLO_ATOL, rtol=0)
    net.free()


def test_network_predict_image_letterbox_path(built, tmp_path):
    """test_detector's input path (examples/detector.c:903-904, :915-922): images of another size -> device letterbox ->
    device quantiser -> forward == oracle letterbox -> oracle quantiser -> oracle walk, per image."""
    cfg, wts, info = _net96(tmp_path, 2)
    im = np.random.default_rng(8).random((2, 3, 120, 160), dtype=np.float32)
    im[1] *= 0.6
    net = darknet.load_network(cfg, wts, batch=2)
    heads = net.split_heads(net.predict_image(im))
    for b in range(2):
        want, _ = _oracle_heads(info, O.letterbox(im[b], 96, 96))
        for h, r in zip(heads, want):
            assert np.allclose(h[b], r, atol=YOLO_ATOL, rtol=0), f"image {b}"
    net.free()


def test_full_yolov3_predict_f32_per_image_quantisation_replans_layer1(built, tmp_path):
    """the full yolov3 (96x96): its production plan hands layer 1 -- a narrow 3x3 stride-2 layer on the resident-bank kernel in patch mode -- a
    flat, halo-padded layer-0 tensor; a batch whose images quantise differently runs layer 0 on the generic per-image flavour, which writes a
    plain tensor: the first such forward re-plans once (layer 1 back on its plain-input flavour) and every image's heads equal the oracle's
    batch-1 walk with that image's own (s_in, zp_in); a uniform u8 batch afterwards still equals the oracle."""
    layers = synth.yolov3_quant()
    cfg, wts = str(tmp_path / "v3.cfg"), str(tmp_path / "v3.weights")
    synth.write_cfg(cfg, layers, batch=3, width=96, height=96)
    info = synth.write_weights(wts, layers, width=96, height=96, seed=5, identity_bn=False)
    net = darknet.load_network(cfg, wts, batch=3)
    assert net.layer_info(1).kernel == 1 and net.layers()[0].type == 0
    rng = np.random.default_rng(8)
    x = rng.random((3, 3, 96, 96), dtype=np.float32)
    x[1] *= 0.6
    x[2] = x[2] * 1.1 - 0.2
    heads = net.split_heads(net.predict_f32(x))
    for b in range(3):
        want, _ = _oracle_heads(info, x[b])
        for h, r in zip(heads, want):
            assert np.allclose(h[b], r, atol=YOLO_ATOL, rtol=0), f"image {b}"
    imgs = np.stack([synth.synthetic_image(s, 3, 96, 96) for s in (1, 2, 3)])
    heads = net.split_heads(net.predict_u8(imgs))
    ref = [r["f32"] for r, sl in zip(O.forward_network(info, imgs[2]), info) if sl.kind == "yolo"]
    for h, r in zip(heads, ref):
        assert np.allclose(h[2], r, atol=YOLO_ATOL, rtol=0)
    net.free()
