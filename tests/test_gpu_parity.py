"""GPU (B200): the CUDA path, called through the C ABI, against the oracle restatement on the same seeded
inputs, against the committed reference-generated goldens, and -- at BASELINE's full sizes -- through
size-independent properties.  Bit-exact for every integer/byte tensor; yolo floats within 1e-6 abs
(double exp on both sides)."""
import glob
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

from oracle import yq_oracle as O
from yolo_quantization_b200 import darknet, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
YOLO_ATOL = 1e-6


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_params(rng, n, K, zp_in, s_in=0.02, s_out=0.05):
    """random but valid per-channel parameters (0 < M < 1), prepared like blas.c:282-334 via the oracle"""
    w = rng.integers(0, 256, size=(n, K), dtype=np.uint8)
    zp_w = rng.integers(0, 256, size=n, dtype=np.uint8)
    s_w = (rng.random(n).astype(np.float32) * 0.01 + 0.001).astype(np.float32)
    bias = (rng.standard_normal(n) * 0.5).astype(np.float32)
    return w, zp_w, s_w, bias


CONV_CASES = [
    # c, h, w, n, k, stride, act, zp_in, zp_out, batch
    (3, 20, 20, 16, 3, 1, "relu6", 0, 0, 2),       # layer-0 shape class (c=3 -> channel stride 4)
    (16, 16, 18, 32, 3, 1, "relu6", 0, 0, 2),
    (32, 13, 13, 64, 3, 1, "leaky", 40, 40, 3),    # zp_in != 0 padding
    (64, 9, 9, 128, 3, 2, "leaky", 17, 33, 2),     # stride 2
    (128, 7, 5, 256, 3, 1, "relu", 5, 9, 1),
    (256, 13, 13, 30, 1, 1, "linear", 0, 128, 2),  # head: n=30 -> stride 32, quant_stop float output
    (1024, 5, 5, 256, 1, 1, "relu6", 0, 0, 1),
    (384, 6, 6, 80, 3, 1, "relu6", 3, 0, 1),       # K = 3456, n not a multiple of 64
    (5, 11, 7, 19, 3, 1, "linear", 200, 77, 2),    # odd everything (channel stride 16 with 11 pad lanes)
    (512, 4, 4, 48, 3, 1, "relu6", 0, 0, 1),       # K = 4608 (largest K of the net)
    (8, 6, 6, 16, 5, 1, "leaky", 9, 20, 1),        # 5x5 kernel
    (16, 1, 1, 16, 3, 1, "relu6", 7, 0, 4),        # 1x1 image: every tap but the centre is padding
]


def run_conv_case(case, kernel=-1, saturate=0, s_out=0.05):
    c, h, w, n, k, stride, act, zp_in, zp_out, batch = case
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()))
    K = c * k * k
    wq, zp_w, s_w, bias = make_params(rng, n, K, zp_in)
    spec = synth.LayerSpec("conv", n, k, stride, 1, 0, act)
    sl = synth.SynthLayer("conv", c, h, w, n, 0, 0, spec, s_out=s_out, biases=bias, s_w=s_w, zp_w=zp_w,
                          w_u8=wq.reshape(n, c, k, k))
    p = O.prepare_conv(sl, 0.02, zp_in)
    x = rng.integers(0, 256, size=(batch, c, h, w), dtype=np.uint8)
    qs = 1 if act == "linear" else 0
    layer = darknet.ConvolutionalLayerQuant(h, w, c, n, k, stride, k // 2, synth.ACT_CODES[act], wq, zp_w,
                                            p["biases_int32"], p["M_value"], p["M0_right_shift_value"], zp_in, zp_out,
                                            s_out, quant_stop_flag=qs, saturate=saturate, kernel=kernel)
    got = layer.forward(x)
    layer.free()
    for b in range(batch):
        acc = O.conv_acc(x[b], wq.reshape(n, c, k, k), zp_w, stride, k // 2, zp_in)
        assert np.array_equal(got["acc"][b], acc), f"int32 accumulator mismatch, image {b}"
        u8 = O.requant(acc, p["biases_int32"], p["M_value"], p["M0_right_shift_value"], synth.ACT_CODES[act], zp_out)
        if saturate:
            continue
        assert np.array_equal(got["u8"][b], u8), f"uint8 mismatch, image {b}"
        if qs:
            assert np.array_equal(got["f32"][b], O.dequant(u8, zp_out, s_out))
    return got


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "c%d_%dx%d_n%d_k%d_s%d_%s_zi%d" % c[:8])
def test_conv_simt_vs_oracle(built, case):
    run_conv_case(case, kernel=0)


TC_CASES = [
    # c, h, w, n, k, stride, act, zp_in, zp_out, batch      (tcgen05 flavour: stride 1, k in {1,3}, c % 64 == 0)
    (64, 52, 52, 128, 3, 1, "relu6", 0, 0, 2),     # layer 6 shape: KC = 64 (SWIZZLE_64B), BN = 128
    (128, 26, 26, 256, 3, 1, "relu6", 0, 0, 3),    # layer 8: KC = 128, two N tiles
    (256, 13, 13, 512, 3, 1, "relu6", 0, 0, 9),    # layer 10/14: 13-wide rows, 9 images stacked per tile
    (512, 13, 13, 64, 3, 1, "relu6", 0, 0, 2),     # K = 4608 (layer 12's K), BN = 64
    (1024, 13, 13, 256, 1, 1, "relu6", 0, 0, 2),   # layer 13: 1x1, K = 1024
    (512, 13, 13, 30, 1, 1, "linear", 0, 128, 2),  # layer 15 head: n = 30 -> BN 32, float side output
    (256, 26, 26, 30, 1, 1, "linear", 0, 128, 1),  # layer 22 head
    (384, 26, 26, 256, 3, 1, "relu6", 0, 0, 1),    # layer 21: c = 384 = 3 chunks of 128
    (256, 13, 13, 128, 1, 1, "relu6", 0, 0, 10),   # layer 18, batch not a multiple of the 9-image tile
    (128, 9, 7, 80, 3, 1, "leaky", 40, 33, 3),     # zp_in != 0: border correction; n = 80 -> partial N tile
    (64, 5, 130, 48, 3, 1, "leaky", 17, 5, 1),     # wide rows (TW = 128 + remainder), n = 48
    (64, 1, 1, 32, 3, 1, "relu", 9, 3, 5),         # 1x1 image: 8 of 9 taps are padding
    (192, 6, 6, 17, 1, 1, "linear", 0, 7, 2),      # c = 192 -> KC = 64, n = 17
]


@pytest.mark.parametrize("case", TC_CASES, ids=lambda c: "c%d_%dx%d_n%d_k%d_s%d_%s_zi%d" % c[:8])
def test_conv_tcgen05_vs_oracle(built, case):
    run_conv_case(case, kernel=1)


TC_SMALL_CASES = [
    # small-c tcgen05 flavour (threads build the im2col rows): c <= 32, 3x3, n in {16, 32, 64}
    (3, 40, 40, 16, 3, 1, "relu6", 0, 0, 2),       # layer 0 shape class
    (3, 33, 17, 16, 3, 1, "leaky", 77, 40, 3),     # zp_in != 0 written straight into the padded taps
    (16, 24, 24, 32, 3, 1, "relu6", 0, 0, 2),      # layer 2
    (32, 20, 20, 64, 3, 1, "relu6", 0, 0, 2),      # layer 4
    (32, 13, 11, 64, 3, 2, "leaky", 40, 40, 3),    # stride 2
    (16, 7, 9, 16, 3, 1, "linear", 5, 100, 1),     # float side output (quant_stop)
    (12, 9, 9, 20, 3, 1, "relu", 9, 3, 2),         # c = 12 -> stride 16 (pad lanes), n = 20 -> stride 32
    (3, 416, 416, 16, 3, 1, "relu6", 0, 0, 1),     # the real layer 0
    (4, 1, 1, 64, 3, 1, "relu6", 200, 0, 130),     # 1x1 images, more than one tile of them
]


@pytest.mark.parametrize("case", TC_SMALL_CASES, ids=lambda c: "c%d_%dx%d_n%d_k%d_s%d_%s_zi%d" % c[:8])
def test_conv_tcgen05_small_c_vs_oracle(built, case):
    run_conv_case(case, kernel=1)


@pytest.mark.parametrize("case", [
    (3, 40, 40, 16, 3, 1, "relu6", 0, 0, 2),
    (3, 33, 17, 16, 3, 1, "leaky", 77, 40, 3),     # odd sizes: the last pooled row/column sees out-of-image taps (count as 0)
    (16, 24, 24, 32, 3, 1, "relu6", 0, 0, 2),
    (32, 21, 19, 64, 3, 1, "relu6", 0, 0, 2),
    (16, 9, 9, 20, 3, 1, "relu", 9, 3, 2),
    (3, 70, 52, 16, 3, 1, "relu6", 0, 7, 2),       # several 16x8 tiles, partial tiles on both borders, zp_out != 0
    (16, 37, 50, 32, 3, 1, "relu6", 11, 0, 1),
], ids=lambda c: "c%d_%dx%d_n%d" % c[:4])
@pytest.mark.parametrize("s_out", [0.05, 6.0], ids=["wrapping", "in_range"])
def test_conv_fused_maxpool_vs_oracle(built, case, s_out):
    """conv + maxpool(2,2) in one launch == oracle conv followed by oracle maxpool (maxpool_layer.c:109-153).
    s_out = 0.05 drives most bytes through the uint8 wrap (the pool-first kernel's per-pixel FP64 path);
    s_out = 6.0 keeps them in range (its max-then-requantize path)."""
    c, h, w, n, k, stride, act, zp_in, zp_out, batch = case
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()) + 1)
    wq, zp_w, s_w, bias = make_params(rng, n, c * k * k, zp_in)
    spec = synth.LayerSpec("conv", n, k, stride, 1, 0, act)
    sl = synth.SynthLayer("conv", c, h, w, n, 0, 0, spec, s_out=s_out, biases=bias, s_w=s_w, zp_w=zp_w, w_u8=wq.reshape(n, c, k, k))
    p = O.prepare_conv(sl, 0.02, zp_in)
    x = rng.integers(0, 256, size=(batch, c, h, w), dtype=np.uint8)
    layer = darknet.ConvolutionalLayerQuant(h, w, c, n, k, stride, 1, synth.ACT_CODES[act], wq, zp_w, p["biases_int32"], p["M_value"],
                                            p["M0_right_shift_value"], zp_in, zp_out, s_out, kernel=1)
    assert layer.can_fuse_maxpool
    for want_conv in (True, False):
        got = layer.forward_pooled(x, want_conv=want_conv)
        for b in range(batch):
            acc = O.conv_acc(x[b], wq.reshape(n, c, k, k), zp_w, stride, 1, zp_in)
            u8 = O.requant(acc, p["biases_int32"], p["M_value"], p["M0_right_shift_value"], synth.ACT_CODES[act], zp_out)
            assert np.array_equal(got["pool"][b], O.maxpool(u8, 2, 2)), f"pooled mismatch image {b}"
            if want_conv:
                assert np.array_equal(got["u8"][b], u8)
    layer.free()


# every distinct convolution shape of the full yolov3 (BASELINE configs[4]: 75 conv layers, leaky, stride-2 down-samplers,
# linear 1x1 heads with 255 filters): (c, n, size, stride, activation).  The reference pins each conv layer on its own
# (SURVEY Appendix F); the quantized shortcut between them does not exist in the reference.
YOLOV3_SHAPES = [
    (3, 32, 3, 1, "leaky"), (32, 64, 3, 2, "leaky"), (64, 32, 1, 1, "leaky"), (32, 64, 3, 1, "leaky"),
    (64, 128, 3, 2, "leaky"), (128, 64, 1, 1, "leaky"), (64, 128, 3, 1, "leaky"),
    (128, 256, 3, 2, "leaky"), (256, 128, 1, 1, "leaky"), (128, 256, 3, 1, "leaky"),
    (256, 512, 3, 2, "leaky"), (512, 256, 1, 1, "leaky"), (256, 512, 3, 1, "leaky"),
    (512, 1024, 3, 2, "leaky"), (1024, 512, 1, 1, "leaky"), (512, 1024, 3, 1, "leaky"),
    (1024, 255, 1, 1, "linear"), (768, 256, 1, 1, "leaky"), (512, 255, 1, 1, "linear"),
    (384, 128, 1, 1, "leaky"), (256, 255, 1, 1, "linear"),
]


@pytest.mark.parametrize("shape", YOLOV3_SHAPES, ids=lambda s: "c%d_n%d_k%d_s%d_%s" % s)
def test_full_yolov3_conv_shapes_vs_oracle(built, shape):
    """BASELINE configs[4]: each conv shape of the full yolov3 through the default flavour selection (flat / flat2 for the
    stride-1 layers with c % 64 == 0, SIMT for the stride-2 down-samplers, small-c for c <= 32) == oracle, bit for bit;
    zp_in = 40 like a leaky net (padding value != 0)."""
    c, n, k, stride, act = shape
    h = w = 12
    rng = np.random.default_rng(zlib.crc32(repr(shape).encode()) + 3)
    wq, zp_w, s_w, bias = make_params(rng, n, c * k * k, 40)
    spec = synth.LayerSpec("conv", n, k, stride, 1, 0, act)
    sl = synth.SynthLayer("conv", c, h, w, n, 0, 0, spec, s_out=0.05, biases=bias, s_w=s_w, zp_w=zp_w, w_u8=wq.reshape(n, c, k, k))
    p = O.prepare_conv(sl, 0.02, 40)
    x = rng.integers(0, 256, size=(2, c, h, w), dtype=np.uint8)
    qs = 1 if act == "linear" else 0
    layer = darknet.ConvolutionalLayerQuant(h, w, c, n, k, stride, k // 2, synth.ACT_CODES[act], wq, zp_w, p["biases_int32"], p["M_value"],
                                            p["M0_right_shift_value"], 40, 33, 0.05, quant_stop_flag=qs)
    got = layer.forward_flat(x, halo_fill=40) if layer.flat_supported else layer.forward(x)
    for b in range(2):
        acc = O.conv_acc(x[b], wq.reshape(n, c, k, k), zp_w, stride, k // 2, 40)
        assert np.array_equal(got["acc"][b], acc), f"int32 accumulator mismatch, image {b}"
        u8 = O.requant(acc, p["biases_int32"], p["M_value"], p["M0_right_shift_value"], synth.ACT_CODES[act], 33)
        assert np.array_equal(got["u8"][b], u8), f"uint8 mismatch, image {b}"
        if qs:
            assert np.array_equal(got["f32"][b], O.dequant(u8, 33, 0.05))
    layer.free()


FLAT_CASES = [
    # c, h, w, n, k, act, zp_in, zp_out, batch
    (64, 13, 13, 128, 3, "relu6", 0, 0, 3),
    (64, 52, 52, 128, 3, "leaky", 40, 40, 1),      # widest supported row (patch = 236 positions), KC = 64
    (128, 7, 5, 256, 3, "relu", 5, 9, 2),
    (128, 26, 26, 64, 3, "relu6", 11, 0, 2),
    (256, 13, 13, 30, 1, "linear", 0, 128, 2),     # head: n = 30 -> stride 32, quant_stop float output
    (1024, 5, 5, 256, 1, "relu6", 0, 0, 1),
    (384, 6, 6, 80, 3, "relu6", 3, 0, 1),          # K = 3456, n not a multiple of 64
    (512, 4, 4, 48, 3, "leaky", 200, 77, 1),       # K = 4608
    (64, 20, 33, 64, 3, "linear", 77, 20, 2),
    (192, 1, 1, 32, 3, "relu6", 7, 0, 4),          # 1x1 image: every tap but the centre is halo
    (32, 20, 33, 64, 3, "leaky", 13, 40, 2),       # c = 32: 32-byte patch rows (SWIZZLE_32B, KC = 32), the persistent two-tile form only
    (32, 208, 208, 64, 3, "leaky", 0, 3, 1),       # layer 3 of the full yolov3 at its real size (three TMA boxes per patch)
    (32, 9, 7, 128, 1, "relu6", 5, 0, 3),
]


@pytest.mark.parametrize("case", FLAT_CASES, ids=lambda c: "c%d_%dx%d_n%d_k%d_%s" % c[:6])
def test_conv_flat_flavour_vs_oracle(built, case):
    """flat-strip tcgen05 flavour (one shared-memory patch per channel chunk, row-shifted descriptors per tap): int32
    accumulators, bytes and floats equal the oracle; the output strip's halo holds exactly halo_fill afterwards."""
    c, h, w, n, k, act, zp_in, zp_out, batch = case
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()) + 9)
    wq, zp_w, s_w, bias = make_params(rng, n, c * k * k, zp_in)
    spec = synth.LayerSpec("conv", n, k, 1, 1, 0, act)
    sl = synth.SynthLayer("conv", c, h, w, n, 0, 0, spec, s_out=0.05, biases=bias, s_w=s_w, zp_w=zp_w, w_u8=wq.reshape(n, c, k, k))
    p = O.prepare_conv(sl, 0.02, zp_in)
    x = rng.integers(0, 256, size=(batch, c, h, w), dtype=np.uint8)
    qs = 1 if act == "linear" and n == 30 else 0
    layer = darknet.ConvolutionalLayerQuant(h, w, c, n, k, 1, k // 2, synth.ACT_CODES[act], wq, zp_w, p["biases_int32"], p["M_value"],
                                            p["M0_right_shift_value"], zp_in, zp_out, 0.05, quant_stop_flag=qs)
    assert layer.flat_supported
    for want_acc in (True, False):   # the side-output variant and the production variant
        got = layer.forward_flat(x, halo_fill=zp_out ^ 0x5A, want_acc=want_acc)
        assert got["halo_ok"], "halo / pad lanes of the output strip"
        for b in range(batch):
            acc = O.conv_acc(x[b], wq.reshape(n, c, k, k), zp_w, 1, k // 2, zp_in)
            if want_acc:
                assert np.array_equal(got["acc"][b], acc), f"int32 accumulator mismatch, image {b}"
            u8 = O.requant(acc, p["biases_int32"], p["M_value"], p["M0_right_shift_value"], synth.ACT_CODES[act], zp_out)
            assert np.array_equal(got["u8"][b], u8), f"uint8 mismatch, image {b}"
            if qs:
                assert np.array_equal(got["f32"][b], O.dequant(u8, zp_out, 0.05))
    layer.free()


PW_CASES = [
    # c, h, w, n, act, zp_in, zp_out, batch, yolo classes (None: a plain 1x1 layer)
    (64, 20, 33, 32, "leaky", 13, 40, 2, None),      # one chunk of 64 channels per tile (one commit per tile), four accumulators, 32-byte rows
    (128, 26, 26, 64, "leaky", 0, 3, 3, None),       # one chunk of 128, 64-byte rows
    (128, 9, 7, 128, "relu6", 5, 0, 2, None),        # one chunk, two accumulators (N = 144 columns each), channel halves per group
    (256, 13, 13, 128, "relu6", 0, 0, 5, None),      # layer 18 of yolov3-tiny: two chunks per tile
    (384, 11, 6, 128, "leaky", 77, 9, 2, None),      # three chunks (layer 99 of the full yolov3)
    (256, 6, 10, 100, "linear", 3, 128, 2, None),    # n = 100 -> stride 112: not a pointwise shape, the flat kernels keep it
    (512, 5, 5, 64, "relu", 9, 20, 1, None),         # four chunks, four accumulators
    (256, 13, 13, 30, "linear", 0, 128, 3, 5),       # head of yolov3-tiny (layer 22 class): 3 anchors x (5 + 5)
    (512, 13, 13, 30, "linear", 0, 100, 2, 5),       # layer 15: four chunks
    (256, 12, 9, 255, "linear", 0, 120, 2, 80),      # heads of the full yolov3: 255 channels, two accumulators of 256 columns, ones row = row 255
    (512, 7, 7, 255, "linear", 4, 131, 1, 80),       # 128 KB filter bank, three ring stages
    (128, 50, 41, 64, "leaky", 0, 0, 4, None),       # several tiles per CTA on a small grid is not reached here; many tiles, ragged last tile
]


@pytest.mark.parametrize("case", PW_CASES, ids=lambda c: "c%d_%dx%d_n%d_%s" % c[:5])
def test_conv_pointwise_flavour_vs_oracle(built, case, monkeypatch):
    """1x1 layers and detection heads on the streaming pointwise flavour (filter bank resident in shared memory, ones row for the
    activation sums): the production launch (no side outputs) equals the oracle byte for byte, heads within the yolo tolerance, the
    output strip's halo holds halo_fill; and equals the flat kernels' result (YQ_PW=0) bit for bit, floats included."""
    c, h, w, n, act, zp_in, zp_out, batch, classes = case
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()) + 21)
    wq, zp_w, s_w, bias = make_params(rng, n, c, zp_in)
    spec = synth.LayerSpec("conv", n, 1, 1, 1, 0, act)
    sl = synth.SynthLayer("conv", c, h, w, n, 0, 0, spec, s_out=0.05, biases=bias, s_w=s_w, zp_w=zp_w, w_u8=wq.reshape(n, c, 1, 1))
    p = O.prepare_conv(sl, 0.02, zp_in)
    x = rng.integers(0, 256, size=(batch, c, h, w), dtype=np.uint8)
    qs = 1 if classes is not None else 0
    layer = darknet.ConvolutionalLayerQuant(h, w, c, n, 1, 1, 0, synth.ACT_CODES[act], wq, zp_w, p["biases_int32"], p["M_value"],
                                            p["M0_right_shift_value"], zp_in, zp_out, 0.05, quant_stop_flag=qs)
    assert layer.flat_supported
    fill = zp_out ^ 0x5A
    got = layer.forward_flat(x, halo_fill=fill, want_acc=False, yolo_classes=classes)
    monkeypatch.setenv("YQ_PW", "0")
    old = layer.forward_flat(x, halo_fill=fill, want_acc=False, yolo_classes=classes)
    monkeypatch.delenv("YQ_PW")
    assert got["halo_ok"], "halo / pad lanes of the output strip"
    assert np.array_equal(got["u8"], old["u8"])
    for b in range(batch):
        acc = O.conv_acc(x[b], wq.reshape(n, c, 1, 1), zp_w, 1, 0, zp_in)
        u8 = O.requant(acc, p["biases_int32"], p["M_value"], p["M0_right_shift_value"], synth.ACT_CODES[act], zp_out)
        assert np.array_equal(got["u8"][b], u8), f"uint8 mismatch, image {b}"
        if classes is not None:
            ref = O.yolo(O.dequant(u8, zp_out, 0.05), n // (classes + 5), classes)
            assert np.allclose(got["yolo"][b], ref, atol=YOLO_ATOL, rtol=0), f"yolo mismatch, image {b}"
            assert np.array_equal(got["yolo"][b], old["yolo"][b])
    layer.free()


@pytest.mark.parametrize("case", [(256, 13, 13, 128, "relu6", 0, 0, 5), (128, 9, 7, 64, "leaky", 13, 40, 2), (64, 6, 11, 32, "leaky", 0, 3, 3),
                                  (512, 5, 5, 100, "linear", 7, 128, 1)],
                         ids=lambda c: "c%d_%dx%d_n%d_%s" % c[:5])
def test_conv_pointwise_fused_upsample_vs_oracle(built, case):
    """1x1 convolution + the stride-2 upsample behind it in one launch (layer 18 -> upsample 19 of yolov3-tiny): every pixel of the
    oracle's convolution output four times (blas.c:334-351), the upsampled strip's halo untouched, pad lanes zero"""
    c, h, w, n, act, zp_in, zp_out, batch = case
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()) + 77)
    wq, zp_w, s_w, bias = make_params(rng, n, c, zp_in)
    spec = synth.LayerSpec("conv", n, 1, 1, 1, 0, act)
    sl = synth.SynthLayer("conv", c, h, w, n, 0, 0, spec, s_out=0.05, biases=bias, s_w=s_w, zp_w=zp_w, w_u8=wq.reshape(n, c, 1, 1))
    p = O.prepare_conv(sl, 0.02, zp_in)
    x = rng.integers(0, 256, size=(batch, c, h, w), dtype=np.uint8)
    layer = darknet.ConvolutionalLayerQuant(h, w, c, n, 1, 1, 0, synth.ACT_CODES[act], wq, zp_w, p["biases_int32"], p["M_value"],
                                            p["M0_right_shift_value"], zp_in, zp_out, 0.05)
    if n == 100:
        assert not layer.flat_up2_supported     # n = 100 -> stride 112: not a pointwise shape
        layer.free()
        return
    assert layer.flat_up2_supported
    got = layer.forward_flat_up2(x)
    assert got["halo_ok"], "halo / pad lanes of the upsampled strip"
    for b in range(batch):
        acc = O.conv_acc(x[b], wq.reshape(n, c, 1, 1), zp_w, 1, 0, zp_in)
        u8 = O.requant(acc, p["biases_int32"], p["M_value"], p["M0_right_shift_value"], synth.ACT_CODES[act], zp_out)
        assert np.array_equal(got["u8"][b], O.upsample(u8, 2)), f"image {b}"
    layer.free()


@pytest.mark.parametrize("case", [(384, 26, 26, 256, 128, "relu6", 0, 0, 3, 3), (384, 9, 12, 256, 256, "leaky", 11, 7, 2, 3), (512, 13, 13, 512, 128, "relu6", 5, 0, 2, 3),
                                  (320, 7, 7, 128, 64, "leaky", 0, 9, 1, 3), (384, 11, 6, 128, 128, "leaky", 77, 9, 2, 1), (192, 20, 9, 64, 64, "leaky", 0, 3, 3, 1)],
                         ids=lambda c: "c%d_%dx%d_n%d_first%d" % c[:5])
def test_conv_reads_two_tensors_as_their_concatenation(built, case):
    """the convolution behind a route that is never materialised (layer 21 of yolov3-tiny reads [upsampled layer 18 | layer 8]; layer 99
    of the full yolov3, a 1x1 layer on the pointwise flavour, reads [upsampled layer 96 | layer 36]): two flat tensors in, bytes equal to
    the oracle's convolution over their concatenation and to the one-tensor launch"""
    c, h, w, n, c_first, act, zp_in, zp_out, batch, k = case
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()) + 5)
    wq, zp_w, s_w, bias = make_params(rng, n, c * k * k, zp_in)
    spec = synth.LayerSpec("conv", n, k, 1, 1, 0, act)
    sl = synth.SynthLayer("conv", c, h, w, n, 0, 0, spec, s_out=0.05, biases=bias, s_w=s_w, zp_w=zp_w, w_u8=wq.reshape(n, c, k, k))
    p = O.prepare_conv(sl, 0.02, zp_in)
    x = rng.integers(0, 256, size=(batch, c, h, w), dtype=np.uint8)
    layer = darknet.ConvolutionalLayerQuant(h, w, c, n, k, 1, k // 2, synth.ACT_CODES[act], wq, zp_w, p["biases_int32"], p["M_value"],
                                            p["M0_right_shift_value"], zp_in, zp_out, 0.05)
    assert layer.flat_cat_supported(c_first) and not layer.flat_cat_supported(c_first + 16)
    got = layer.forward_flat_cat(x, c_first, halo_fill=zp_out)
    one = layer.forward_flat(x, halo_fill=zp_out, want_acc=False)
    assert np.array_equal(got["u8"], one["u8"])
    for b in range(batch):
        acc = O.conv_acc(x[b], wq.reshape(n, c, k, k), zp_w, 1, k // 2, zp_in)
        u8 = O.requant(acc, p["biases_int32"], p["M_value"], p["M0_right_shift_value"], synth.ACT_CODES[act], zp_out)
        assert np.array_equal(got["u8"][b], u8), f"uint8 mismatch, image {b}"
    layer.free()


ROWS_CASES = [
    # c, h, w, n, zp_in, zp_out, batch
    (3, 48, 64, 16, 0, 0, 2),
    (3, 34, 70, 16, 9, 5, 1),        # partial tiles in x and y, w % 4 != 0
    (3, 16, 32, 32, 0, 0, 1),
    (16, 32, 48, 32, 0, 0, 2),
    (16, 18, 22, 32, 40, 0, 1),      # partial tiles, halo = zp_in = 40
    (16, 16, 16, 16, 0, 3, 1),
    (16, 20, 36, 64, 0, 0, 1),
    (32, 16, 16, 64, 0, 0, 2),
    (32, 26, 40, 64, 17, 0, 1),
    (32, 14, 18, 32, 0, 0, 1),
    (20, 12, 20, 32, 0, 0, 1),       # c = 20 -> channel stride 32 with 12 pad lanes
    (64, 20, 36, 128, 0, 0, 1),      # c = 64 (layer 6 class): four 16-channel blocks, n = 128 as two launches of 64 channels
    (64, 16, 16, 64, 7, 0, 2),
    (64, 52, 52, 128, 0, 0, 1),      # layer 6 itself
]


@pytest.mark.parametrize("case", ROWS_CASES, ids=lambda c: "c%d_%dx%d_n%d" % c[:4])
@pytest.mark.parametrize("s_out", [0.05, 6.0], ids=["wrapping", "in_range"])
@pytest.mark.parametrize("out_pad", [0, 1])
@pytest.mark.parametrize("variant", [2, 1], ids=["two_signed_blocks", "ones_rows"])
def test_conv_rows_flavour_vs_oracle(built, case, s_out, out_pad, variant, monkeypatch):
    """halo-input conv + RELU6 + maxpool(2,2) (no im2col; Toeplitz / even-odd MMA groups) == oracle conv + oracle maxpool;
    with out_pad = 1 the halo of the output tensor must stay untouched.  Both weight forms: two signed blocks
    h + l = w - zp_w (every difference <= 254) and, when some w - zp_w = 255, all-ones rows + zero-point correction."""
    c, h, w, n, zp_in, zp_out, batch = case
    k = 3
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()) + 5)
    wq, zp_w, s_w, bias = make_params(rng, n, c * k * k, zp_in)
    if variant == 2:
        monkeypatch.setenv("YQ_ROWS_TWO", "1")          # (the default takes this form only where it pays: c <= 16, n <= 32)
        zp_w = np.maximum(zp_w, 1).astype(zp_w.dtype)
        zp_w[n // 2] = 255                              # w - zp_w down to -255
        wq[n // 2, 0] = 0
    else:
        zp_w[n - 1] = 0
        wq[n - 1, -1] = 255                             # the one difference the signed blocks cannot hold
    spec = synth.LayerSpec("conv", n, k, 1, 1, 0, "relu6")
    sl = synth.SynthLayer("conv", c, h, w, n, 0, 0, spec, s_out=s_out, biases=bias, s_w=s_w, zp_w=zp_w, w_u8=wq.reshape(n, c, k, k))
    p = O.prepare_conv(sl, 0.02, zp_in)
    x = rng.integers(0, 256, size=(batch, c, h, w), dtype=np.uint8)
    layer = darknet.ConvolutionalLayerQuant(h, w, c, n, k, 1, 1, synth.ACT_CODES["relu6"], wq, zp_w, p["biases_int32"], p["M_value"],
                                            p["M0_right_shift_value"], zp_in, zp_out, s_out)
    assert layer.rows_supported and layer.rows_variant == (variant if c != 64 else 1)    # (c = 64 has no room for the second block)
    got = layer.forward_rows_pooled(x, out_pad=out_pad)
    for b in range(batch):
        acc = O.conv_acc(x[b], wq.reshape(n, c, k, k), zp_w, 1, 1, zp_in)
        u8 = O.requant(acc, p["biases_int32"], p["M_value"], p["M0_right_shift_value"], synth.ACT_CODES["relu6"], zp_out)
        assert np.array_equal(got[b], O.maxpool(u8, 2, 2)), f"pooled mismatch image {b}"
    if out_pad:
        raw, pad, pitch, rows = layer.last_rows_raw
        cs = darknet.channel_stride(n)
        t = raw[:batch * rows * pitch * cs].reshape(batch, rows, pitch, cs).copy()
        t[:, pad:pad + h // 2, pad:pad + w // 2, :] = 0xEE
        assert (t == 0xEE).all(), "the rows kernel wrote outside the pooled interior"
    layer.free()


ROWS_NCHW_CASES = [
    # h, w, n, batch   (c = 3, zp_in = 0, w % 16 == 0, w >= 64)
    (48, 64, 16, 2),
    (34, 80, 16, 3),       # partial tiles in x and y: the TMA box hangs over the right and bottom edges (zero fill)
    (16, 64, 32, 1),
    (50, 64, 16, 2),
    (64, 416, 16, 1),      # the real row width: 6.5 tiles of 64 pixels (dense layer-0 flavour), 13 of 32 (rows flavour)
    (22, 208, 16, 2),
    (130, 96, 16, 3),      # more tiles than one wave of a small grid: both accumulators and every staging slot get reused
]


@pytest.mark.parametrize("case", ROWS_NCHW_CASES, ids=lambda c: "%dx%d_n%d_b%d" % c)
@pytest.mark.parametrize("variant", [2, 1, 3], ids=["two_signed_blocks", "ones_rows", "rows_kernel_two_blocks"])
@pytest.mark.parametrize("s_out", [0.05, 6.0], ids=["wrapping", "in_range"])
def test_conv_rows_flavour_reads_nchw_planes(built, case, variant, s_out, monkeypatch):
    """layer-0 class: the kernel fetches the [b,3,h,w] planes itself (TMA, zero fill outside the image) and rearranges
    them on chip == oracle conv + maxpool; same result as through the padded NHWC4 copy.  n = 16 with two signed weight blocks
    runs the dense layer-0 flavour (yq_conv_tc_l0.cu: 8 pixels per MMA row), everything else -- and variant 3, which switches
    the dense flavour off -- the rows flavour's planar form."""
    h, w, n, batch = case
    if variant == 3:
        monkeypatch.setenv("YQ_NO_L0", "1")
        variant = 2
    c, k, zp_in, zp_out = 3, 3, 0, 0
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()) + 11)
    wq, zp_w, s_w, bias = make_params(rng, n, c * k * k, zp_in)
    if variant == 2:
        monkeypatch.setenv("YQ_ROWS_TWO", "1")
        zp_w = np.maximum(zp_w, 1).astype(zp_w.dtype)
    else:
        zp_w[n - 1] = 0
        wq[n - 1, -1] = 255
    spec = synth.LayerSpec("conv", n, k, 1, 1, 0, "relu6")
    sl = synth.SynthLayer("conv", c, h, w, n, 0, 0, spec, s_out=s_out, biases=bias, s_w=s_w, zp_w=zp_w, w_u8=wq.reshape(n, c, k, k))
    p = O.prepare_conv(sl, 0.02, zp_in)
    x = rng.integers(0, 256, size=(batch, c, h, w), dtype=np.uint8)
    zp_out = 3 if s_out > 1 else 0
    layer = darknet.ConvolutionalLayerQuant(h, w, c, n, k, 1, 1, synth.ACT_CODES["relu6"], wq, zp_w, p["biases_int32"], p["M_value"],
                                            p["M0_right_shift_value"], zp_in, zp_out, s_out)
    assert layer.rows_nchw_supported and layer.rows_variant == variant
    got = layer.forward_rows_pooled(x, out_pad=1, nchw=True)
    for b in range(batch):
        acc = O.conv_acc(x[b], wq.reshape(n, c, k, k), zp_w, 1, 1, zp_in)
        u8 = O.requant(acc, p["biases_int32"], p["M_value"], p["M0_right_shift_value"], synth.ACT_CODES["relu6"], zp_out)
        assert np.array_equal(got[b], O.maxpool(u8, 2, 2)), f"pooled mismatch image {b}"
    assert np.array_equal(got, layer.forward_rows_pooled(x, out_pad=1, nchw=False))
    layer.free()


def test_conv_fused_maxpool_large_accumulators(built):
    """accumulators beyond 2^22 with in-range bytes: the integer-form requantize is no longer guaranteed to equal
    the reference's double multiply, so the pool-first kernel must take its FP64 path and still match bit for bit."""
    c, h, w, n, k = 32, 20, 24, 64, 3
    rng = np.random.default_rng(77)
    wq = rng.integers(180, 256, size=(n, c * k * k), dtype=np.uint8)
    zp_w = rng.integers(0, 20, size=n, dtype=np.uint8)
    s_w = (rng.random(n).astype(np.float32) * 0.01 + 0.001).astype(np.float32)
    bias = (rng.standard_normal(n) * 0.5).astype(np.float32)
    s_out = 25.0
    spec = synth.LayerSpec("conv", n, k, 1, 1, 0, "relu6")
    sl = synth.SynthLayer("conv", c, h, w, n, 0, 0, spec, s_out=s_out, biases=bias, s_w=s_w, zp_w=zp_w, w_u8=wq.reshape(n, c, k, k))
    p = O.prepare_conv(sl, 0.02, 0)
    x = rng.integers(150, 256, size=(2, c, h, w), dtype=np.uint8)
    layer = darknet.ConvolutionalLayerQuant(h, w, c, n, k, 1, 1, synth.ACT_CODES["relu6"], wq, zp_w, p["biases_int32"], p["M_value"],
                                            p["M0_right_shift_value"], 0, 0, s_out, kernel=1)
    got = layer.forward_pooled(x, want_conv=False)
    for b in range(2):
        acc = O.conv_acc(x[b], wq.reshape(n, c, k, k), zp_w, 1, 1, 0)
        assert acc.max() > (1 << 22)
        u8 = O.requant(acc, p["biases_int32"], p["M_value"], p["M0_right_shift_value"], synth.ACT_CODES["relu6"], 0)
        assert np.array_equal(got["pool"][b], O.maxpool(u8, 2, 2))
    layer.free()


def test_conv_wrap_semantics(built):
    """tiny s_out forces q + zp_out far outside [0,255]: the store must WRAP like the reference (A.3)."""
    case = (16, 8, 8, 32, 3, 1, "linear", 0, 100, 1)
    got = run_conv_case(case, kernel=0, s_out=0.0005)
    assert len(np.unique(got["u8"])) > 200          # wrapped values cover the byte range


def test_conv_saturate_switch(built):
    case = (16, 8, 8, 32, 3, 1, "linear", 0, 100, 1)
    got = run_conv_case(case, kernel=0, saturate=1, s_out=0.0005)
    frac_sat = np.isin(got["u8"], (0, 255)).mean()
    assert frac_sat > 0.9


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "layer_*.npz"))), ids=os.path.basename)
def test_conv_vs_reference_goldens(built, path):
    """single-conv fixtures produced by the compiled reference (per-layer oracle trick)"""
    g = np.load(path)
    c, h, w, n, k, stride, pad, act, bn, qs, zp_in, zp_out = (int(v) for v in g["geom"])
    layer = darknet.ConvolutionalLayerQuant(h, w, c, n, k, stride, pad, act, g["w_u8"], g["zp_w"], g["biases_int32"],
                                            g["M_value"], g["M0_right_shift_value"], zp_in, zp_out, float(g["scales"][1]),
                                            quant_stop_flag=qs)
    got = layer.forward(g["x"][None])
    assert np.array_equal(got["acc"][0], g["ref_int32"])
    assert np.array_equal(got["u8"][0], g["ref_uint8"])
    if qs:
        assert np.array_equal(got["f32"][0], g["ref_f32"].reshape(got["f32"][0].shape))


@pytest.mark.parametrize("c,h,w,size,stride", [(16, 16, 16, 2, 2), (512, 13, 13, 2, 1), (3, 9, 11, 2, 2), (32, 7, 9, 3, 2),
                                               (20, 6, 6, 2, 2)])
def test_maxpool(built, c, h, w, size, stride):
    x = np.random.default_rng(c + h).integers(0, 256, size=(2, c, h, w), dtype=np.uint8)
    got = darknet.forward_maxpool_layer_quant_gpu(x, size, stride)
    for b in range(2):
        assert np.array_equal(got[b], O.maxpool(x[b], size, stride))


@pytest.mark.parametrize("c,h,w,stride", [(128, 13, 13, 2), (3, 5, 4, 2), (16, 3, 3, 3)])
def test_upsample(built, c, h, w, stride):
    x = np.random.default_rng(c).integers(0, 256, size=(2, c, h, w), dtype=np.uint8)
    got = darknet.forward_upsample_layer_quant_gpu(x, stride)
    for b in range(2):
        assert np.array_equal(got[b], O.upsample(x[b], stride))


@pytest.mark.parametrize("cs", [(128, 256), (16,), (30, 16), (3, 5, 8)])
def test_route(built, cs):
    rng = np.random.default_rng(sum(cs))
    xs = [rng.integers(0, 256, size=(2, c, 6, 5), dtype=np.uint8) for c in cs]
    got = darknet.forward_route_layer_quant_gpu(xs)
    assert np.array_equal(got, np.concatenate(xs, axis=1))


@pytest.mark.parametrize("cs,ups,hw", [((128, 256), (2, 1), (26, 26)), ((16, 16), (1, 2), (6, 10)), ((3, 5, 8), (2, 1, 2), (4, 6)), ((32,), (3,), (9, 6))])
def test_route_reads_upsampled_inputs(built, cs, ups, hw):
    # an upsample -> route pair as one launch: equals the oracle's upsample followed by the plain concat
    rng = np.random.default_rng(sum(cs) + sum(ups))
    xs = [rng.integers(0, 256, size=(2, c, hw[0] // u, hw[1] // u), dtype=np.uint8) for c, u in zip(cs, ups)]
    got = darknet.forward_route_layer_quant_gpu(xs, ups)
    want = np.concatenate([np.stack([O.upsample(x[b], u) for b in range(2)]) if u > 1 else x for x, u in zip(xs, ups)], axis=1)
    assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.parametrize("cs,ups,hw", [((128, 256), (2, 1), (26, 26)), ((3, 5, 8), (2, 1, 2), (4, 6)), ((16, 32, 16), (1, 1, 1), (5, 7))])
def test_route_input_by_input(built, cs, ups, hw):
    # the inputs of one route written by separate launches: each launch touches its own channels only
    rng = np.random.default_rng(sum(cs) * 7 + sum(ups))
    xs = [rng.integers(0, 256, size=(2, c, hw[0] // u, hw[1] // u), dtype=np.uint8) for c, u in zip(cs, ups)]
    want = np.concatenate([np.stack([O.upsample(x[b], u) for b in range(2)]) if u > 1 else x for x, u in zip(xs, ups)], axis=1)
    n = len(cs)
    got = darknet.forward_route_layer_quant_gpu(xs, ups, parts=[1 << k for k in reversed(range(n))], fill=0xEE)
    assert np.array_equal(got, want)
    only_last = darknet.forward_route_layer_quant_gpu(xs, ups, parts=[1 << (n - 1)], fill=0xEE)
    off = sum(cs[:-1])
    assert np.array_equal(only_last[:, off:], want[:, off:]) and np.all(only_last[:, :off] == 0xEE)
    if n > 2:
        ends = darknet.forward_route_layer_quant_gpu(xs, ups, parts=[1 | 1 << (n - 1)], fill=0xEE)   # a mask with a gap
        assert np.array_equal(ends[:, :cs[0]], want[:, :cs[0]]) and np.array_equal(ends[:, off:], want[:, off:])
        assert np.all(ends[:, cs[0]:off] == 0xEE)


@pytest.mark.parametrize("cs,ups,hw", [((128, 256), (2, 1), (26, 26)), ((16, 16), (1, 2), (6, 10)), ((48, 16, 32), (2, 1, 2), (4, 6)), ((32,), (3,), (9, 6))])
def test_route_rows_kernel(built, monkeypatch, cs, ups, hw):
    # YQ_ROUTE_ROWS=1: one block per (output row, input); off by default until it has been timed
    monkeypatch.setenv("YQ_ROUTE_ROWS", "1")
    rng = np.random.default_rng(sum(cs) + 3 * sum(ups))
    xs = [rng.integers(0, 256, size=(2, c, hw[0] // u, hw[1] // u), dtype=np.uint8) for c, u in zip(cs, ups)]
    want = np.concatenate([np.stack([O.upsample(x[b], u) for b in range(2)]) if u > 1 else x for x, u in zip(xs, ups)], axis=1)
    assert np.array_equal(darknet.forward_route_layer_quant_gpu(xs, ups), want)
    n = len(cs)
    assert np.array_equal(darknet.forward_route_layer_quant_gpu(xs, ups, parts=[1 << k for k in range(n)], fill=0xEE), want)


def test_yolo(built):
    x = (np.random.default_rng(0).standard_normal((3, 30, 13, 13)) * 3).astype(np.float32)
    got = darknet.forward_yolo_layer_gpu(x, 3, 5)
    for b in range(3):
        ref = O.yolo(x[b], 3, 5)
        assert np.allclose(got[b], ref, atol=YOLO_ATOL, rtol=0)
        assert np.array_equal(got[b][2::10], x[b][2::10]) and np.array_equal(got[b][3::10], x[b][3::10])


def _net_vs_oracle(net, info, imgs, check_acc=True):
    heads = net.split_heads(net.predict_u8(imgs))
    for b in range(imgs.shape[0]):
        ref = O.forward_network(info, imgs[b])
        hi = 0
        for i, (sl, r) in enumerate(zip(info, ref)):
            if sl.kind == "conv" and check_acc:
                assert np.array_equal(net.pull_layer(i, "acc")[b], r["acc"]), f"layer {i} int32 mismatch (image {b})"
            if sl.kind == "yolo":
                assert np.allclose(heads[hi][b], r["f32"], atol=YOLO_ATOL, rtol=0), f"yolo layer {i}"
                hi += 1
            else:
                assert np.array_equal(net.pull_layer(i, "u8")[b], r["u8"]), f"layer {i} uint8 mismatch (image {b})"
            if sl.kind == "conv" and sl.spec.quant_stop:
                assert np.array_equal(net.pull_layer(i, "f32")[b], r["f32"]), f"layer {i} dequant mismatch"


def test_tiny416_batch1_vs_oracle_and_reference_golden(built, tiny_net_files):
    """BASELINE configs[1]: yolov3-tiny INT8 per-channel, batch 1, bit-exact int32 vs the CPU path."""
    cfg, wts, info, _ = tiny_net_files
    gold = json.load(open(os.path.join(GOLD, "tiny416.json")))
    net = darknet.load_network(cfg, wts, batch=1)
    net.set_debug(True)
    im = synth.synthetic_image(gold["seed_image"])[None]
    _net_vs_oracle(net, info, im)
    # and directly against the SHA-256 of the compiled reference's dumps
    heads = iter(net.split_heads(net.predict_u8(im)))
    for e in gold["layers"]:
        i = e["index"]
        if "output_int32" in e:
            assert sha(net.pull_layer(i, "acc")[0]) == e["output_int32"], f"layer {i} int32 vs reference"
        if "output_uint8" in e:
            assert sha(net.pull_layer(i, "u8")[0]) == e["output_uint8"], f"layer {i} uint8 vs reference"
        if e["type"] == "conv":
            p = net.conv_params(i)
            assert sha(p["M0"]) == e["M0"] and sha(p["biases_int32"]) == e["biases_int32"], f"layer {i} host prep"
            assert sha(p["M0_right_shift"]) == e["M0_right_shift"]
    net.free()


def test_fused_network_equals_unfused_and_oracle(built, tiny_net_files):
    """production configuration (fusion on, no debug buffers): every tensor that is still materialised and the
    yolo heads equal the oracle; heads equal the unfused run bit for bit."""
    cfg, wts, info, _ = tiny_net_files
    im = synth.synthetic_image(31)[None]
    net = darknet.load_network(cfg, wts, batch=1)
    fused_flat = net.predict_u8(im).copy()
    assert net.launches_per_forward < 1 + 24 - 1          # pools folded into conv launches
    ref = O.forward_network(info, im[0])
    for i, (sl, r) in enumerate(zip(info, ref)):
        # (the conv tensors of layers 0, 2, 4, 6 never exist: their launches write the pooled tensor only; nor does the
        # upsampled tensor of layer 19: route 20 reads layer 18 through the upsample, and pulling it fails loudly)
        if sl.kind in ("upsample", "route") and net.layers()[i].fused:
            # (likewise route 20: layer 21 reads [layer 18 through the upsample | layer 8] itself)
            with pytest.raises(Exception, match="not materialised"):
                net.pull_layer(i, "u8")
        elif sl.kind == "conv" and net.layers()[i].fused == 6:
            # (layer 18: its launch writes the upsampled tensor, the first part of what layer 21 reads)
            with pytest.raises(Exception, match="not materialised"):
                net.pull_layer(i, "u8")
        elif sl.kind == "maxpool" or (sl.kind == "conv" and i >= 8) or sl.kind in ("route", "upsample"):
            assert np.array_equal(net.pull_layer(i, "u8")[0], r["u8"]), f"layer {i}"
    for h, i in zip(net.split_heads(fused_flat), (16, 23)):
        assert np.allclose(h[0], ref[i]["f32"], atol=YOLO_ATOL, rtol=0)
    net.set_fusion(False)
    assert np.array_equal(net.predict_u8(im), fused_flat)
    net.free()


def test_full_size_batch_128_equals_per_image_runs_and_oracle(built, tiny_net_files):
    """BASELINE's bench configuration (batch 128 at 416x416, production plan, CUDA graph): every image's heads equal the
    heads of the same image in a batch-4 run bit for bit (tiles of the persistent kernels span image boundaries: no
    cross-talk), whatever its slot in the batch, and two of them equal the oracle."""
    cfg, wts, info, _ = tiny_net_files
    four = np.stack([synth.synthetic_image(s) for s in (41, 42, 43, 44)])
    slot = np.random.default_rng(9).integers(0, 4, size=128)
    slot[:4] = (0, 1, 2, 3)
    small = darknet.load_network(cfg, wts, batch=4)
    heads4 = [h.copy() for h in small.split_heads(small.predict_u8(four))]
    small.free()
    net = darknet.load_network(cfg, wts, batch=128)
    net.use_graph(True)
    for _ in range(2):                                   # second pass = graph replay
        heads128 = net.split_heads(net.predict_u8(four[slot]))
        for h4, h128 in zip(heads4, heads128):
            assert np.array_equal(h128, h4[slot])
    net.free()
    for img in (0, 3):
        ref = O.forward_network(info, four[img])
        for h4, i in zip(heads4, (16, 23)):
            assert np.allclose(h4[img], ref[i]["f32"], atol=YOLO_ATOL, rtol=0)


def test_tiny96_leaky_vs_oracle(built, tmp_path):
    """leaky net: zp_in = 40 padding in every conv after layer 0, leaky epilogue; GPU == exact-integer spec."""
    layers = synth.yolov3_tiny_quant("leaky")
    cfg, wts = str(tmp_path / "l.cfg"), str(tmp_path / "l.weights")
    synth.write_cfg(cfg, layers, batch=2, width=96, height=96)
    info = synth.write_weights(wts, layers, width=96, height=96, seed=3, identity_bn=False)
    net = darknet.load_network(cfg, wts, batch=2)
    net.set_debug(True)
    imgs = np.stack([synth.synthetic_image(s, 3, 96, 96) for s in (5, 6)])
    _net_vs_oracle(net, info, imgs)
    net.free()


@pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref binary not shipped")
def test_tiny416_vs_compiled_reference_live(built, tiny_net_files, tmp_path):
    """the compiled reference itself, run on this box, against the CUDA path (fresh image seed)"""
    cfg, wts, info, _ = tiny_net_files
    im = synth.synthetic_image(77)
    synth.image_to_float(im).tofile(str(tmp_path / "img.f32"))
    O.run_reference("net", cfg, wts, str(tmp_path / "img.f32"), str(tmp_path / "dump"))
    ref = O.read_dump(str(tmp_path / "dump"))
    net = darknet.load_network(cfg, wts, batch=1)
    net.set_debug(True)
    heads = iter(net.split_heads(net.predict_u8(im[None])))
    for r in ref:
        i = r["index"]
        if r["type"] == "conv":
            # relu6 / s_out = 12/255 keeps the reference's float-carried accumulator inside its exact range
            # (SURVEY 0.4); were it to leave it, this assert names the layer instead of silently passing
            assert np.array_equal(r["output_int32"], net.pull_layer(i, "acc")[0]), \
                f"layer {i}: reference int32 differs from the exact-integer GPU accumulator"
        if r["type"] == "yolo":
            assert np.allclose(next(heads)[0].ravel(), r["output_f32"], atol=YOLO_ATOL, rtol=0)
        elif "output_uint8" in r:
            assert np.array_equal(net.pull_layer(i, "u8")[0], r["output_uint8"]), f"layer {i}"
    net.free()


def test_graph_replay_equals_eager_and_is_deterministic(built, tiny_net_files):
    cfg, wts, info, _ = tiny_net_files
    imgs = np.stack([synth.synthetic_image(s) for s in (1, 2, 3)])
    net = darknet.load_network(cfg, wts, batch=3)
    a = net.predict_u8(imgs).copy()
    net.use_graph(True)
    b = net.predict_u8(imgs).copy()
    c = net.predict_u8(imgs).copy()
    assert np.array_equal(a, b) and np.array_equal(b, c)
    net.free()


def test_pipelined_submit_collect_equals_predict(built, tiny_net_files):
    """yq_network_submit_u8 / yq_network_collect (2-deep pipeline) == yq_network_predict_u8, batch by batch."""
    cfg, wts, info, _ = tiny_net_files
    net = darknet.load_network(cfg, wts, batch=2)
    net.use_graph(True)
    batches = [np.stack([synth.synthetic_image(10 * k + s) for s in (1, 2)]) for k in range(4)]
    want = [net.predict_u8(b).copy() for b in batches]
    outs = [np.empty(net.output_floats, np.float32) for _ in batches]
    inflight = []
    for k, b in enumerate(batches):
        b = np.ascontiguousarray(b)
        inflight.append((net.submit_raw(b.ctypes.data), k, b))
        if len(inflight) == 2:
            sl, j, _keep = inflight.pop(0)
            net.collect_raw(sl, outs[j].ctypes.data)
    for sl, j, _keep in inflight:
        net.collect_raw(sl, outs[j].ctypes.data)
    for w, o in zip(want, outs):
        assert np.array_equal(w, o)
    from yolo_quantization_b200._lib import YqError
    s0 = net.submit_raw(batches[0].ctypes.data)
    s1 = net.submit_raw(batches[1].ctypes.data)
    with pytest.raises(YqError, match="in flight"):
        net.submit_raw(batches[2].ctypes.data)
    net.collect_raw(s0, outs[0].ctypes.data)
    net.collect_raw(s1, outs[1].ctypes.data)
    net.free()


def test_layout_transform_roundtrip(built):
    """NCHW -> NHWC (both the c<=4 fast path and the generic one) and back."""
    rng = np.random.default_rng(4)
    for c, h, w in ((3, 8, 12), (3, 5, 7), (4, 4, 4), (16, 6, 6), (30, 5, 3), (1, 2, 2)):
        x = rng.integers(0, 256, size=(2, c, h, w), dtype=np.uint8)
        d = darknet.push_nchw_u8(x)
        cs = darknet.channel_stride(c)
        raw = d.pull((2, h, w, cs), np.uint8)
        assert np.array_equal(raw[..., :c], x.transpose(0, 2, 3, 1))
        assert not raw[..., c:].any()                      # pad channels are zero
        assert np.array_equal(darknet.pull_nhwc_u8(d, 2, c, h, w), x)
        d.free()


def test_input_quantiser_vs_oracle(built):
    """the layer-0 dynamic input quantiser on the device == the oracle restatement of blas.c:108-168
    (which tests/test_oracle_vs_reference_live.py pins against the compiled reference), per image, bit for bit."""
    rng = np.random.default_rng(12)
    imgs = [rng.random((3, 40, 56), dtype=np.float32),                                   # [0, 1): zp = 0
            (rng.standard_normal((3, 40, 56)) * 0.7).astype(np.float32),                 # negative values: zp != 0
            synth.image_to_float(synth.synthetic_image(5, 3, 40, 56)),                   # the benchmark's image class (u8 / 255)
            np.full((3, 40, 56), 0.25, np.float32)]                                      # constant image
    imgs[3][0, 0, 0] = -0.5
    x = np.stack(imgs)
    u8, scales, zps = darknet.quantize_input(x)
    for b in range(len(imgs)):
        ru8, rs, rz = O.quantize_input(x[b])
        assert scales[b] == np.float32(rs) and int(zps[b]) == int(rz), (b, scales[b], rs, zps[b], rz)
        assert np.array_equal(u8[b], ru8), f"image {b}: {np.count_nonzero(u8[b] != ru8)} bytes differ"


def test_packed_weight_arena_round_trip(built, tiny_net_files, tmp_path):
    """the first load builds the kernel-layout filter images and writes the arena; the second load of the same
    weights fetches every image (no miss, nothing rewritten) and the network computes the same bytes; other weights
    can only miss."""
    cfg, wts, _, layers = tiny_net_files
    arena = str(tmp_path / "tiny.yqpk")
    x = np.random.default_rng(3).integers(0, 256, size=(2, 3, 416, 416), dtype=np.uint8)
    net = darknet.load_network(cfg, wts, batch=2, packed=arena)
    first = darknet.pack_arena_stats()
    # (hits during the first load: the flat kernels of one layer share one image)
    assert first["misses"] > 0 and first["entries"] == first["misses"] and os.path.getsize(arena) > 8_000_000
    ref = net.predict_u8(x)
    net.free()
    mtime = os.stat(arena).st_mtime_ns
    net = darknet.load_network(cfg, wts, batch=2, packed=arena)
    second = darknet.pack_arena_stats()
    assert second["misses"] == 0 and second["hits"] == first["hits"] + first["misses"] and second["dirty"] == 0
    assert os.stat(arena).st_mtime_ns == mtime
    assert np.array_equal(net.predict_u8(x), ref)
    net.free()
    other = str(tmp_path / "other.weights")
    synth.write_weights(other, layers, seed=1)
    net = darknet.load_network(cfg, other, batch=2, packed=arena)
    third = darknet.pack_arena_stats()
    assert third["hits"] == first["hits"] and third["misses"] == first["misses"]
    plain = darknet.load_network(cfg, other, batch=2)
    assert np.array_equal(net.predict_u8(x), plain.predict_u8(x))
    net.free()
    plain.free()


def test_network_predict_f32_equals_predict_u8(built, tiny_net_files):
    """float images through the device quantiser + layer-0 re-prep == the uint8 path fed with the oracle-quantized image
    and the same (s_in, zp_in)."""
    cfg, wts, info, _ = tiny_net_files
    im = synth.synthetic_image(77)
    xf = synth.image_to_float(im)[None]
    ru8, rs, rz = O.quantize_input(xf[0])
    net = darknet.load_network(cfg, wts, batch=1)
    got = net.predict_f32(xf).copy()
    net.set_input_quant(float(rs), int(rz))
    want = net.predict_u8(ru8[None])
    assert np.array_equal(got, want)
    net.free()


def test_network_input_paths_agree(built, tiny_net_files, monkeypatch):
    """layer 0 fed by the in-kernel plane fetch (default) and by the layout-transform launch + padded NHWC4 copy
    (YQ_NO_PLANAR=1, also the path of nets with zp_in != 0 or w % 16 != 0): same bytes out, one launch apart."""
    cfg, wts, _, _ = tiny_net_files
    x = np.random.default_rng(17).integers(0, 256, size=(3, 3, 416, 416), dtype=np.uint8)
    net = darknet.load_network(cfg, wts, batch=3)
    planar, n_planar = net.predict_u8(x).copy(), net.launches_per_forward
    net.free()
    monkeypatch.setenv("YQ_NO_PLANAR", "1")
    net = darknet.load_network(cfg, wts, batch=3)
    assert np.array_equal(net.predict_u8(x), planar) and net.launches_per_forward == n_planar + 1
    net.free()


@pytest.mark.parametrize("uproute,branch,early,nocat", [(1, 0, 0, 0), (0, 1, 0, 0), (1, 1, 0, 0), (0, 1, 1, 0), (1, 1, 1, 0), (1, 0, 1, 0), (1, 1, 1, 1), (1, 1, 0, 1)])
def test_network_schedule_switches_keep_every_byte(built, tiny_net_files, monkeypatch, uproute, branch, early, nocat):
    """YQ_UPROUTE (upsample folded into the route behind it: one launch less), YQ_BRANCH_STREAM (the first detection
    head on a second stream beside the layers after it) and YQ_EARLY_ROUTE (layer 8 copied into route 20's tensor on
    that stream right behind layer 8: one launch more) change the schedule only: same bytes out, eager and replayed.
    With the upsample folded in, layer 21 reads [layer 18 through the upsample | layer 8] itself (route 20 is never
    written, there is nothing to copy early) unless YQ_NO_CAT=1 keeps the materialised route."""
    cfg, wts, _, _ = tiny_net_files
    x = np.random.default_rng(23).integers(0, 256, size=(4, 3, 416, 416), dtype=np.uint8)
    for k in ("YQ_UPROUTE", "YQ_BRANCH_STREAM", "YQ_EARLY_ROUTE"):
        monkeypatch.setenv(k, "0")
    net = darknet.load_network(cfg, wts, batch=4)
    base, n_base = net.predict_u8(x).copy(), net.launches_per_forward
    net.free()
    monkeypatch.setenv("YQ_UPROUTE", str(uproute))
    monkeypatch.setenv("YQ_BRANCH_STREAM", str(branch))
    monkeypatch.setenv("YQ_EARLY_ROUTE", str(early))
    monkeypatch.setenv("YQ_NO_CAT", str(nocat))
    net = darknet.load_network(cfg, wts, batch=4)
    cat = uproute and not nocat
    assert bool(net.layers()[20].fused) == bool(cat) and (net.layers()[18].fused == 6) == bool(cat)
    # (with the route gone, layer 18 writes the upsampled tensor itself: the route's launch goes too)
    assert net.launches_per_forward == n_base - uproute - (1 if cat else 0) + (1 if early and branch and not cat else 0)
    for graph in (False, True):
        net.use_graph(graph)
        for _ in range(4):
            assert np.array_equal(net.predict_u8(x), base)
    y = np.random.default_rng(24).integers(0, 256, size=(4, 3, 416, 416), dtype=np.uint8)
    other = net.predict_u8(y).copy()
    assert not np.array_equal(other, base) and np.array_equal(net.predict_u8(x), base)
    net.free()


def test_layout_transform_padded_geometry(built):
    """NCHW -> halo-padded NHWC and back: the interior round-trips, the halo keeps what the runtime put there."""
    import ctypes as C
    from yolo_quantization_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(5)
    for c, h, w, pad, pitch, rows in ((3, 8, 12, 1, 16, 10), (3, 6, 20, 1, 24, 9), (3, 5, 7, 1, 12, 8), (16, 6, 6, 1, 9, 8), (3, 4, 8, 2, 13, 9),
                                      (3, 6, 32, 1, 40, 9), (4, 5, 48, 1, 52, 8), (1, 7, 16, 1, 20, 9), (3, 9, 64, 1, 68, 11)):   # 16-pixel fast path
        x = rng.integers(0, 256, size=(2, c, h, w), dtype=np.uint8)
        g = _lib.ActGeom(pad, pitch, rows)
        cs = darknet.channel_stride(c)
        src = darknet.DeviceBuffer.from_numpy(x)
        dst = darknet.DeviceBuffer(lib.yq_act_geom_bytes(C.byref(g), 2, c), zero=False)
        _lib.check(lib.yq_cuda_memset(dst.ptr, 0xA5, dst.nbytes, None))
        _lib.check(lib.yq_nchw_to_nhwc_u8_geom(src.ptr, dst.ptr, 2, c, h, w, C.byref(g), None))
        raw = dst.pull((2, rows, pitch, cs), np.uint8)
        assert np.array_equal(raw[:, pad:pad + h, pad:pad + w, :c], x.transpose(0, 2, 3, 1))
        halo = raw.copy()
        halo[:, pad:pad + h, pad:pad + w, :] = 0xA5
        assert (halo == 0xA5).all(), "the transform wrote outside the interior"
        back = darknet.DeviceBuffer(x.nbytes)
        _lib.check(lib.yq_nhwc_to_nchw_u8_geom(dst.ptr, back.ptr, 2, c, h, w, C.byref(g), None))
        assert np.array_equal(back.pull(x.shape, np.uint8), x)
        for d in (src, dst, back):
            d.free()


def test_full_size_batch128_properties(built, tiny_net_files):
    """BASELINE configs[2] size (batch 128 @ 416x416) through size-independent properties:
    batch-independence (image i of the batch == the same image alone), permutation equivariance and
    idempotence; a sample of images is also checked against the oracle."""
    cfg, wts, info, _ = tiny_net_files
    B = 128
    rng = np.random.default_rng(2024)
    imgs = rng.integers(0, 256, size=(B, 3, 416, 416), dtype=np.uint8)
    net = darknet.load_network(cfg, wts, batch=B)
    net.use_graph(True)
    flat = net.predict_u8(imgs).copy()
    heads = [h.copy() for h in net.split_heads(flat)]
    # idempotence
    assert np.array_equal(net.predict_u8(imgs), flat)
    # permutation equivariance
    perm = rng.permutation(B)
    heads_p = net.split_heads(net.predict_u8(imgs[perm]))
    for h, hp in zip(heads, heads_p):
        assert np.array_equal(h[perm], hp)
    l21 = net.pull_layer(21, "u8")   # of the permuted run
    net.free()
    # batch independence + oracle on a sample
    net1 = darknet.load_network(cfg, wts, batch=1)
    for b in (0, 63, 127):
        single = net1.split_heads(net1.predict_u8(imgs[b:b + 1]))
        for h, s in zip(heads, single):
            assert np.array_equal(h[b], s[0])
    ref = O.forward_network(info, imgs[perm[5]])
    assert np.array_equal(l21[5], ref[21]["u8"])
    assert np.allclose(heads_p[1][5], ref[23]["f32"], atol=YOLO_ATOL, rtol=0)
    net1.free()


def test_device_box_decode_and_nms_vs_oracle(built, tiny_net_files):
    """row 8f-2: get_network_boxes + do_nms_sort on the device == the oracle restatement (which equals the compiled
    reference incl. its stable-sort tie behaviour, tests/test_oracle_golden.py).  Coordinates within 1e-5, the set of
    surviving (box, class) pairs identical."""
    cfg, wts, info, _ = tiny_net_files
    imgs = np.stack([synth.synthetic_image(s) for s in (1, 41, 42)])
    net = darknet.load_network(cfg, wts, batch=3)
    heads = net.split_heads(net.predict_u8(imgs))
    for nms in (0.0, 0.45):
        got = net.get_boxes(416, 416, 0.5, nms, 1)
        for b in range(3):
            want = O.yolo_boxes([heads[0][b], heads[1][b]], [(3, 4, 5), (0, 1, 2)], 5, 416, 416, 416, 416, 0.5, nms)
            assert got[b].shape == want.shape, (nms, b, got[b].shape, want.shape)
            assert np.allclose(got[b][:, :5], want[:, :5], rtol=1e-5, atol=1e-6)
            assert np.array_equal(got[b][:, 5:] > 0, want[:, 5:] > 0), f"NMS survivors differ (image {b}, nms {nms})"
            assert np.allclose(got[b][:, 5:], want[:, 5:], rtol=1e-6, atol=0)
    # non-square "original image" exercises correct_yolo_boxes' letterbox arithmetic, absolute coordinates
    got = net.get_boxes(640, 480, 0.6, 0.0, 0)
    want = O.yolo_boxes([heads[0][1], heads[1][1]], [(3, 4, 5), (0, 1, 2)], 5, 416, 416, 640, 480, 0.6, 0.0, relative=0)
    assert got[1].shape == want.shape and np.allclose(got[1], want, rtol=1e-5, atol=1e-4)
    net.free()


def test_accumulator_linearity(built):
    """int32 accumulators are linear in the input when zp_in = 0: acc(2x) == 2 acc(x) for x <= 127."""
    rng = np.random.default_rng(9)
    c, n, k = 64, 64, 3
    wq = rng.integers(0, 256, size=(n, c * k * k), dtype=np.uint8)
    zp_w = rng.integers(0, 256, size=n, dtype=np.uint8)
    ones = np.ones(n)
    layer = darknet.ConvolutionalLayerQuant(10, 10, c, n, k, 1, 1, synth.ACT_CODES["linear"], wq, zp_w, np.zeros(n, np.int32),
                                            ones * 0.5, ones, 0, 0, 1.0)
    x = rng.integers(0, 128, size=(1, c, 10, 10), dtype=np.uint8)
    a1 = layer.forward(x)["acc"].astype(np.int64)
    a2 = layer.forward((x * 2).astype(np.uint8))["acc"].astype(np.int64)
    assert np.array_equal(2 * a1, a2)
    layer.free()


def test_error_paths(built, tmp_path):
    from yolo_quantization_b200._lib import YqError
    with pytest.raises(YqError, match="cannot open cfg"):
        darknet.load_network(str(tmp_path / "missing.cfg"), str(tmp_path / "missing.weights"))
    layers = synth.yolov3_tiny_quant()
    cfg, wts = str(tmp_path / "t.cfg"), str(tmp_path / "t.weights")
    synth.write_cfg(cfg, layers, width=64, height=64)
    synth.write_weights(wts, layers, width=64, height=64)
    with open(wts, "r+b") as f:
        f.truncate(1000)
    with pytest.raises(YqError, match="truncated"):
        darknet.load_network(cfg, wts)
    # multiplier outside (0,1): the reference assert()s (blas.c:391-392); we report it
    synth.write_weights(wts, layers, width=64, height=64, relu6_scale=1e-9)
    with pytest.raises(YqError, match="outside"):
        darknet.load_network(cfg, wts)
