"""CPU: the oracle restatement (oracle/yq_oracle.c) against the golden fixtures generated from the
COMPILED REFERENCE (oracle/gen_golden.py).  This is what pins the oracle; the CUDA path is then pinned
to the oracle (tests/test_gpu_parity.py)."""
import glob
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import yq_oracle as O
from yolo_quantization_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _check_net(gold, outs, strict_acc=True):
    bad = []
    for o, e in zip(outs, gold["layers"]):
        for mine, key in (("acc", "output_int32"), ("u8", "output_uint8"), ("f32", "output_f32"),
                          ("M0", "M0"), ("M0_right_shift", "M0_right_shift"), ("biases_int32", "biases_int32")):
            if mine in o and key in e and sha(o[mine]) != e[key]:
                bad.append((e["index"], key))
    return bad


def test_tiny416_exact_integer_oracle_equals_reference(built):
    """relu6 net, 2^24-safe distribution: exact-integer oracle == reference on all 24 layers."""
    gold = json.load(open(os.path.join(GOLD, "tiny416.json")))
    info = synth.write_weights(None, synth.yolov3_tiny_quant(), seed=gold["seed_weights"])
    im = synth.synthetic_image(gold["seed_image"])
    assert sha(im) == gold["input_sha256"]
    outs = O.forward_network(info, im)
    assert _check_net(gold, outs) == []


def test_tiny96_leaky_float_carried_oracle_equals_reference(built):
    """leaky net (zp_in = 40 padding, K up to 4608): the reference's float-carried accumulator deviates from
    exact integers; the float-carried restatement reproduces it bit for bit, the exact one does not."""
    gold = json.load(open(os.path.join(GOLD, "tiny96_leaky.json")))
    info = synth.write_weights(None, synth.yolov3_tiny_quant("leaky"), width=96, height=96, seed=gold["seed_weights"])
    im = synth.synthetic_image(gold["seed_image"], 3, 96, 96)
    assert sha(im) == gold["input_sha256"]
    assert _check_net(gold, O.forward_network(info, im, reffloat=True)) == []
    bad_exact = _check_net(gold, O.forward_network(info, im, reffloat=False))
    # layers 0..9 (K <= 1152) are 2^24-safe and agree; the first deviation is in a K >= 2304 layer
    assert bad_exact and min(i for i, _ in bad_exact) >= 10


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "layer_*.npz"))), ids=os.path.basename)
def test_single_layer_cases(built, path):
    g = np.load(path)
    c, h, w, n, k, stride, pad, act, bn, qs, zp_in, zp_out = (int(v) for v in g["geom"])
    s_in, s_out = (float(v) for v in g["scales"])
    acc = O.conv_acc(g["x"], g["w_u8"], g["zp_w"], stride, pad, zp_in)
    assert np.array_equal(acc, g["ref_int32"])
    u8 = O.requant(acc, g["biases_int32"], g["M_value"], g["M0_right_shift_value"], act, zp_out)
    assert np.array_equal(u8, g["ref_uint8"])
    if qs:
        assert np.array_equal(O.dequant(u8, zp_out, s_out), g["ref_f32"].reshape(u8.shape))
    # the float-carried restatement must agree too on these small-K cases
    assert np.array_equal(O.conv_acc(g["x"], g["w_u8"], g["zp_w"], stride, pad, zp_in, reffloat=True), g["ref_int32"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "layer_*.npz"))), ids=os.path.basename)
def test_prep_restatement_matches_reference_params(built, path):
    """quantization_weights_and_activations restated (blas.c:282-334) vs the reference's own M0 / shift / bias,
    including non-identity batch-norm folding."""
    g = np.load(path)
    c, h, w, n, k, stride, pad, act, bn, qs, zp_in, zp_out = (int(v) for v in g["geom"])
    s_in, s_out = (np.float32(v) for v in g["scales"])
    spec = synth.LayerSpec("conv", n, k, stride, 1, bn, "linear")
    sl = synth.SynthLayer("conv", c, h, w, n, 0, 0, spec, s_out=float(s_out), biases=g["biases"], bn_scales=g["bn_scales"],
                          bn_mean=g["bn_mean"], bn_var=g["bn_var"], s_w=g["s_w"], zp_w=g["zp_w"], w_u8=g["w_u8"])
    p = O.prepare_conv(sl, float(s_in), zp_in)
    assert np.array_equal(p["M0"], g["M0"]) and np.array_equal(p["M0_right_shift"], g["M0_right_shift"])
    assert np.array_equal(p["M_value"], g["M_value"])
    assert np.array_equal(p["M0_right_shift_value"], g["M0_right_shift_value"])
    assert np.array_equal(p["biases_int32"], g["biases_int32"])


def test_wrap_case_really_wraps():
    """the relu_3x3_wrap fixture must exercise uint8 wrap (q + zp_out outside [0,255])."""
    g = np.load(os.path.join(GOLD, "layer_relu_3x3_wrap.npz"))
    acc = g["ref_int32"].astype(np.int64) + g["biases_int32"].astype(np.int64)[:, None, None]
    q = np.trunc(np.trunc(acc * g["M_value"][:, None, None]) * g["M0_right_shift_value"][:, None, None]) + int(g["geom"][11])
    assert ((q < 0) | (q > 255)).sum() > 10


def test_maxpool_upsample_yolo_small_cases(built):
    rng = np.random.default_rng(0)
    x = rng.integers(0, 256, size=(5, 13, 13), dtype=np.uint8)
    # 2x2 stride 1 on 13x13 (layer 11 of yolov3-tiny): right/bottom taps out of bounds are ignored
    y = O.maxpool(x, 2, 1)
    assert y.shape == (5, 13, 13)
    assert y[0, 12, 12] == x[0, 12, 12] and y[0, 0, 0] == x[0, :2, :2].max()
    y2 = O.maxpool(x[:, :12, :12], 2, 2)
    assert np.array_equal(y2, x[:, :12, :12].reshape(5, 6, 2, 6, 2).max(axis=(2, 4)))
    u = O.upsample(x, 2)
    assert np.array_equal(u, x.repeat(2, 1).repeat(2, 2))
    f = rng.standard_normal((30, 4, 4)).astype(np.float32)
    yo = O.yolo(f, 3, 5)
    sig = (1.0 / (1.0 + np.exp(-f.astype(np.float64)))).astype(np.float32)
    for a in range(3):
        for e in range(10):
            exp = f[a * 10 + e] if e in (2, 3) else sig[a * 10 + e]
            assert np.array_equal(yo[a * 10 + e], exp)
