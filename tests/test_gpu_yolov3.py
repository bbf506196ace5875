"""GPU (B200): BASELINE configs[4] -- the full yolov3 (75 conv, 23 shortcut, 4 route, 2 upsample, 3 yolo) as a network.

* every distinct conv shape at its REAL spatial size (416 ... 13) against the oracle and, live, against the compiled
  reference through the one-layer trick on a 2^24-safe input distribution;
* the stride-2 down-samplers on the tcgen05 per-tap flavour (TMA box with elementStrides = 2), between plain, flat and
  padded tensors, reading their padding from the halo or restoring it with the border-correction table;
* the quantized shortcut (this repo's integer extension -- the reference has none, SURVEY 0.10: pinned by the oracle's
  restatement of the spec only);
* the 107-layer network against the oracle's layer walk, debug plan (every tensor + int32 accumulators) and production plan.
Bit-exact for every integer / byte tensor; yolo floats within 1e-6 abs."""
import zlib

import numpy as np
import pytest

from oracle import yq_oracle as O
from yolo_quantization_b200 import darknet, synth

pytestmark = pytest.mark.gpu
YOLO_ATOL = 1e-6


def _rand_layer(rng, c, n, k, stride, act, zp_in, zp_out=33, s_out=0.05, h=12, w=12, quant_stop=None):
    wq = rng.integers(0, 256, size=(n, c * k * k), dtype=np.uint8)
    zp_w = rng.integers(0, 256, size=n, dtype=np.uint8)
    s_w = (rng.random(n).astype(np.float32) * 0.01 + 0.001).astype(np.float32)
    bias = (rng.standard_normal(n) * 0.5).astype(np.float32)
    spec = synth.LayerSpec("conv", n, k, stride, 1, 0, act)
    sl = synth.SynthLayer("conv", c, h, w, n, 0, 0, spec, s_out=s_out, biases=bias, s_w=s_w, zp_w=zp_w, w_u8=wq.reshape(n, c, k, k))
    p = O.prepare_conv(sl, 0.02, zp_in)
    qs = (1 if act == "linear" else 0) if quant_stop is None else quant_stop
    layer = darknet.ConvolutionalLayerQuant(h, w, c, n, k, stride, k // 2, synth.ACT_CODES[act], wq, zp_w, p["biases_int32"], p["M_value"],
                                            p["M0_right_shift_value"], zp_in, zp_out, s_out, quant_stop_flag=qs)
    return layer, wq.reshape(n, c, k, k), zp_w, p


def _check(got, x, wq, zp_w, p, stride, k, act, zp_in, zp_out, s_out=0.05, qs=0):
    for b in range(x.shape[0]):
        acc = O.conv_acc(x[b], wq, zp_w, stride, k // 2, zp_in)
        if "acc" in got:
            assert np.array_equal(got["acc"][b], acc), f"int32 accumulator mismatch, image {b}"
        u8 = O.requant(acc, p["biases_int32"], p["M_value"], p["M0_right_shift_value"], synth.ACT_CODES[act], zp_out)
        assert np.array_equal(got["u8"][b], u8), f"uint8 mismatch, image {b}"
        if qs:
            assert np.array_equal(got["f32"][b], O.dequant(u8, zp_out, s_out))


STRIDE2_CASES = [
    # c, h, w, n, zp_in, batch
    (64, 16, 16, 128, 40, 2),
    (128, 26, 26, 256, 17, 3),       # odd output rows per tile
    (256, 12, 20, 512, 0, 2),        # zp_in = 0: no correction table at all
    (512, 26, 26, 1024, 40, 1),      # yolov3's last down-sampler at its real size
    (64, 9, 7, 48, 200, 2),          # odd sizes: (h + 2 - 3) / 2 + 1 rows, right / bottom taps stay inside
    (192, 2, 2, 64, 9, 5),           # 1x1 outputs
]


@pytest.mark.parametrize("case", STRIDE2_CASES, ids=lambda c: "c%d_%dx%d_n%d_zi%d" % c[:5])
def test_conv_stride2_tcgen05_vs_oracle(built, case):
    """forward_convolutional_layer_quant_inputi_outputi is generic in stride (convolutional_layer.c:699-716, im2col.c:26-50);
    3x3 / stride 2 with c % 64 == 0 runs the per-tap tcgen05 flavour.  Four tensor situations, all == oracle:
    plain -> plain (zero fill + border correction), flat -> flat (padding read from the halo), padded with a foreign halo
    byte -> plain (correction on a padded tensor), plain -> a wide padded tensor (only the interior may be written)."""
    c, h, w, n, zp_in, batch = case
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()) + 11)
    layer, wq, zp_w, p = _rand_layer(rng, c, n, 3, 2, "leaky", zp_in, h=h, w=w)
    assert layer.kernel == 1 and layer.geom_supported
    x = rng.integers(0, 256, size=(batch, c, h, w), dtype=np.uint8)
    _check(layer.forward(x), x, wq, zp_w, p, 2, 3, "leaky", zp_in, 33)
    oh, ow = layer.out_h, layer.out_w
    for gi, go, fill in (("flat", "flat", None), ((2, w + 5, h + 3), None, (zp_in + 7) & 255), (None, (1, ow + 4, oh + 2), None),
                         ((1, w + 2, h + 2), "flat", None)):
        got = layer.forward_geom(x, gi, go, in_fill=fill)
        assert got["halo_ok"], f"halo of the output was written ({gi} -> {go})"
        _check(got, x, wq, zp_w, p, 2, 3, "leaky", zp_in, 33)
    layer.free()


PATCH_CASES = [
    # c, h, w, n, zp_in, batch
    (32, 32, 32, 64, 40, 2),         # layer 1 of the full yolov3 in small: c = 32 (32-byte rows), whole filter per ring stage, one commit per tile
    (64, 32, 48, 128, 17, 2),        # layer 5 in small: c = 64, a filter row per stage
    (32, 37, 45, 64, 0, 3),          # odd sizes: partial tiles in x and y
    (64, 26, 20, 32, 200, 1),        # narrow output (32-byte staging rows)
    (32, 416, 416, 64, 3, 1),        # layer 1 at its real size
    (64, 208, 208, 128, 9, 1),       # layer 5 at its real size
    (96, 16, 24, 100, 5, 2),         # n = 100 -> stride 112: not a patch-mode shape
]


@pytest.mark.parametrize("form", ["default", "nine_boxes"])
@pytest.mark.parametrize("case", PATCH_CASES, ids=lambda c: "c%d_%dx%d_n%d_zi%d" % c[:5])
def test_conv_stride2_resident_bank_patch_mode_vs_oracle(built, case, form, monkeypatch):
    """narrow 3x3 / stride-2 layers (layers 1 and 5 of the full yolov3) on the resident-bank kernel in patch mode: flat input whose halo
    holds zp_in, no side outputs -> bytes equal the oracle and the older flavours (YQ_PW=0), the output's halo is untouched; with side
    outputs, a plain input or a foreign halo byte the older flavours still serve the call.  c = 32 runs the PAIR form by default (the
    input as 64-byte pixel pairs, two boxes per filter row); its nine-box form is a tested switch (slower than the small-c flavour)."""
    c, h, w, n, zp_in, batch = case
    if form == "nine_boxes":
        if c != 32:
            pytest.skip("only c = 32 has two forms")
        monkeypatch.setenv("YQ_PWT_C32", "1")
        monkeypatch.setenv("YQ_NO_PWT_PAIR", "1")
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()) + 19)
    layer, wq, zp_w, p = _rand_layer(rng, c, n, 3, 2, "leaky", zp_in, h=h, w=w)
    x = rng.integers(0, 256, size=(batch, c, h, w), dtype=np.uint8)
    assert layer.patch_supported == (n != 100)
    if n == 100:
        _check(layer.forward(x), x, wq, zp_w, p, 2, 3, "leaky", zp_in, 33)
        layer.free()
        return
    if c % 64 == 0:      # (a c = 32 layer falls back to the small-c flavour, which reads plain tensors only)
        oh, ow = layer.out_h, layer.out_w
        for gi, go in (("flat", "flat"), ("flat", None), ((2, w + 5, h + 3), (1, ow + 4, oh + 2))):
            got = layer.forward_geom(x, gi, go, want_acc=False)
            assert got["halo_ok"], f"halo of the output was written ({gi} -> {go})"
            _check(got, x, wq, zp_w, p, 2, 3, "leaky", zp_in, 33)
        monkeypatch.setenv("YQ_PW", "0")
        old = layer.forward_geom(x, "flat", "flat", want_acc=False)
        monkeypatch.delenv("YQ_PW")
        assert np.array_equal(old["u8"], layer.forward_geom(x, "flat", "flat", want_acc=False)["u8"])
        _check(layer.forward_geom(x, "flat", "flat", want_acc=True), x, wq, zp_w, p, 2, 3, "leaky", zp_in, 33)          # side output: per-tap flavour
        _check(layer.forward_geom(x, "flat", "flat", in_fill=(zp_in + 7) & 255, want_acc=False), x, wq, zp_w, p, 2, 3, "leaky", zp_in, 33)
    else:
        for gi, go in (("flat", "flat"), ("flat", None), ((2, w + 5, h + 3), "flat")):
            got = layer.forward_geom(x, gi, go, want_acc=False)
            assert got["halo_ok"], f"halo of the output was written ({gi} -> {go})"
            _check(got, x, wq, zp_w, p, 2, 3, "leaky", zp_in, 33)
        _check(layer.forward(x), x, wq, zp_w, p, 2, 3, "leaky", zp_in, 33)                                                # plain: small-c flavour
    layer.free()


@pytest.mark.parametrize("case", [(64, 13, 13, 128, 3, 40), (128, 9, 11, 64, 1, 5), (256, 6, 6, 255, 1, 0)], ids=lambda c: "c%d_%dx%d_n%d_k%d" % c[:5])
def test_conv_stride1_per_tap_between_padded_tensors(built, case, monkeypatch):
    """the per-tap flavour at stride 1 between tensors of any geometry (what a conv next to a conflicting tensor falls to)"""
    c, h, w, n, k, zp_in = case
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()) + 12)
    act = "linear" if n == 255 else "leaky"
    layer, wq, zp_w, p = _rand_layer(rng, c, n, k, 1, act, zp_in, h=h, w=w)
    x = rng.integers(0, 256, size=(2, c, h, w), dtype=np.uint8)
    for gi, go, fill in (("flat", "flat", None), ("flat", None, zp_in ^ 0x55), (None, (2, w + 4, h + 4), None)):
        got = layer.forward_geom(x, gi, go, in_fill=fill)
        assert got["halo_ok"]
        _check(got, x, wq, zp_w, p, 1, k, act, zp_in, 33, qs=1 if act == "linear" else 0)
    layer.free()


FLAT2_CASES = [
    # c, h, w, n, k, act, zp_in, batch      (the persistent two-tile flat form beyond yolov3-tiny's shapes)
    (256, 13, 13, 128, 1, "leaky", 40, 3),     # 1x1: one weight stage per patch chunk
    (512, 9, 9, 256, 1, "leaky", 7, 2),        # 1x1, two n-tiles, four K chunks
    (128, 20, 20, 64, 1, "leaky", 40, 2),      # 1x1, n = 64: half a tile of channels is stored (SWIZZLE_64B staging)
    (64, 24, 24, 32, 1, "relu6", 0, 2),        # 1x1, n = 32 (SWIZZLE_32B staging), KC = 64
    (64, 6, 104, 128, 3, "leaky", 40, 2),      # 3x3 on 104-wide rows: the patch arrives as two 240-row boxes
    (64, 4, 208, 128, 3, "leaky", 33, 1),      # 3x3 on 208-wide rows: three boxes
    (128, 5, 104, 64, 3, "leaky", 9, 2),       # wide rows, KC = 128, n = 64
    (128, 3, 150, 120, 3, "linear", 200, 1),   # n = 120: pad lanes of the last 16 channels stay zero
]


@pytest.mark.parametrize("case", FLAT2_CASES, ids=lambda c: "c%d_%dx%d_n%d_k%d_%s" % c[:6])
def test_conv_flat2_wide_rows_1x1_and_narrow_outputs(built, case, monkeypatch):
    """conv_u8_tc_flat2_kernel on the shapes the full yolov3 adds: 1x1 layers (forced here: the dispatch takes it from a few waves
    of tile pairs on), rows up to 222 pixels wide (patch in up to three TMA boxes) and layers with n < 128."""
    c, h, w, n, k, act, zp_in, batch = case
    monkeypatch.setenv("YQ_FLAT2_1X1", "1")
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()) + 13)
    layer, wq, zp_w, p = _rand_layer(rng, c, n, k, 1, act, zp_in, h=h, w=w, quant_stop=0)
    assert layer.flat_supported
    x = rng.integers(0, 256, size=(batch, c, h, w), dtype=np.uint8)
    for want_acc in (True, False):                 # SLOW (side outputs) and production variants
        got = layer.forward_flat(x, halo_fill=zp_in ^ 0x33, want_acc=want_acc)
        assert got["halo_ok"], "halo / pad lanes of the output"
        _check(got, x, wq, zp_w, p, 1, k, act, zp_in, 33)
    layer.free()


# (c, n, size, stride, activation, input h = w at 416x416)
YOLOV3_REAL = [
    (3, 32, 3, 1, "leaky", 416), (32, 64, 3, 2, "leaky", 416), (64, 32, 1, 1, "leaky", 208), (32, 64, 3, 1, "leaky", 208),
    (64, 128, 3, 2, "leaky", 208), (128, 64, 1, 1, "leaky", 104), (64, 128, 3, 1, "leaky", 104),
    (128, 256, 3, 2, "leaky", 104), (256, 128, 1, 1, "leaky", 52), (128, 256, 3, 1, "leaky", 52),
    (256, 512, 3, 2, "leaky", 52), (512, 256, 1, 1, "leaky", 26), (256, 512, 3, 1, "leaky", 26),
    (512, 1024, 3, 2, "leaky", 26), (1024, 512, 1, 1, "leaky", 13), (512, 1024, 3, 1, "leaky", 13),
    (1024, 255, 1, 1, "linear", 13), (768, 256, 1, 1, "leaky", 26), (512, 255, 1, 1, "linear", 26),
    (384, 128, 1, 1, "leaky", 52), (256, 255, 1, 1, "linear", 52),
]


@pytest.mark.parametrize("shape", YOLOV3_REAL, ids=lambda s: "c%d_n%d_k%d_s%d_%s_%d" % s)
def test_full_yolov3_conv_shapes_at_real_sizes(built, shape, tmp_path):
    """Every distinct convolution of the full yolov3 at the spatial size it has in the 416x416 network, through the flavour the
    network's planner picks for it (flat / flat2 / flat2x, per-tap for the down-samplers, small-c), int32 accumulators and
    uint8 outputs == oracle.  Where oracle/_ref is present the same layer also runs through the COMPILED REFERENCE (one-layer
    cfg, l.forward called directly: SURVEY Appendix F) on the same bytes; the input stays below 25 so that the reference's
    float-carried accumulator (gemm.c:279-296) remains exact -- asserted, not assumed."""
    c, n, k, stride, act, hw = shape
    # (small values AND a small zero point: the padding value counts in the reference's running sums too -- 13 x 13 corner
    # outputs of a K = 4608 layer see 5 of 9 taps of padding)
    s_in, zp_in = 0.03, 8 if c > 3 else 0
    layers = synth.single_conv(n, k, stride, act, 1, 1 if act == "linear" else 0, act_scale=0.05, act_zp=33)
    cfg, wts, img = (str(tmp_path / f) for f in ("l.cfg", "l.weights", "img.f32"))
    synth.write_cfg(cfg, layers, width=hw, height=hw, channels=c)
    info = synth.write_weights(wts, layers, width=hw, height=hw, channels=c, seed=hw + c, input_quant=(s_in, zp_in), identity_bn=False)
    rng = np.random.default_rng(zlib.crc32(repr(shape).encode()) + 5)
    x = rng.integers(0, max(6, min(25, int(1.0e7 / (c * k * k * 160)))), size=(c, hw, hw), dtype=np.uint8)
    x.flat[0], x.flat[1] = 0, 255                  # the reference's dynamic input quantiser re-derives (s_in, zp_in) from these
    sl = info[0]
    r = None
    if O.have_reference():
        # the reference first: its own host prep (quantization_weights_and_activations, src/blas.c:259-346 -- "reused verbatim"
        # in a drop-in) supplies the per-channel parameters, exactly as a binding would hand them over in yq_conv_desc
        synth.image_to_float(x, s_in, zp_in).tofile(img)
        O.run_reference("layer", cfg, wts, img, str(tmp_path / "dump"), omp=True)
        r = O.read_dump(str(tmp_path / "dump"))[0]
        ref_in = np.fromfile(str(tmp_path / "dump" / "L00_input_uint8.bin"), dtype=np.uint8).reshape(c, hw, hw)
        assert np.array_equal(ref_in, x) and r["zp_in"] == zp_in, "the reference's input quantiser did not reproduce the test tensor"
        p = {k_: r[k_] for k_ in ("biases_int32", "M_value", "M0_right_shift_value", "M0")}
        mine = O.prepare_conv(sl, float(np.float32(r["s_in"])), zp_in)
        assert np.array_equal(mine["M0"], p["M0"]) and np.abs(mine["biases_int32"] - p["biases_int32"]).max() <= 1   # (-Ofast: oracle header)
    else:
        p = O.prepare_conv(sl, s_in, zp_in)
    qs = 1 if act == "linear" else 0
    layer = darknet.ConvolutionalLayerQuant(hw, hw, c, n, k, stride, k // 2, synth.ACT_CODES[act], sl.w_u8, sl.zp_w, p["biases_int32"], p["M_value"],
                                            p["M0_right_shift_value"], zp_in, 33, sl.s_out, quant_stop_flag=qs)
    if layer.flat_supported and stride == 1:
        got = layer.forward_flat(x[None], halo_fill=zp_in)
    elif layer.geom_supported:
        got = layer.forward_geom(x[None], "flat", "flat")
    else:
        got = layer.forward(x[None])
    layer.free()
    acc = O.conv_acc(x, sl.w_u8, sl.zp_w, stride, k // 2, zp_in)
    assert np.array_equal(got["acc"][0], acc), "int32 accumulator vs oracle"
    u8 = O.requant(acc, p["biases_int32"], p["M_value"], p["M0_right_shift_value"], synth.ACT_CODES[act], 33)
    assert np.array_equal(got["u8"][0], u8), "uint8 vs oracle"
    if r is not None:
        assert np.array_equal(r["output_int32"], got["acc"][0]), "int32 accumulator vs the compiled reference"
        assert np.array_equal(r["output_uint8"], got["u8"][0]), "uint8 vs the compiled reference"
        if qs:
            assert np.array_equal(r["output_f32"].reshape(got["f32"][0].shape), got["f32"][0]), "dequantised head vs the compiled reference"


@pytest.mark.parametrize("c,h,w,batch", [(64, 16, 16, 2), (256, 13, 13, 3), (1024, 5, 7, 1), (24, 9, 4, 2), (3, 6, 5, 2)])
def test_shortcut_vs_oracle(built, c, h, w, batch):
    """the quantized shortcut (extension; spec in include/yq_b200.h) == its plain-C restatement, incl. saturation at both ends,
    channel counts that leave pad lanes (24 -> stride 32, 3 -> stride 4: pad lanes stay zero) and halo-padded tensors."""
    rng = np.random.default_rng(c * 31 + h)
    a = rng.integers(0, 256, size=(batch, c, h, w), dtype=np.uint8)
    b = rng.integers(0, 256, size=(batch, c, h, w), dtype=np.uint8)
    for qa, qb, qo in (((0.02, 40), (0.03, 50), (0.03, 50)), ((0.05, 0), (0.05, 255), (0.011, 128)), ((0.001, 7), (0.5, 200), (0.02, 3))):
        ref = np.stack([O.shortcut(a[i], b[i], qa, qb, qo) for i in range(batch)])
        assert np.array_equal(darknet.forward_shortcut_layer_quant_gpu(a, b, qa, qb, qo), ref)
        geoms = ((1, w + 1, h + 1), (0, w, h), (2, w + 7, h + 3))
        assert np.array_equal(darknet.forward_shortcut_layer_quant_gpu(a, b, qa, qb, qo, geoms=geoms), ref)
    assert ref.min() == 0 or ref.max() == 255       # the last case saturates
    with pytest.raises(Exception, match="scale ratio"):
        darknet.shortcut_multiplier(100.0, 0.01)


@pytest.mark.parametrize("case", [(128, 13, 13, 256, 3, 3), (64, 20, 20, 128, 3, 2), (256, 9, 9, 512, 3, 2), (256, 13, 13, 128, 1, 2), (512, 7, 7, 1024, 3, 1)],
                         ids=lambda c: "c%d_%dx%d_n%d_k%d" % c[:5])
def test_conv_with_fused_shortcut_vs_oracle(built, case):
    """a flat convolution (flat2 for c < 256, the CTA-pair flat2x from c = 256 on) with the FOLLOWING quantized shortcut fused into
    its epilogue == oracle conv -> oracle shortcut, incl. halo bytes and saturation at both ends"""
    import ctypes as C
    from yolo_quantization_b200 import _lib
    from yolo_quantization_b200._lib import ActGeom, check
    c, h, w, n, k, batch = case
    lib = _lib.load()
    rng = np.random.default_rng(zlib.crc32(repr(case).encode()) + 21)
    layer, wq, zp_w, p = _rand_layer(rng, c, n, k, 1, "leaky", 40, zp_out=33, h=h, w=w, quant_stop=0)
    assert lib.yq_conv_flat_shortcut_supported(layer.handle)
    x = rng.integers(0, 256, size=(batch, c, h, w), dtype=np.uint8)
    b = rng.integers(0, 256, size=(batch, n, h, w), dtype=np.uint8)
    qa, qb, qo = (0.05, 33), (0.03, 50), (0.021, 60)
    ka, kb = darknet.shortcut_multiplier(qa[0], qo[0]), darknet.shortcut_multiplier(qb[0], qo[0])
    g = ActGeom()
    check(lib.yq_act_geom_flat(h, w, C.byref(g)))

    def stage(t, ch, fill):
        d = darknet.DeviceBuffer(lib.yq_act_geom_bytes(C.byref(g), batch, ch), zero=False)
        check(lib.yq_cuda_memset(d.ptr, fill, d.nbytes, None))
        src = darknet.DeviceBuffer.from_numpy(np.ascontiguousarray(t))
        check(lib.yq_nchw_to_nhwc_u8_geom(src.ptr, d.ptr, batch, ch, h, w, C.byref(g), None))
        check(lib.yq_stream_synchronize(None))
        src.free()
        return d
    din, dres = stage(x, c, 40), stage(b, n, 0x77)
    dout = darknet.DeviceBuffer(lib.yq_act_geom_bytes(C.byref(g), batch, n), zero=False)
    check(lib.yq_cuda_memset(dout.ptr, 0xEE, dout.nbytes, None))
    check(lib.yq_forward_convolutional_layer_quant_flat_shortcut_gpu(layer.handle, din.ptr, dres.ptr, dout.ptr, 0x5C, qb[1], ka, kb, qo[1], batch, None),
          "yq_forward_convolutional_layer_quant_flat_shortcut_gpu")
    tmp = darknet.DeviceBuffer(batch * n * h * w)
    check(lib.yq_nhwc_to_nchw_u8_geom(dout.ptr, tmp.ptr, batch, n, h, w, C.byref(g), None))
    check(lib.yq_stream_synchronize(None))
    got = tmp.pull((batch, n, h, w), np.uint8)
    raw = dout.pull((dout.nbytes // n, n), np.uint8)
    rows = raw[: batch * g.rows_h * g.pitch_w].reshape(batch, g.rows_h, g.pitch_w, n).copy()
    rows[:, 1:1 + h, 1:1 + w, :] = 0x5C
    assert (rows == 0x5C).all() and (raw[batch * g.rows_h * g.pitch_w:] == 0x5C).all(), "halo of the fused launch's output"
    for i in range(batch):
        acc = O.conv_acc(x[i], wq, zp_w, 1, k // 2, 40)
        a = O.requant(acc, p["biases_int32"], p["M_value"], p["M0_right_shift_value"], synth.ACT_CODES["leaky"], 33)
        assert np.array_equal(got[i], O.shortcut(a, b[i], qa, qb, qo)), f"image {i}"
    assert got.min() == 0 and got.max() == 255
    for d in (din, dres, dout, tmp):
        d.free()
    layer.free()


@pytest.mark.parametrize("full_m0", [True, False, 2], ids=["M0_31_bits", "M0_from_float", "M0_from_float_shift_3"])
@pytest.mark.parametrize("act", ["leaky", "relu6"])
def test_requant_beyond_2_22(built, full_m0, act):
    """|acc + bias| far beyond 2^22.  The integer epilogue equals the reference's double arithmetic (convolutional_layer.c:732-733)
    as long as |x| * M0 fits 53 bits: with the reference's own host prep M0 = round(M * 2^31) of a FLOAT M has >= 7 trailing zeros
    and the integer form serves |x| < 2^29; a binding that hands in a full 31-bit (odd) M0 must get the FP64 re-do from 2^22 on,
    where the double product really rounds.  Both against the oracle's literal double arithmetic, flat2x + flat2 + per-tap."""
    c, h, w, n, k = 512, 6, 6, 128, 3
    rng = np.random.default_rng(77 + full_m0)
    wq = rng.integers(200, 256, size=(n, c * k * k), dtype=np.uint8)          # far above the zero points: huge accumulators
    zp_w = rng.integers(0, 40, size=n, dtype=np.uint8)
    x = rng.integers(180, 256, size=(2, c, h, w), dtype=np.uint8)
    s_w = (rng.random(n).astype(np.float32) * 1e-5 + 1e-6).astype(np.float32)
    spec = synth.LayerSpec("conv", n, k, 1, 1, 0, act)
    sl = synth.SynthLayer("conv", c, h, w, n, 0, 0, spec, s_out=0.05, biases=np.zeros(n, np.float32), s_w=s_w, zp_w=zp_w, w_u8=wq.reshape(n, c, k, k))
    p = O.prepare_conv(sl, 0.02, 3)
    if full_m0 == 2:
        # a LARGE multiplier (M in [2^-4, 2^-3)): h = trunc(|x| M) runs into the millions, far beyond the 81 914 up to which the LEAKY
        # epilogue's divide-by-ten multiply-add is exact -- make_epi() must have lowered xlim so that these chunks take the FP64 re-do
        # (and the uint8 store wraps, as the reference's does)
        p["M0_right_shift_value"] = np.full(n, 2.0 ** -3)
        p["M_value"] = (rng.random(n).astype(np.float32) * 0.5 + 0.5).astype(np.float64)
        p["M_value"] = np.round(p["M_value"] * 2.0 ** 31) * 2.0 ** -31
    elif full_m0:
        m0 = (rng.integers(1 << 30, 1 << 31, size=n, dtype=np.int64) | 1)       # odd: all 31 bits significant
        p["M_value"] = m0.astype(np.float64) * 2.0 ** -31
        p["M0_right_shift_value"] = np.full(n, 2.0 ** -24)
    for stride in (1, 2):
        layer = darknet.ConvolutionalLayerQuant(h, w, c, n, k, stride, 1, synth.ACT_CODES[act], wq, zp_w, p["biases_int32"], p["M_value"],
                                                p["M0_right_shift_value"], 3, 33, 0.05)
        got = layer.forward_flat(x, halo_fill=3) if stride == 1 else layer.forward(x)
        layer.free()
        for b in range(2):
            acc = O.conv_acc(x[b], wq.reshape(n, c, k, k), zp_w, stride, 1, 3)
            assert np.abs(acc + p["biases_int32"][:, None, None]).max() > (1 << 24)
            assert np.array_equal(got["acc"][b], acc)
            u8 = O.requant(acc, p["biases_int32"], p["M_value"], p["M0_right_shift_value"], synth.ACT_CODES[act], 33)
            assert np.array_equal(got["u8"][b], u8), f"stride {stride} image {b}"


def _yolov3_files(tmp_path, size, batch, seed=2):
    layers = synth.yolov3_quant()
    cfg, wts = str(tmp_path / "v3.cfg"), str(tmp_path / "v3.weights")
    synth.write_cfg(cfg, layers, batch=batch, width=size, height=size)
    info = synth.write_weights(wts, layers, width=size, height=size, seed=seed, identity_bn=False)
    return cfg, wts, info


def _walk(net, info, imgs, debug):
    heads = net.split_heads(net.predict_u8(imgs))
    for b in range(imgs.shape[0]):
        ref = O.forward_network(info, imgs[b])
        hi = 0
        for i, (sl, r) in enumerate(zip(info, ref)):
            li = net.layer_info(i)
            if sl.kind == "yolo":
                assert np.allclose(heads[hi][b], r["f32"], atol=YOLO_ATOL, rtol=0), f"yolo layer {i}"
                hi += 1
                continue
            if sl.kind == "conv" and debug:
                assert np.array_equal(net.pull_layer(i, "acc")[b], r["acc"]), f"layer {i} int32 mismatch (image {b})"
            if not debug and ((li.fused and sl.kind != "shortcut") or (sl.kind == "conv" and sl.spec.quant_stop)):
                continue                                  # not materialised in the production plan (a fused shortcut's tensor is: its conv wrote it)
            assert np.array_equal(net.pull_layer(i, "u8")[b], r["u8"]), f"layer {i} ({sl.kind}) uint8 mismatch (image {b})"


def test_yolov3_network_96_vs_oracle(built, tmp_path):
    """BASELINE configs[4] as a NETWORK: all 107 layers at 96x96, batch 2 -- debug plan (every tensor, every int32
    accumulator) and production plan (fused heads, flat strips, side streams) == the oracle's layer walk."""
    cfg, wts, info = _yolov3_files(tmp_path, 96, 2)
    assert sum(s.kind == "conv" for s in info) == 75 and sum(s.kind == "shortcut" for s in info) == 23
    net = darknet.load_network(cfg, wts, batch=2)
    assert net.n == 107
    imgs = np.stack([synth.synthetic_image(s, 3, 96, 96) for s in (21, 22)])
    net.set_debug(True)
    _walk(net, info, imgs, debug=True)
    dbg = net.predict_u8(imgs).copy()
    net.set_debug(False)
    _walk(net, info, imgs, debug=False)
    assert np.array_equal(net.predict_u8(imgs), dbg)
    kinds = [net.layer_info(i).kernel for i in range(net.n) if net.layer_info(i).type == 0]
    assert 0 not in kinds, f"a convolution of the production plan runs the SIMT flavour: {kinds}"
    net.use_graph(True)
    for _ in range(2):
        assert np.array_equal(net.predict_u8(imgs), dbg)
    net.free()


def test_yolov3_network_416_heads_vs_oracle(built, tmp_path):
    """the real size: one 416x416 image in a batch of 3 (production plan, CUDA graph) -- the three yolo heads and the last
    shortcut of every stage == oracle; every image equals its batch-1 run bit for bit."""
    cfg, wts, info = _yolov3_files(tmp_path, 416, 3, seed=4)
    imgs = np.stack([synth.synthetic_image(s) for s in (51, 52, 53)])
    net = darknet.load_network(cfg, wts, batch=3)
    net.use_graph(True)
    flat = net.predict_u8(imgs).copy()
    heads = net.split_heads(flat)
    ref = O.forward_network(info, imgs[1])
    for h, i in zip(heads, (82, 94, 106)):
        assert np.allclose(h[1], ref[i]["f32"], atol=YOLO_ATOL, rtol=0), f"yolo layer {i}"
    for i in (4, 11, 36, 61, 74):
        assert np.array_equal(net.pull_layer(i, "u8")[1], ref[i]["u8"]), f"shortcut {i}"
    net.free()
    one = darknet.load_network(cfg, wts, batch=1)
    for b in (0, 2):
        h1 = one.split_heads(one.predict_u8(imgs[b:b + 1]))
        for h3, h in zip(heads, h1):
            assert np.array_equal(h3[b], h[0])
    one.free()


def test_shortcut_cfg_errors(built, tmp_path):
    layers = [synth.LayerSpec("conv", 16, 3, activation="leaky"), synth.LayerSpec("conv", 32, 3, activation="leaky"),
              synth.LayerSpec("shortcut", layers=(-2,))]
    cfg, wts = str(tmp_path / "e.cfg"), str(tmp_path / "e.weights")
    synth.write_cfg(cfg, layers, width=16, height=16)
    with open(wts, "wb") as f:
        f.write(b"\0" * 64)
    with pytest.raises(Exception, match="differ in shape"):
        darknet.load_network(cfg, wts)
