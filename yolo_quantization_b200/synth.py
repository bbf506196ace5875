"""Synthetic fixtures in the reference's own on-disk formats.

The reference ships no weights (README.md:48-51) and no tests, so every fixture is generated:

* ``write_cfg``      -- a darknet ``.cfg`` (the INI dialect parsed by ``src/parser.c:682-850``) for the
                        quantized yolov3-tiny of ``cfg/yolov3_tiny_quant_channelwise.cfg`` (same 24 layers,
                        regenerated from the table below -- not a copy of the file), for single-conv nets
                        (the per-layer oracle trick) and for leaky / stride-2 variants.
* ``write_weights``  -- a ``.weights`` byte stream exactly as ``load_weights_upto`` consumes it under
                        ``-DQUANTIZATION`` (``src/parser.c:1201-1305``; per-layer fields ``:1124-1199``).
* ``synthetic_image``-- seeded uint8 CHW image that contains at least one 0 and one 255 so the reference's
                        dynamic layer-0 quantiser (``src/blas.c:108-168``) reproduces it with s=1/255, zp=0.

Everything is seeded (numpy ``default_rng``) so the GPU box regenerates identical bytes.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

# activation codes follow the reference's ACTIVATION enum order (include/darknet.h:87-89):
# LOGISTIC, RELU, RELIE, LINEAR, RAMP, TANH, PLSE, LEAKY6, RELU6, LEAKY, ELU, ...
ACT_CODES = {"logistic": 0, "relu": 1, "linear": 3, "relu6": 8, "leaky": 9}


@dataclass
class LayerSpec:
    kind: str                       # conv | maxpool | route | upsample | yolo | shortcut
    filters: int = 0
    size: int = 1
    stride: int = 1
    pad: int = 1                    # cfg "pad=1" => padding = size/2 (parser.c:175-178)
    bn: int = 1
    activation: str = "relu6"
    quant_stop: int = 0
    layers: Tuple[int, ...] = ()    # route; shortcut: (from,)
    mask: Tuple[int, ...] = ()      # yolo
    classes: Optional[int] = None   # yolo: None => module CLASSES / ANCHORS (the shipped 5-class, 6-anchor cfg)
    anchors: Optional[str] = None
    num: Optional[int] = None
    # activation quantisation of this layer's output (s_out, zp_out); None => defaults by activation
    act_scale: Optional[float] = None
    act_zp: Optional[int] = None


ANCHORS = "25,39, 29,88, 405,102, 407,109,408,113,420,129"
CLASSES = 5


def yolov3_tiny_quant(activation: str = "relu6", head_filters: int = 3 * (5 + CLASSES)) -> List[LayerSpec]:
    """The 24-layer net of cfg/yolov3_tiny_quant_channelwise.cfg:27-230."""
    a = activation
    C, M, R, U, Y = "conv", "maxpool", "route", "upsample", "yolo"
    return [
        LayerSpec(C, 16, 3, activation=a), LayerSpec(M, size=2, stride=2),
        LayerSpec(C, 32, 3, activation=a), LayerSpec(M, size=2, stride=2),
        LayerSpec(C, 64, 3, activation=a), LayerSpec(M, size=2, stride=2),
        LayerSpec(C, 128, 3, activation=a), LayerSpec(M, size=2, stride=2),
        LayerSpec(C, 256, 3, activation=a), LayerSpec(M, size=2, stride=2),
        LayerSpec(C, 512, 3, activation=a), LayerSpec(M, size=2, stride=1),
        LayerSpec(C, 1024, 3, activation=a),
        LayerSpec(C, 256, 1, activation=a),
        LayerSpec(C, 512, 3, activation=a),
        LayerSpec(C, head_filters, 1, bn=0, activation="linear", quant_stop=1),
        LayerSpec(Y, mask=(3, 4, 5)),
        LayerSpec(R, layers=(-4,)),
        LayerSpec(C, 128, 1, activation=a),
        LayerSpec(U, stride=2),
        LayerSpec(R, layers=(-1, 8)),
        LayerSpec(C, 256, 3, activation=a),
        LayerSpec(C, head_filters, 1, bn=0, activation="linear", quant_stop=1),
        LayerSpec(Y, mask=(0, 1, 2)),
    ]


YOLOV3_ANCHORS = "10,13, 16,30, 33,23, 30,61, 62,45, 59,119, 116,90, 156,198, 373,326"


def yolov3_quant(classes: int = 80, activation: str = "leaky") -> List[LayerSpec]:
    """Full yolov3 (BASELINE configs[4]): 107 layers = 75 conv (3x3/1, 3x3/2 down-samplers, 1x1; BN + leaky except the three
    linear heads), 23 shortcut, 4 route, 2 upsample, 3 yolo.  The reference ships no such cfg and has no quantized shortcut
    (SURVEY 0.10, Appendix F); the layer table is the public darknet yolov3 topology re-derived here, every layer quantized=1,
    the shortcut being this repo's integer extension (spec: include/yq_b200.h, yq_forward_shortcut_layer_quant_gpu)."""
    a = activation
    hf = 3 * (5 + classes)
    L: List[LayerSpec] = []
    conv = lambda f, k, s=1: L.append(LayerSpec("conv", f, k, s, activation=a))
    def block(f: int, reps: int) -> None:
        conv(f, 3, 2)                                  # down-sampler
        for _ in range(reps):
            conv(f // 2, 1)
            conv(f, 3)
            L.append(LayerSpec("shortcut", layers=(-3,)))
    def head(mask) -> None:
        L.append(LayerSpec("conv", hf, 1, bn=0, activation="linear", quant_stop=1))
        L.append(LayerSpec("yolo", mask=mask, classes=classes, anchors=YOLOV3_ANCHORS, num=9))
    conv(32, 3)
    for f, reps in ((64, 1), (128, 2), (256, 8), (512, 8), (1024, 4)):
        block(f, reps)
    for f, mask, skip in ((512, (6, 7, 8), 61), (256, (3, 4, 5), 36), (128, (0, 1, 2), None)):
        for _ in range(3):
            conv(f, 1)
            conv(f * 2, 3)
        head(mask)
        if skip is not None:
            L.append(LayerSpec("route", layers=(-4,)))
            conv(f // 2, 1)
            L.append(LayerSpec("upsample", stride=2))
            L.append(LayerSpec("route", layers=(-1, skip)))
    assert len(L) == 107 and sum(l.kind == "conv" for l in L) == 75 and sum(l.kind == "shortcut" for l in L) == 23
    return L


def single_conv(filters: int, size: int, stride: int = 1, activation: str = "relu6", bn: int = 1,
                quant_stop: int = 0, act_scale: Optional[float] = None, act_zp: Optional[int] = None
                ) -> List[LayerSpec]:
    return [LayerSpec("conv", filters, size, stride, 1, bn, activation, quant_stop,
                      act_scale=act_scale, act_zp=act_zp)]


def write_cfg(path: str, layers: Sequence[LayerSpec], batch: int = 1, width: int = 416, height: int = 416,
              channels: int = 3) -> None:
    out = ["[net]", f"batch={batch}", "subdivisions=1", f"width={width}", f"height={height}",
           f"channels={channels}", "momentum=0.9", "decay=0.0005", "learning_rate=0.001",
           "max_batches=10", "policy=steps", "steps=5", "scales=.1", "start_quantization_step=10000", ""]
    for l in layers:
        if l.kind == "conv":
            out += ["[convolutional]"]
            if l.bn:
                out += ["batch_normalize=1"]
            out += [f"filters={l.filters}", f"size={l.size}", f"stride={l.stride}", f"pad={l.pad}",
                    f"activation={l.activation}", "quantized=1", f"quant_stop={l.quant_stop}", ""]
        elif l.kind == "maxpool":
            out += ["[maxpool]", f"size={l.size}", f"stride={l.stride}", "quantized=1", "quant_stop=0", ""]
        elif l.kind == "route":
            out += ["[route]", "layers = " + ", ".join(str(i) for i in l.layers), "quantized=1", "quant_stop=0", ""]
        elif l.kind == "upsample":
            out += ["[upsample]", f"stride={l.stride}", "quantized=1", "quant_stop=0", ""]
        elif l.kind == "yolo":
            out += ["[yolo]", "mask = " + ",".join(str(i) for i in l.mask), f"anchors = {l.anchors or ANCHORS}",
                    f"classes={CLASSES if l.classes is None else l.classes}", f"num={l.num or 6}", "jitter=.3",
                    "ignore_thresh = .7", "truth_thresh = 1", "random=1", ""]
        elif l.kind == "shortcut":
            # the quantized shortcut is this repo's extension (the reference's shortcut is float only)
            out += ["[shortcut]", f"from={l.layers[0]}", "activation=linear", "quantized=1", "quant_stop=0", ""]
        else:
            raise ValueError(l.kind)
    with open(path, "w") as f:
        f.write("\n".join(out))


def default_act_quant(l: LayerSpec, relu6_scale: float) -> Tuple[float, int]:
    if l.act_scale is not None:
        return float(np.float32(l.act_scale)), int(l.act_zp or 0)
    if l.activation == "relu6":
        return float(np.float32(relu6_scale)), 0
    if l.activation == "leaky":
        return float(np.float32(0.02)), 40
    return float(np.float32(0.08)), 128          # linear heads


@dataclass
class SynthLayer:
    """What the generator wrote for one layer (host-side truth for tests)."""
    kind: str
    c: int = 0
    h: int = 0
    w: int = 0
    out_c: int = 0
    out_h: int = 0
    out_w: int = 0
    spec: Optional[LayerSpec] = None
    s_in: float = 0.0
    zp_in: int = 0
    s_out: float = 0.0
    zp_out: int = 0
    biases: Optional[np.ndarray] = None        # f32 [n] as stored (before BN folding)
    bn_scales: Optional[np.ndarray] = None
    bn_mean: Optional[np.ndarray] = None
    bn_var: Optional[np.ndarray] = None
    s_w: Optional[np.ndarray] = None           # f32 [n]
    zp_w: Optional[np.ndarray] = None          # u8 [n]
    w_u8: Optional[np.ndarray] = None          # u8 [n, c, k, k]  (OIHW, parser.c:1143-1145)
    inputs: Tuple[int, ...] = field(default_factory=tuple)


def write_weights(path: Optional[str], layers: Sequence[LayerSpec], width: int = 416, height: int = 416,
                  channels: int = 3, seed: int = 0, relu6_scale: float = 12.0 / 255.0,
                  input_quant: Tuple[float, int] = (1.0 / 255.0, 0), identity_bn: bool = True
                  ) -> List[SynthLayer]:
    """Write the QUANTIZATION-layout ``.weights`` stream; returns the per-layer truth.

    Field order per layer: parser.c:1124-1159 (conv), :1161-1172 (maxpool), :1174-1183 (route),
    :1185-1199 (upsample).  Header: parser.c:1213-1225 (major,minor,revision int32 + size_t seen).
    """
    rng = np.random.default_rng(seed)
    chunks: List[bytes] = [struct.pack("<iii", 0, 2, 0), struct.pack("<Q", 0)]
    res: List[SynthLayer] = []
    c, h, w = channels, height, width
    prev = (float(np.float32(input_quant[0])), int(input_quant[1]))
    for i, l in enumerate(layers):
        if l.kind == "conv":
            n, k = l.filters, l.size
            pad = k // 2 if l.pad else 0
            oh = (h + 2 * pad - k) // l.stride + 1
            ow = (w + 2 * pad - k) // l.stride + 1
            K = c * k * k
            wf = (rng.standard_normal((n, K)) * np.sqrt(2.0 / K)).astype(np.float32)
            b = (rng.standard_normal(n) * 0.1).astype(np.float32)
            sl = SynthLayer("conv", c, h, w, n, oh, ow, l, biases=b)
            chunks.append(b.tobytes())
            if l.bn:
                if identity_bn:
                    sc, mean, var = np.ones(n, "f4"), np.zeros(n, "f4"), np.ones(n, "f4")
                else:
                    sc = (1.0 + 0.2 * rng.standard_normal(n)).astype("f4")
                    mean = (0.1 * rng.standard_normal(n)).astype("f4")
                    var = (0.5 + rng.random(n)).astype("f4")
                sl.bn_scales, sl.bn_mean, sl.bn_var = sc, mean, var
                chunks += [sc.tobytes(), mean.tobytes(), var.tobytes()]
            mn = np.minimum(wf.min(1), 0).astype("f4")
            mx = np.maximum(wf.max(1), 0).astype("f4")
            s_w = ((mx - mn) / np.float32(255.0)).astype("f4")
            zp_w = np.clip(np.round(-mn / s_w), 0, 255).astype("u1")
            q = np.clip(np.round(wf / s_w[:, None]) + zp_w[:, None], 0, 255).astype("u1")
            s_out, zp_out = default_act_quant(l, relu6_scale)
            chunks.append(struct.pack("<f", prev[0]) + struct.pack("<B", prev[1]))
            chunks.append(struct.pack("<f", s_out) + struct.pack("<B", zp_out))
            chunks += [s_w.tobytes(), zp_w.tobytes(), q.tobytes(), wf.tobytes()]
            sl.s_in, sl.zp_in, sl.s_out, sl.zp_out = prev[0], prev[1], s_out, zp_out
            sl.s_w, sl.zp_w, sl.w_u8 = s_w, zp_w, q.reshape(n, c, k, k)
            prev = (s_out, zp_out)
            c, h, w = n, oh, ow
        elif l.kind == "maxpool":
            padding = l.size - 1                                         # parser.c:415
            oh = (h + padding - l.size) // l.stride + 1                  # maxpool_layer.c:31-32
            ow = (w + padding - l.size) // l.stride + 1
            chunks.append(struct.pack("<f", prev[0]) + struct.pack("<B", prev[1]))
            sl = SynthLayer("maxpool", c, h, w, c, oh, ow, l, s_out=prev[0], zp_out=prev[1])
            h, w = oh, ow
        elif l.kind == "upsample":
            chunks.append(struct.pack("<f", prev[0]) + struct.pack("<B", prev[1]))
            sl = SynthLayer("upsample", c, h, w, c, h * l.stride, w * l.stride, l, s_out=prev[0], zp_out=prev[1])
            h, w = h * l.stride, w * l.stride
        elif l.kind == "route":
            idx = tuple(j if j >= 0 else i + j for j in l.layers)
            first = res[idx[0]]
            oc = sum(res[j].out_c for j in idx)
            if len(idx) > 1:
                # route with n>1 carries its own (s, zp) on disk; we give it the first input's
                prev = (first.s_out, first.zp_out)
                chunks.append(struct.pack("<f", prev[0]) + struct.pack("<B", prev[1]))
            else:
                prev = (first.s_out, first.zp_out)                       # inherited, nothing on disk
            sl = SynthLayer("route", oc, first.out_h, first.out_w, oc, first.out_h, first.out_w, l,
                            s_out=prev[0], zp_out=prev[1], inputs=idx)
            c, h, w = oc, first.out_h, first.out_w
        elif l.kind == "shortcut":
            j = l.layers[0] if l.layers[0] >= 0 else i + l.layers[0]
            assert (res[j].out_c, res[j].out_h, res[j].out_w) == (c, h, w), "shortcut inputs must agree in shape"
            # extension record: the layer's own output (scale, zero point), 5 bytes like a maxpool's (parser.c:1161-1172)
            s_out = float(np.float32(l.act_scale if l.act_scale is not None else 0.03))
            zp_out = int(l.act_zp if l.act_zp is not None else 50)
            chunks.append(struct.pack("<f", s_out) + struct.pack("<B", zp_out))
            sl = SynthLayer("shortcut", c, h, w, c, h, w, l, s_in=prev[0], zp_in=prev[1], s_out=s_out, zp_out=zp_out, inputs=(j,))
            prev = (s_out, zp_out)
        elif l.kind == "yolo":
            sl = SynthLayer("yolo", c, h, w, c, h, w, l, s_out=prev[0], zp_out=prev[1])
        else:
            raise ValueError(l.kind)
        res.append(sl)
    if path is not None:
        with open(path, "wb") as f:
            for ch in chunks:
                f.write(ch)
    return res


def synthetic_image(seed: int = 1, channels: int = 3, height: int = 416, width: int = 416) -> np.ndarray:
    """uint8 CHW image, uniform in [0,255], guaranteed to contain 0 and 255 (see module docstring)."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(channels, height, width), dtype=np.uint8)
    img.flat[0], img.flat[1] = 0, 255
    return img


def image_to_float(img_u8: np.ndarray, scale: float = 1.0 / 255.0, zp: int = 0) -> np.ndarray:
    """float image x=(u8-zp)*s that the reference's layer-0 quantiser maps back to ``img_u8``."""
    return ((img_u8.astype(np.float32) - np.float32(zp)) * np.float32(scale)).astype(np.float32)
