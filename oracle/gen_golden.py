"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/ from the COMPILED REFERENCE (oracle/_ref, built by
`make -C oracle ref` from the unmodified sources under /root/reference).  Run in the build container:

    python -m oracle.gen_golden

The reference has no tests or golden vectors of its own; these fixtures pin the oracle
(and through it the CUDA path) to outputs of the reference itself:

  tests/golden/tiny416.json      SHA-256 of every per-layer dump (int32 accumulators, uint8 outputs, yolo floats,
                                 prepared per-channel params) of the 24-layer yolov3-tiny at 416x416, seeded
                                 synthetic weights (seed 0) and image (seed 1).
  tests/golden/tiny96_leaky.json same net with leaky activations (zp_out = 40 -> zp_in != 0 padding) at 96x96.
  tests/golden/layer_*.npz       small single-conv cases (the per-layer oracle trick):
                                 inputs, weights, prepared params and the reference's int32 / uint8 / f32 outputs.
"""
from __future__ import annotations

import hashlib
import json
import os
import tempfile
import zlib

import numpy as np

from oracle import yq_oracle as O
from yolo_quantization_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_net(layers, size, seed_w, seed_img, name, **wkw):
    with tempfile.TemporaryDirectory() as d:
        cfg, wts, img = (os.path.join(d, x) for x in ("n.cfg", "n.weights", "img.f32"))
        synth.write_cfg(cfg, layers, batch=1, width=size, height=size)
        synth.write_weights(wts, layers, width=size, height=size, seed=seed_w, **wkw)
        im = synth.synthetic_image(seed_img, 3, size, size)
        synth.image_to_float(im).tofile(img)
        O.run_reference("net", cfg, wts, img, os.path.join(d, "dump"))
        dump = O.read_dump(os.path.join(d, "dump"))
    out = {"generator": "oracle/gen_golden.py", "net": name, "size": size, "seed_weights": seed_w, "seed_image": seed_img,
           "weights_kwargs": wkw, "input_sha256": sha(im), "layers": []}
    for dl in dump:
        e = {"index": dl["index"], "type": dl["type"]}
        for k in ("output_int32", "output_uint8", "output_f32", "M0", "M0_right_shift", "biases_int32"):
            if k in dl:
                e[k] = sha(dl[k])
        out["layers"].append(e)
    return out


# (name, c, h, w, filters, size, stride, activation, bn, quant_stop, s_in, zp_in, act_scale, act_zp, identity_bn)
LAYER_CASES = [
    ("relu6_3x3", 8, 12, 12, 16, 3, 1, "relu6", 1, 0, 0.05, 0, 0.05, 0, True),
    ("leaky_3x3_s2_zp37", 16, 13, 13, 32, 3, 2, "leaky", 1, 0, 0.02, 37, 0.02, 40, False),
    ("linear_1x1_head", 32, 7, 7, 30, 1, 1, "linear", 0, 1, 0.047, 0, 0.08, 128, True),
    ("relu_3x3_wrap", 4, 9, 11, 8, 3, 1, "relu", 1, 0, 0.05, 5, 0.004, 3, False),      # small s_out -> uint8 wrap
    ("leaky_1x1_c3", 3, 10, 10, 16, 1, 1, "leaky", 1, 0, 1.0 / 255.0, 0, 0.01, 60, True),
]


def run_layer(case):
    name, c, h, w, n, k, stride, act, bn, qs, s_in, zp_in, a_s, a_z, ident = case
    layers = synth.single_conv(n, k, stride, act, bn, qs, act_scale=a_s, act_zp=a_z)
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    x = rng.integers(0, 256, size=(c, h, w), dtype=np.uint8)
    x.flat[0], x.flat[1] = 0, 255           # so the reference's dynamic input quantiser reproduces (s_in, zp_in)
    with tempfile.TemporaryDirectory() as d:
        cfg, wts, img = (os.path.join(d, f) for f in ("l.cfg", "l.weights", "img.f32"))
        synth.write_cfg(cfg, layers, batch=1, width=w, height=h, channels=c)
        info = synth.write_weights(wts, layers, width=w, height=h, channels=c, seed=7, input_quant=(s_in, zp_in),
                                   identity_bn=ident)
        synth.image_to_float(x, s_in, zp_in).tofile(img)
        O.run_reference("layer", cfg, wts, img, os.path.join(d, "dump"))
        dl = O.read_dump(os.path.join(d, "dump"))[0]
        got_in = np.fromfile(os.path.join(d, "dump", "L00_input_uint8.bin"), dtype=np.uint8).reshape(c, h, w)
    assert np.array_equal(got_in, x), f"{name}: reference input quantiser did not reproduce the intended uint8 tensor"
    assert dl["zp_in"] == zp_in, (name, dl["zp_in"], zp_in)
    sl = info[0]
    pad = k // 2
    np.savez_compressed(
        os.path.join(GOLD, f"layer_{name}.npz"), x=x, w_u8=sl.w_u8, zp_w=sl.zp_w, s_w=sl.s_w, biases=sl.biases,
        bn_scales=sl.bn_scales if bn else np.zeros(0, "f4"), bn_mean=sl.bn_mean if bn else np.zeros(0, "f4"),
        bn_var=sl.bn_var if bn else np.zeros(0, "f4"),
        geom=np.array([c, h, w, n, k, stride, pad, synth.ACT_CODES[act], bn, qs, zp_in, a_z], np.int32),
        scales=np.array([dl["s_in"], dl["s_out"]], np.float32),
        M_value=dl["M_value"], M0_right_shift_value=dl["M0_right_shift_value"], biases_int32=dl["biases_int32"],
        M0=dl["M0"], M0_right_shift=dl["M0_right_shift"],
        ref_int32=dl["output_int32"], ref_uint8=dl["output_uint8"],
        ref_f32=dl.get("output_f32", np.zeros(0, "f4")))
    return name


def main():
    os.makedirs(GOLD, exist_ok=True)
    O.build()
    assert O.have_reference(), "oracle/_ref is not built (needs /root/reference)"
    with open(os.path.join(GOLD, "tiny416.json"), "w") as f:
        json.dump(run_net(synth.yolov3_tiny_quant(), 416, 0, 1, "yolov3_tiny_quant relu6"), f, indent=1)
    with open(os.path.join(GOLD, "tiny96_leaky.json"), "w") as f:
        json.dump(run_net(synth.yolov3_tiny_quant("leaky"), 96, 3, 5, "yolov3_tiny_quant leaky"), f, indent=1)
    for case in LAYER_CASES:
        print("layer case", run_layer(case))


if __name__ == "__main__":
    main()
