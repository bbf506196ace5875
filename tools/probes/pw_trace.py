"""Per-CTA event timeline of the pointwise flavour (yq_conv_tc_pw.cu) on one 1x1 layer: c h w n batch [classes].
    YQ_PW_TRACE=1 python tools/probes/pw_trace.py 256 26 26 30 128 5      (head 22 of yolov3-tiny)"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from yolo_quantization_b200 import _lib, darknet, synth  # noqa: E402

os.environ["YQ_PW_TRACE"] = "1"
args = [int(x) for x in sys.argv[1:]]
c, h, w, n, batch = args[:5]
classes = args[5] if len(args) > 5 else None
rng = np.random.default_rng(0)
wq = rng.integers(0, 256, size=(n, c), dtype=np.uint8)
zp_w = rng.integers(0, 256, size=n, dtype=np.uint8)
head = classes is not None
layer = darknet.ConvolutionalLayerQuant(h, w, c, n, 1, 1, 0, synth.ACT_CODES["linear" if head else "leaky"], wq, zp_w, np.zeros(n, np.int32),
                                        np.full(n, 0.25 * 2.0 ** -8), np.ones(n), 0, 128 if head else 0, 0.05, quant_stop_flag=1 if head else 0)
x = rng.integers(0, 256, size=(batch, c, h, w), dtype=np.uint8)
for _ in range(3):
    layer.forward_flat(x, want_acc=False, yolo_classes=classes)
lib = _lib.load()
buf = np.zeros(256 * 64, np.uint64)
get = lib.yq_debug_pw_trace
get.restype = C.c_int
get.argtypes = [C.c_void_p, C.c_size_t]
assert get(buf.ctypes.data, buf.nbytes) == 0
t = buf.reshape(-1, 64).astype(np.int64)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
r = lambda a: (np.median(a) / 1e3).round(2)
print(f"{len(t)} CTAs; kernel span {(t[:, 41].max() - t0) / 1e3:.1f} us; CTA start spread {(t[:, 0].max() - t0) / 1e3:.1f} us")
print("medians over CTAs, us since the CTA's own start:")
print("  setup done", r(t[:, 1] - t[:, 0]), " previous grid complete", r(t[:, 2] - t[:, 0]), " filter bank landed", r(t[:, 3] - t[:, 0]))
for k in range(8):
    m = t[:, 4 + 4 * k] > 0
    if not m.any():
        break
    s = t[m]
    e = s[:, 4 + 4 * k + 2] > 0
    print(f"  tile {k} ({m.sum()} CTAs): loads issued {r(s[:, 4 + 4 * k + 3] - s[:, 0])}  operands landed + MMAs issued {r(s[:, 4 + 4 * k] - s[:, 0])}"
          f"  accumulator seen by the epilogue {r(s[e][:, 4 + 4 * k + 1] - s[e][:, 0])}  epilogue done {r(s[e][:, 4 + 4 * k + 2] - s[e][:, 0])}")
print("  stores drained (warp 0)", r(t[:, 40] - t[:, 0]), " CTA end", r(t[:, 41] - t[:, 0]), " max CTA end", ((t[:, 41] - t[:, 0]).max() / 1e3).round(2))
layer.free()
