"""Per-phase clocks of the rows conv kernel (thread 0 of every CTA), on a yolov3-tiny layer 0 / 2 / 4 shaped convolution.
Needs libyq_b200.so built with -DYQ_TIMELINE (make NVFLAGS_EXTRA=-DYQ_TIMELINE).
usage: rows_timeline.py c h w n batch"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from yolo_quantization_b200 import _lib, darknet, synth  # noqa: E402

c, h, w, n, batch = (int(x) for x in (sys.argv[1:6] + ["3", "416", "416", "16", "128"][len(sys.argv) - 1:]))
rng = np.random.default_rng(0)
wq = rng.integers(0, 256, size=(n, c * 9), dtype=np.uint8)
zp_w = rng.integers(0, 256, size=n, dtype=np.uint8)
layer = darknet.ConvolutionalLayerQuant(h, w, c, n, 3, 1, 1, synth.ACT_CODES["relu6"], wq, zp_w, np.zeros(n, np.int32), np.full(n, 0.25 * 2.0 ** -8),
                                        np.ones(n), 0, 0, 0.05)
x = rng.integers(0, 256, size=(batch, c, h, w), dtype=np.uint8)
for _ in range(2):
    layer.forward_rows_pooled(x)
lib = _lib.load()
buf = np.zeros(8 * 1024, np.uint64)
get = lib.yq_debug_rows_timeline
get.restype = C.c_int
get.argtypes = [C.c_void_p, C.c_size_t]
assert get(buf.ctypes.data, buf.nbytes) == 0
t = buf.reshape(-1, 8).astype(np.float64)
t = t[t[:, 7] > 0]
names = ["syncthreads (all warps done, copy fenced)", "MMA issue", "next copy issue + tile split", "wait for the MMAs", "TMEM loads + requant math",
         "cp.async wait + proxy fence", "global stores"]
per = t[:, :7] / t[:, 7:8]
print(f"{len(t)} CTAs, {t[:, 7].mean():.1f} tiles per CTA, {per.sum(1).mean():.0f} clocks per tile (thread 0)")
for i, nm in enumerate(names):
    print(f"{nm:>45s}: mean {per[:, i].mean():7.0f} clk   p10 {np.percentile(per[:, i], 10):7.0f}   p90 {np.percentile(per[:, i], 90):7.0f}")
layer.free()
