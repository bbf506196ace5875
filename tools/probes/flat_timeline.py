"""Per-CTA timeline of the flat conv kernel on a layer-12-shaped convolution (512 -> 1024, 13x13, batch 128).
Needs libyq_b200.so built with -DYQ_TIMELINE (make NVFLAGS_EXTRA=-DYQ_TIMELINE)."""
import ctypes as C
import os
import sys
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from yolo_quantization_b200 import _lib, darknet, synth  # noqa: E402

c, h, w, n, k, batch = (int(x) for x in (sys.argv[1:7] + ["512", "13", "13", "1024", "3", "128"][len(sys.argv) - 1:]))
rng = np.random.default_rng(0)
wq = rng.integers(0, 256, size=(n, c * k * k), dtype=np.uint8)
zp_w = rng.integers(0, 256, size=n, dtype=np.uint8)
head = n % 128 != 0
layer = darknet.ConvolutionalLayerQuant(h, w, c, n, k, 1, k // 2, synth.ACT_CODES["linear" if head else "relu6"], wq, zp_w, np.zeros(n, np.int32),
                                        np.full(n, 0.25 * 2.0 ** -8), np.ones(n), 0, 128 if head else 0, 0.05, quant_stop_flag=1 if head else 0)
x = rng.integers(0, 256, size=(batch, c, h, w), dtype=np.uint8)
for _ in range(3):
    layer.forward_flat(x, want_acc=False)
lib = _lib.load()
cudart = C.CDLL("libcudart.so.12") if False else None
# read the device symbol through the library's own export
buf = np.zeros(8 * 16384, np.uint64)
get = lib.yq_debug_flat_timeline
get.restype = C.c_int
get.argtypes = [C.c_void_p, C.c_size_t]
assert get(buf.ctypes.data, buf.nbytes) == 0
t = buf.reshape(-1, 8).astype(np.int64)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
names = ["start", "setup", "first_operands", "last_mma_issued", "acc_complete", "epi_math", "store_drained"]
print(f"{len(t)} CTAs; kernel span {(t[:, 6].max() - t0) / 1e3:.1f} us")
for a, b in ((0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (0, 6)):
    d = (t[:, b] - t[:, a]) / 1e3
    print(f"{names[a]:>16s} -> {names[b]:<16s} median {np.median(d):7.2f} us   p10 {np.percentile(d, 10):7.2f}   p90 {np.percentile(d, 90):7.2f}")
starts = np.sort(t[:, 0] - t0) / 1e3
print("CTA start times (us), every 148th:", starts[::148].round(1).tolist())
layer.free()
