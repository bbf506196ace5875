"""Where the producer lane of the rows conv kernels spends its clocks, inside the yolov3-tiny forward at batch 128 (the bench's
data distribution).  Needs libyq_b200.so built with  make NVFLAGS_EXTRA=-DYQ_ROWS_TRACE  (non-planar producers only)."""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("YQ_NO_PLANAR", "1")      # the planar producer (two warps) is not instrumented
from yolo_quantization_b200 import _lib, darknet, synth  # noqa: E402

B = 128
layers = synth.yolov3_tiny_quant()
with tempfile.TemporaryDirectory() as d:
    cfg, wts = os.path.join(d, "t.cfg"), os.path.join(d, "t.weights")
    synth.write_cfg(cfg, layers, batch=B)
    synth.write_weights(wts, layers)
    net = darknet.load_network(cfg, wts, batch=B)
    x = np.random.default_rng(0).integers(0, 256, size=(B, 3, 416, 416), dtype=np.uint8)
    dev = darknet.DeviceBuffer.from_numpy(x)
    for _ in range(3):
        net.forward_device(dev.ptr)
        net.synchronize()
    lib = _lib.load()
    buf = np.zeros(5 * 8 * 1024, np.uint64)
    get = lib.yq_debug_rows_trace
    get.restype = C.c_int
    get.argtypes = [C.c_void_p, C.c_size_t]
    assert get(buf.ctypes.data, buf.nbytes) == 0
    buf2 = np.zeros(5 * 4 * 1024, np.uint64)
    get2 = lib.yq_debug_rows_trace2
    get2.restype = C.c_int
    get2.argtypes = [C.c_void_p, C.c_size_t]
    assert get2(buf2.ctypes.data, buf2.nbytes) == 0
    names = ["wait acc_empty (epilogue of the tile before the previous)", "wait full (TMA landed)", "MMA issue + commit", "wait acc_full of the previous tile",
             "refill: tile split + TMA issue"]
    for slot, what in enumerate(("layer 0 (c = 4)", "layer 2 (c = 16)", "layer 4 (c = 32)", "layer 6, channels 0-63", "layer 6, channels 64-127")):
        t = buf.reshape(5, 1024, 8)[slot].astype(np.float64)
        t = t[t[:, 7] > 0]
        if not len(t):
            continue
        per = t[:, :5] / t[:, 7:8]
        ghz = (t[:, 5] / np.maximum(t[:, 6], 1)).mean()
        print(f"{what}: {len(t)} CTAs, {t[:, 7].mean():.1f} tiles per CTA, {per.sum(1).mean():.0f} clocks per tile (producer lane); "
              f"loop {t[:, 6].mean() / 1e3:.1f} us at {ghz:.2f} GHz SM clock")
        t2 = buf2.reshape(5, 1024, 4)[slot].astype(np.float64)
        t2 = t2[t2[:, 2] > 0]
        if len(t2):
            print(f"    kernel entry -> loop start {t2[:, 0].mean() / 1e3:.1f} us (of which after the wait on the previous kernel: {t2[:, 1].mean() / 1e3:.1f} us); "
                  f"loop end .. CTA exit: {(t2[:, 3] - t[:len(t2), 6]).mean() / 1e3:.1f} us; first CTA loop start .. last CTA exit: "
                  f"{((t2[:, 2] + t2[:, 3]).max() - t2[:, 2].min()) / 1e3:.1f} us")
        for i, nm in enumerate(names):
            print(f"{nm:>60s}: mean {per[:, i].mean():7.0f} clk   p10 {np.percentile(per[:, i], 10):7.0f}   p90 {np.percentile(per[:, i], 90):7.0f}")
    net.free()
