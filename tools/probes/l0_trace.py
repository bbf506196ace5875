"""Where the producer lane and an epilogue warp of the dense layer-0 kernel spend their clocks, inside the yolov3-tiny forward at
batch 128.  Needs libyq_b200.so built with  make NVFLAGS_EXTRA=-DYQ_L0_TRACE."""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
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
    buf = np.zeros(1024 * 16, np.uint64)
    get = lib.yq_debug_l0_trace
    get.restype = C.c_int
    get.argtypes = [C.c_void_p, C.c_size_t]
    assert get(buf.ctypes.data, buf.nbytes) == 0
    t = buf.reshape(1024, 16).astype(np.float64)
    t = t[t[:, 4] > 0]
    ghz = (t[:, 5] / np.maximum(t[:, 6], 1)).mean()
    print(f"{len(t)} CTAs, producer warp 0: {t[:, 4].mean():.1f} tiles, loop {t[:, 6].mean() / 1e3:.1f} us at {ghz:.2f} GHz; entry -> loop start {t[:, 12].mean() / 1e3:.1f} us; "
          f"first loop start .. last exit {(t[:, 14].max() - t[:, 13].min()) / 1e3:.1f} us; loop starts spread {(t[:, 13].max() - t[:, 13].min()) / 1e3:.1f} us")
    for i, nm in enumerate(["wait full (TMA landed)", "rearrange + fence", "refill + wait acc_empty", "MMA issue + commit"]):
        per = t[:, i] / t[:, 4]
        print(f"  producer {nm:>28s}: mean {per.mean():7.0f} clk per tile   p10 {np.percentile(per, 10):7.0f}   p90 {np.percentile(per, 90):7.0f}")
    e = t[t[:, 10] > 0]
    for i, nm in ((8, "wait acc_full"), (9, "tile body")):
        per = e[:, i] / e[:, 10]
        print(f"  epilogue {nm:>28s}: mean {per.mean():7.0f} clk per tile   p10 {np.percentile(per, 10):7.0f}   p90 {np.percentile(per, 90):7.0f}   ({e[:, 10].mean():.1f} tiles)")
    net.free()
