"""Profiling driver (run under ncu on the GPU box): N warm-up forwards + 1 forward of yolov3-tiny at batch B
(YQ_NET=yolov3: the full yolov3, default batch 64).
    ncu --set full --import-source on -k regex:conv_u8 -s 26 -c 13 -o gpurun_out/prof python tools/prof_forward.py
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_quantization_b200 import darknet, synth  # noqa: E402

NET = os.environ.get("YQ_NET", "tiny")
B = int(os.environ.get("YQ_BATCH", "128" if NET == "tiny" else "64"))
WARM = int(os.environ.get("YQ_WARM", "2"))
KERNEL = int(os.environ.get("YQ_KERNEL", "-1"))
layers = synth.yolov3_tiny_quant() if NET == "tiny" else synth.yolov3_quant()
with tempfile.TemporaryDirectory() as d:
    cfg, wts = os.path.join(d, "t.cfg"), os.path.join(d, "t.weights")
    synth.write_cfg(cfg, layers, batch=B)
    synth.write_weights(wts, layers)
    net = darknet.load_network(cfg, wts, batch=B)
    if KERNEL >= 0:
        net.set_conv_kernel(KERNEL)
    x = np.random.default_rng(0).integers(0, 256, size=(B, 3, 416, 416), dtype=np.uint8)
    dev = darknet.DeviceBuffer.from_numpy(x)
    for _ in range(WARM + 1):
        net.forward_device(dev.ptr)
        net.synchronize()
    print("launches per forward", net.launches_per_forward)
    # the layer behind every launch of one forward, in issue order (tools/make_traffic.py maps ncu's launch list with it)
    import json
    print("LAUNCH_ORDER", json.dumps(net.launch_order()))
    if not os.environ.get("YQ_NO_PROFILE_FORWARD"):
        print("done", net.profile_forward(dev.ptr).round(3).tolist())
    net.free()
