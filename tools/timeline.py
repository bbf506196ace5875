"""Kernel timeline of the forward in steady state (CUDA-graph replays back to back, ONE stream): start / end of every kernel from
the CUPTI activity records torch.profiler collects -- per kernel of the forward: mean duration and the mean gap between the previous
kernel's end and its start (negative: programmatic dependent launch lets its prologue overlap the previous kernel's tail).
    YQ_NET=tiny|yolov3 python tools/timeline.py [out.json]"""
import json
import os
import sys
import tempfile

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_quantization_b200 import darknet, synth  # noqa: E402

NET = os.environ.get("YQ_NET", "tiny")
B = int(os.environ.get("YQ_BATCH", "128" if NET == "tiny" else "64"))
N = int(os.environ.get("YQ_FORWARDS", "20"))
layers = synth.yolov3_tiny_quant() if NET == "tiny" else synth.yolov3_quant()
with tempfile.TemporaryDirectory() as d:
    cfg, wts = os.path.join(d, "t.cfg"), os.path.join(d, "t.weights")
    synth.write_cfg(cfg, layers, batch=B)
    synth.write_weights(wts, layers)
    net = darknet.load_network(cfg, wts, batch=B)
    net.use_graph(os.environ.get("YQ_GRAPH", "1") != "0")
    xs = [darknet.DeviceBuffer.from_numpy(np.random.default_rng(i).integers(0, 256, size=(B, 3, 416, 416), dtype=np.uint8)) for i in range(4)]
    for i in range(10):
        net.forward_device(xs[i % 4].ptr)
    net.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(N):
            net.forward_device(xs[i % 4].ptr)
        net.synchronize()
    order = net.launch_order()
    per = net.launches_per_forward
    net.free()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
ev.sort(key=lambda e: e.time_range.start)
print(f"{len(ev)} kernel records, {per} launches per forward, {N} forwards")
if len(ev) != per * N:
    print("record count does not match launches x forwards; names:", sorted({e.name[:60] for e in ev})[:40])
    sys.exit(1)
dur = np.array([[e.time_range.end - e.time_range.start for e in ev[f * per:(f + 1) * per]] for f in range(N)], float)
start = np.array([[e.time_range.start for e in ev[f * per:(f + 1) * per]] for f in range(N)], float)
end = start + dur
gap = np.zeros_like(dur)
gap[:, 1:] = start[:, 1:] - end[:, :-1]
gap[1:, 0] = start[1:, 0] - end[:-1, -1]
span = (end[-1, -1] - start[1, 0]) / (N - 1)
rows = []
for k in range(per):
    rows.append({"launch": k, "layer": order[k], "kernel": ev[k].name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:70], "us": round(float(dur[1:, k].mean()), 2), "gap_us": round(float(gap[1:, k].mean()), 2)})
    print(f"{k:3d} layer {order[k]:3d}  {rows[-1]['us']:8.2f} us  gap {rows[-1]['gap_us']:7.2f}  {rows[-1]['kernel']}")
print("(launches are listed in start order; the layer column follows the issue order: a side-stream launch may sit under its neighbour's layer)")
print(f"forward period {span:.1f} us; sum of kernel durations {dur[1:].sum(1).mean():.1f} us; sum of positive gaps {np.clip(gap[1:], 0, None).sum(1).mean():.1f} us; "
      f"overlap (negative gaps) {-np.clip(gap[1:], None, 0).sum(1).mean():.1f} us")
if len(sys.argv) > 1:
    json.dump({"net": NET, "batch": B, "forwards": N, "period_us": span, "launches": rows}, open(sys.argv[1], "w"), indent=1)
