"""group the per-layer times of a `bench.py --net yolov3` line by layer shape (helper for profiles/)"""
import collections
import json
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from yolo_quantization_b200 import synth

d = json.load(open(sys.argv[1]))
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 4579.0
B = d["config"]["global_batch"] // d["n_gpus"]
info = synth.write_weights(None, synth.yolov3_quant())
g = collections.OrderedDict()
for r in d["layers"]:
    s = info[r["layer"]]
    key = (s.kind, s.c, s.out_c, s.spec.size if s.kind == "conv" else 0, s.spec.stride if s.kind == "conv" else 0, s.h, r["kernel"])
    e = g.setdefault(key, [0, 0.0])
    e[0] += 1
    e[1] += r["ms"]
tot = sum(v[1] for v in g.values())
print(f"{d['value']:.0f} img/s, {d['ms_per_step']:.3f} ms per {B}-image step; sum of per-layer events {tot:.3f} ms; int8 peak taken as {peak} TOP/s, HBM 6553 GB/s")
for k, v in sorted(g.items(), key=lambda kv: -kv[1][1]):
    kind, c, n, sz, st, h, kern = k
    oh = h // max(st, 1)
    t_tc = B * 2 * oh * oh * c * n * sz * sz / 1e12 / peak * 1e3 if kind == "conv" else 0
    byts = B * (h * h * c + oh * oh * n) if kind == "conv" else (3 * B * h * h * c if kind == "shortcut" else 2 * B * h * h * n)
    t_mem = byts / 6553e9 * 1e3
    each = v[1] / v[0]
    print(f"{kind:9s} c{c:4d} n{n:4d} k{sz} s{st} {h:3d}x{h:<3d} flavour {str(kern):4s} x{v[0]:2d}  {v[1]:.3f} ms ({100 * v[1] / tot:4.1f}%)  each {each:.4f}"
          f"  roofline {max(t_tc, t_mem):.4f} ({'tensor' if t_tc > t_mem else 'hbm'})  frac {max(t_tc, t_mem) / each:.2f}")
