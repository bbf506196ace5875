"""debug helper (GPU box): full yolov3 at 416, which layer first differs from the oracle, per plan"""
import sys, os, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import yq_oracle as O
from yolo_quantization_b200 import darknet, synth

size = int(sys.argv[1]) if len(sys.argv) > 1 else 416
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
d = tempfile.mkdtemp()
layers = synth.yolov3_quant()
cfg, wts = os.path.join(d, "v3.cfg"), os.path.join(d, "v3.weights")
synth.write_cfg(cfg, layers, batch=batch, width=size, height=size)
info = synth.write_weights(wts, layers, width=size, height=size, seed=4, identity_bn=False)
imgs = np.stack([synth.synthetic_image(51 + b, 3, size, size) for b in range(batch)])
ref = O.forward_network(info, imgs[batch - 1])
net = darknet.load_network(cfg, wts, batch=batch)
for debug in (True, False):
    net.set_debug(debug)
    heads = net.split_heads(net.predict_u8(imgs))
    bad = []
    hi = 0
    for i, (sl, r) in enumerate(zip(info, ref)):
        li = net.layer_info(i)
        if sl.kind == "yolo":
            ok = np.allclose(heads[hi][batch - 1], r["f32"], atol=1e-6, rtol=0)
            hi += 1
        else:
            if not debug and (li.fused or (sl.kind == "conv" and sl.spec.quant_stop)):
                continue
            got = net.pull_layer(i, "u8")[batch - 1]
            ok = np.array_equal(got, r["u8"])
            if not ok and len(bad) < 6:
                diff = np.argwhere(got != r["u8"])
                print(f"  layer {i} {sl.kind} kernel {li.kernel} c{sl.c} {sl.h}x{sl.w}->{sl.out_c}: {len(diff)} of {got.size} differ; first {diff[:3].tolist()} "
                      f"rows {sorted(set(diff[:,1].tolist()))[:8]} cols {sorted(set(diff[:,2].tolist()))[:8]}")
            if debug and sl.kind == "conv":
                okacc = np.array_equal(net.pull_layer(i, "acc")[batch - 1], r["acc"])
                if not okacc:
                    print(f"  layer {i} acc differs")
        if not ok:
            bad.append(i)
    print("debug" if debug else "production", "plan: mismatching layers", bad[:20])
print("kinds", [(i, net.layer_info(i).kernel) for i in range(net.n) if net.layer_info(i).type == 0][:20])
