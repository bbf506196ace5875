"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/liboracle.so plus a layer-by-layer network
walk that mirrors the reference's ``forward_network`` (src/network.c:229-261) on the CPU.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module.  The product package (yolo_quantization_b200/) never does.

Also here: helpers to run the compiled reference (oracle/_ref/ref_harness, built by
``make -C oracle ref`` from the sources under /root/reference) and read its per-layer dumps.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
from typing import Dict, List, Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_HARNESS = os.path.join(HERE, "_ref", "ref_harness")
REF_HARNESS_OMP = os.path.join(HERE, "_ref", "ref_harness_omp")

ACT = {"logistic": 0, "relu": 1, "linear": 3, "relu6": 8, "leaky": 9}   # include/darknet.h:87-89

_lib = None


def build(force: bool = False) -> None:
    """Compile the C restatement (and, when /root/reference is present, the reference itself)."""
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
            os.path.join(HERE, "yq_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/src") and (force or not os.path.exists(REF_HARNESS)):
        subprocess.check_call(["make", "-C", HERE, "ref", "-j8"], stdout=subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def conv_acc(x: np.ndarray, w_u8: np.ndarray, zp_w: np.ndarray, stride: int, pad: int, zp_in: int,
             reffloat: bool = False) -> np.ndarray:
    """x: u8 [c,h,w]; w_u8: u8 [n,c,k,k]; returns int32 [n,oh,ow] (exact integer, or float-carried)."""
    x = np.ascontiguousarray(x, np.uint8)
    w_u8 = np.ascontiguousarray(w_u8, np.uint8)
    zp_w = np.ascontiguousarray(zp_w, np.uint8)
    c, h, w = x.shape
    n, _, k, _ = w_u8.shape
    oh = (h + 2 * pad - k) // stride + 1
    ow = (w + 2 * pad - k) // stride + 1
    out = np.empty((n, oh, ow), np.int32)
    fn = lib().yq_oracle_conv_acc_reffloat if reffloat else lib().yq_oracle_conv_acc
    fn(_p(x, C.c_uint8), c, h, w, _p(w_u8, C.c_uint8), _p(zp_w, C.c_uint8), n, k, stride, pad, int(zp_in),
       _p(out, C.c_int32))
    return out


def requant(acc: np.ndarray, bias_i32: np.ndarray, M_value: np.ndarray, rshift_value: np.ndarray,
            activation: int, zp_out: int) -> np.ndarray:
    acc = np.ascontiguousarray(acc, np.int32)
    n = acc.shape[0]
    spatial = int(np.prod(acc.shape[1:]))
    out = np.empty(acc.shape, np.uint8)
    lib().yq_oracle_requant(_p(acc, C.c_int32), n, spatial, _p(np.ascontiguousarray(bias_i32, np.int32), C.c_int32),
                            _p(np.ascontiguousarray(M_value, np.float64), C.c_double),
                            _p(np.ascontiguousarray(rshift_value, np.float64), C.c_double),
                            int(activation), int(zp_out), _p(out, C.c_uint8))
    return out


def dequant(u8: np.ndarray, zp_out: int, s_out: float) -> np.ndarray:
    u8 = np.ascontiguousarray(u8, np.uint8)
    out = np.empty(u8.shape, np.float32)
    lib().yq_oracle_dequant(_p(u8, C.c_uint8), C.c_size_t(u8.size), int(zp_out), C.c_float(s_out),
                            _p(out, C.c_float))
    return out


def maxpool(x: np.ndarray, size: int, stride: int, pad: Optional[int] = None) -> np.ndarray:
    x = np.ascontiguousarray(x, np.uint8)
    c, h, w = x.shape
    pad = size - 1 if pad is None else pad
    oh = (h + pad - size) // stride + 1
    ow = (w + pad - size) // stride + 1
    out = np.empty((c, oh, ow), np.uint8)
    lib().yq_oracle_maxpool(_p(x, C.c_uint8), c, h, w, size, stride, pad, _p(out, C.c_uint8))
    return out


def upsample(x: np.ndarray, stride: int) -> np.ndarray:
    x = np.ascontiguousarray(x, np.uint8)
    c, h, w = x.shape
    out = np.empty((c, h * stride, w * stride), np.uint8)
    lib().yq_oracle_upsample(_p(x, C.c_uint8), c, h, w, stride, _p(out, C.c_uint8))
    return out


def shortcut_mult(s_x: float, s_out: float) -> int:
    """K = round(s_x / s_out * 2^16) of the quantized-shortcut extension (oracle/yq_oracle.c: yq_oracle_shortcut_mult)."""
    k = C.c_int32()
    if lib().yq_oracle_shortcut_mult(C.c_float(s_x), C.c_float(s_out), C.byref(k)) != 0:
        raise ValueError(f"shortcut multiplier {s_x}/{s_out} outside [2^-16, 64)")
    return int(k.value)


def shortcut(a: np.ndarray, b: np.ndarray, q_a, q_b, q_out) -> np.ndarray:
    """Quantized shortcut (extension, not in the reference): a, b u8 of one shape; q_* = (scale, zero point)."""
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    assert a.shape == b.shape
    out = np.empty(a.shape, np.uint8)
    lib().yq_oracle_shortcut(_p(a, C.c_uint8), _p(b, C.c_uint8), C.c_size_t(a.size), int(q_a[1]), int(q_b[1]),
                             shortcut_mult(q_a[0], q_out[0]), shortcut_mult(q_b[0], q_out[0]), int(q_out[1]), _p(out, C.c_uint8))
    return out


def yolo(x: np.ndarray, n_anchors: int, classes: int) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    _, h, w = x.shape
    out = np.empty(x.shape, np.float32)
    lib().yq_oracle_yolo(_p(x, C.c_float), n_anchors, classes, h, w, _p(out, C.c_float))
    return out


def letterbox(im: np.ndarray, h: int, w: int) -> np.ndarray:
    """letterbox_image (src/image.c:812-831): float CHW image -> float [c, h, w] canvas (0.5 border)."""
    im = np.ascontiguousarray(im, np.float32)
    c, ih, iw = im.shape
    out = np.empty((c, h, w), np.float32)
    lib().yq_oracle_letterbox(_p(im, C.c_float), c, ih, iw, _p(out, C.c_float), h, w)
    return out


def quantize_input(x: np.ndarray):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty(x.shape, np.uint8)
    s = C.c_float()
    z = C.c_uint8()
    lib().yq_oracle_quantize_input(_p(x, C.c_float), C.c_size_t(x.size), _p(out, C.c_uint8), C.byref(s), C.byref(z))
    return out, float(s.value), int(z.value)


def prepare_conv(sl, s_in: float, zp_in: int) -> Dict[str, np.ndarray]:
    """Host prep of one conv (blas.c:282-334) from a synth.SynthLayer; returns M0/shift/M_value/... arrays."""
    n = sl.out_c
    K = sl.c * sl.spec.size * sl.spec.size
    M0 = np.empty(n, np.int32)
    sh = np.empty(n, np.int32)
    Mv = np.empty(n, np.float64)
    rs = np.empty(n, np.float64)
    bi = np.empty(n, np.int32)
    bn = int(sl.spec.bn)
    z = np.zeros(n, np.float32)
    lib().yq_oracle_prepare_conv(
        n, K, _p(np.ascontiguousarray(sl.w_u8), C.c_uint8), _p(sl.zp_w, C.c_uint8), _p(sl.s_w, C.c_float),
        C.c_float(s_in), int(zp_in), C.c_float(sl.s_out), _p(sl.biases, C.c_float), bn,
        _p(sl.bn_scales if bn else z, C.c_float), _p(sl.bn_mean if bn else z, C.c_float),
        _p(sl.bn_var if bn else z, C.c_float), _p(M0, C.c_int32), _p(sh, C.c_int), _p(Mv, C.c_double),
        _p(rs, C.c_double), _p(bi, C.c_int32))
    return {"M0": M0, "M0_right_shift": sh, "M_value": Mv, "M0_right_shift_value": rs, "biases_int32": bi}


ANCHORS = [float(v) for v in "25,39, 29,88, 405,102, 407,109,408,113,420,129".split(",")]


def yolo_boxes(heads, masks, classes: int, netw: int, neth: int, w: int, h: int, thresh: float = 0.5, nms: float = 0.45,
               relative: int = 1, anchors=ANCHORS) -> np.ndarray:
    """get_network_boxes (+ do_nms_sort when nms > 0) for ONE image. heads: list of f32 CHW yolo outputs in network order."""
    per = 5 + classes
    cap = sum(int(hd.shape[1] * hd.shape[2] * len(m)) for hd, m in zip(heads, masks))
    dets = np.zeros((cap, per), np.float32)
    n = 0
    for hd, m in zip(heads, masks):
        hd = np.ascontiguousarray(hd, np.float32)
        aw = np.array([anchors[2 * q] for q in m], np.float32)
        ah = np.array([anchors[2 * q + 1] for q in m], np.float32)
        sub = dets[n:]
        n += lib().yq_oracle_yolo_detections(_p(hd, C.c_float), hd.shape[2], hd.shape[1], len(m), classes, _p(aw, C.c_float),
                                             _p(ah, C.c_float), netw, neth, w, h, relative, C.c_float(thresh), _p(sub, C.c_float))
    dets = np.ascontiguousarray(dets[:n])
    if nms > 0 and n:
        lib().yq_oracle_nms_sort(_p(dets, C.c_float), n, classes, C.c_float(nms))
    return dets


def forward_network(info: Sequence, img_u8: np.ndarray, input_quant=(1.0 / 255.0, 0),
                    params_override: Optional[Dict[int, Dict[str, np.ndarray]]] = None,
                    reffloat: bool = False) -> List[Dict[str, np.ndarray]]:
    """Walk the net like forward_network (network.c:229-261) at batch 1.

    info: List[synth.SynthLayer]; img_u8: u8 CHW (already quantised layer-0 input).
    Returns one dict per layer with 'u8' (output_uint8_final), conv extras 'acc' / 'f32', yolo 'f32'.
    params_override[i] may carry the reference's dumped M_value / M0_right_shift_value / biases_int32.
    """
    outs: List[Dict[str, np.ndarray]] = []
    cur = np.ascontiguousarray(img_u8, np.uint8)
    cur_f32 = None
    prev_q = (float(np.float32(input_quant[0])), int(input_quant[1]))
    for i, sl in enumerate(info):
        o: Dict[str, np.ndarray] = {}
        if sl.kind == "conv":
            sp = sl.spec
            pad = sp.size // 2 if sp.pad else 0
            # conv i>0 takes (s_in, zp_in) from layer i-1's activation params (blas.c:301-305)
            s_in, zp_in = prev_q
            prm = prepare_conv(sl, s_in, zp_in)
            if params_override and i in params_override:
                prm.update(params_override[i])
            acc = conv_acc(cur, sl.w_u8, sl.zp_w, sp.stride, pad, zp_in, reffloat=reffloat)
            u8 = requant(acc, prm["biases_int32"], prm["M_value"], prm["M0_right_shift_value"],
                         ACT[sp.activation], sl.zp_out)
            o.update(acc=acc, u8=u8, **prm)
            if sp.quant_stop:
                o["f32"] = dequant(u8, sl.zp_out, sl.s_out)
                cur_f32 = o["f32"]
            cur = u8
        elif sl.kind == "maxpool":
            cur = maxpool(cur, sl.spec.size, sl.spec.stride)
            o["u8"] = cur
        elif sl.kind == "upsample":
            cur = upsample(cur, sl.spec.stride)
            o["u8"] = cur
        elif sl.kind == "route":
            cur = np.concatenate([outs[j]["u8"] for j in sl.inputs], axis=0)   # route_layer.c:107-117
            o["u8"] = cur
        elif sl.kind == "shortcut":
            j = sl.inputs[0]                                                   # extension: see yq_oracle_shortcut
            cur = shortcut(cur, outs[j]["u8"], prev_q, (info[j].s_out, info[j].zp_out), (sl.s_out, sl.zp_out))
            o["u8"] = cur
        elif sl.kind == "yolo":
            o["f32"] = yolo(cur_f32, len(sl.spec.mask), (sl.c // len(sl.spec.mask)) - 5)
        if sl.kind != "yolo":
            prev_q = (sl.s_out, sl.zp_out)
        outs.append(o)
    return outs


# ----------------------------------------------------------------------------------------------
# the compiled reference
# ----------------------------------------------------------------------------------------------

def have_reference() -> bool:
    return os.path.exists(REF_HARNESS)


def run_reference(mode: str, cfg: str, weights: str, input_f32: str, out: str, omp: bool = False,
                  threads: Optional[int] = None) -> str:
    exe = REF_HARNESS_OMP if omp else REF_HARNESS
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    r = subprocess.run([exe, mode, cfg, weights, input_f32, out], stdout=subprocess.DEVNULL,
                       stderr=subprocess.PIPE, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError(f"ref_harness failed ({r.returncode}): {r.stderr[-2000:]}")
    return r.stderr


def ref_times(stderr: str) -> List[float]:
    # (the reference writes its own progress text to stderr without a trailing newline: a REF_TIME record may not start a line)
    return [float(t) for t in re.findall(r"REF_TIME\s+\d+\s+([0-9.eE+-]+)", stderr)]


_DT = {"output_int32": np.int32, "output_uint8": np.uint8, "M0": np.int32, "M0_right_shift": np.int32,
       "M_value": np.float64, "M0_right_shift_value": np.float64, "biases_int32": np.int32,
       "weight_zero_point": np.uint8, "weight_scales": np.float32, "biases_folded": np.float32,
       "output_f32": np.float32, "input_uint8": np.uint8, "boxes_pre_nms": np.float32, "boxes_post_nms": np.float32}


def read_dump(dirname: str) -> List[Dict]:
    """Parse manifest.txt + L%02d_*.bin written by ref_harness into one dict per layer."""
    layers: List[Dict] = []
    with open(os.path.join(dirname, "manifest.txt")) as f:
        for line in f:
            t = line.split()
            if t[0] != "layer":
                continue
            d: Dict = {"index": int(t[1]), "type": t[3]}
            for k, v in zip(t[4::2], t[5::2]):
                d[k] = float(v) if ("." in v or "e" in v) else int(v)
            layers.append(d)
    for d in layers:
        i = d["index"]
        pre = f"L{i:02d}_"
        for fn in os.listdir(dirname):
            if fn.startswith(pre) and fn.endswith(".bin"):
                name = fn[len(pre):-4]
                arr = np.fromfile(os.path.join(dirname, fn), dtype=_DT[name])
                if name.startswith("output_") and d["type"] != "yolo":
                    arr = arr.reshape(d["out_c"], d["out_h"], d["out_w"])
                d[name] = arr
    return layers
