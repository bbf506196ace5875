"""Thin Python mirror of the reference's public C API for the quantized inference path, over the
C ABI of libyq_b200.so (include/yq_b200.h).  Python is glue only: every byte of compute happens in
the sm_100a kernels; nothing here touches the oracle or has a CPU fallback.

Reference names kept (include/darknet.h): ``load_network`` (:730), ``network_predict`` (:904),
``forward_network`` (:856), ``free_network`` (:907), ``set_batch_network`` (batch is fixed at load,
see ``load_network``), and per-layer ``forward_*_layer_quant_gpu`` forms.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import ActGeom, ConvDesc, LayerInfo, YqError, check

LAYER_TYPES = {0: "conv", 1: "maxpool", 2: "route", 3: "upsample", 4: "yolo", 5: "shortcut"}
ACTIVATIONS = {"logistic": 0, "relu": 1, "linear": 3, "relu6": 8, "leaky": 9}


def channel_stride(c: int) -> int:
    return _lib.load().yq_channel_stride(int(c))


class DeviceBuffer:
    """cuda_make_array / cuda_free (src/cuda.c:90-104,145-151) for raw bytes."""

    def __init__(self, nbytes: int, zero: bool = True):
        lib = _lib.load()
        _lib.require_gpu()
        self.nbytes = int(nbytes)
        self.ptr = lib.yq_cuda_malloc(self.nbytes)
        if not self.ptr:
            raise YqError(_lib.last_error())
        if zero:
            check(lib.yq_cuda_memset(self.ptr, 0, self.nbytes, None))

    @classmethod
    def from_numpy(cls, a: np.ndarray) -> "DeviceBuffer":
        a = np.ascontiguousarray(a)
        b = cls(a.nbytes, zero=False)
        b.push(a)
        return b

    def push(self, a: np.ndarray) -> None:
        a = np.ascontiguousarray(a)
        assert a.nbytes <= self.nbytes
        check(_lib.load().yq_cuda_push(self.ptr, a.ctypes.data, a.nbytes, None))
        check(_lib.load().yq_stream_synchronize(None))

    def pull(self, shape, dtype) -> np.ndarray:
        out = np.empty(shape, dtype)
        assert out.nbytes <= self.nbytes
        check(_lib.load().yq_cuda_pull(out.ctypes.data, self.ptr, out.nbytes, None))
        return out

    def free(self) -> None:
        if self.ptr:
            _lib.load().yq_cuda_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------
# layout helpers at the boundary (the reference's tensors are CHW per image)
# ----------------------------------------------------------------------------------------------

def push_nchw_u8(x: np.ndarray) -> DeviceBuffer:
    """Host uint8 [b,c,h,w] -> device NHWC with the library's channel stride."""
    x = np.ascontiguousarray(x, np.uint8)
    b, c, h, w = x.shape
    src = DeviceBuffer.from_numpy(x)
    dst = DeviceBuffer(b * h * w * channel_stride(c))
    check(_lib.load().yq_nchw_to_nhwc_u8(src.ptr, dst.ptr, b, c, h, w, None))
    check(_lib.load().yq_stream_synchronize(None))
    src.free()
    return dst


def pull_nhwc_u8(buf: DeviceBuffer, b: int, c: int, h: int, w: int) -> np.ndarray:
    tmp = DeviceBuffer(b * c * h * w)
    check(_lib.load().yq_nhwc_to_nchw_u8(buf.ptr, tmp.ptr, b, c, h, w, None))
    out = tmp.pull((b, c, h, w), np.uint8)
    tmp.free()
    return out


def pull_nhwc_i32(buf: DeviceBuffer, b: int, c: int, h: int, w: int) -> np.ndarray:
    tmp = DeviceBuffer(b * c * h * w * 4)
    check(_lib.load().yq_nhwc_to_nchw_i32(buf.ptr, tmp.ptr, b, c, h, w, None))
    out = tmp.pull((b, c, h, w), np.int32)
    tmp.free()
    return out


def letterbox_image(images: np.ndarray, h: int, w: int) -> np.ndarray:
    """letterbox_image (src/image.c:812-831) on the device: float [b,c,ih,iw] -> float [b,c,h,w]."""
    lib = _lib.load()
    x = np.ascontiguousarray(images, np.float32)
    b, c, ih, iw = x.shape
    din = DeviceBuffer.from_numpy(x)
    dout = DeviceBuffer(b * c * h * w * 4)
    check(lib.yq_letterbox_image_gpu(din.ptr, b, c, ih, iw, dout.ptr, h, w, None), "yq_letterbox_image_gpu")
    check(lib.yq_stream_synchronize(None))
    out = dout.pull((b, c, h, w), np.float32)
    din.free(); dout.free()
    return out


def quantize_input(x: np.ndarray):
    """quant_weights_with_min_max_channel(1, net->input, ...) (src/blas.c:108-168) on the device, per image.
    x: float32 [b, ...]; returns (uint8 array of x's shape, scales [b], zero points [b])."""
    lib = _lib.load()
    x = np.ascontiguousarray(x, np.float32)
    b, n = x.shape[0], int(np.prod(x.shape[1:]))
    din = DeviceBuffer.from_numpy(x)
    dout = DeviceBuffer(b * n)
    meta = DeviceBuffer(16 * b)
    check(lib.yq_quantize_input_gpu(din.ptr, dout.ptr, meta.ptr, meta.ptr + 4 * b, meta.ptr + 8 * b, b, n, None), "yq_quantize_input_gpu")
    check(lib.yq_stream_synchronize(None))
    u8 = dout.pull(x.shape, np.uint8)
    raw = meta.pull((4 * b,), np.int32)
    scales, zps = raw[:b].view(np.float32).copy(), raw[b:2 * b].copy()
    for d in (din, dout, meta):
        d.free()
    return u8, scales, zps


# ----------------------------------------------------------------------------------------------
# layer level
# ----------------------------------------------------------------------------------------------

class ConvolutionalLayerQuant:
    """make_convolutional_layer's quantized GPU state + forward_convolutional_layer_quant_gpu.

    Arguments follow ``struct layer`` after quantization_weights_and_activations (src/blas.c:259-346):
    weights_uint8 OIHW, per-channel weight zero points, biases_int32, M_value, M0_right_shift_value.
    """

    def __init__(self, h: int, w: int, c: int, n: int, size: int, stride: int, pad: int, activation: int,
                 weights_uint8: np.ndarray, weight_zero_point: np.ndarray, biases_int32: np.ndarray,
                 M_value: np.ndarray, M0_right_shift_value: np.ndarray, zp_in: int, zp_out: int, s_out: float,
                 quant_stop_flag: int = 0, saturate: int = 0, kernel: int = -1):
        lib = _lib.load()
        _lib.require_gpu()
        self._keep = [np.ascontiguousarray(weights_uint8, np.uint8), np.ascontiguousarray(weight_zero_point, np.uint8),
                      np.ascontiguousarray(biases_int32, np.int32), np.ascontiguousarray(M_value, np.float64),
                      np.ascontiguousarray(M0_right_shift_value, np.float64)]
        assert self._keep[0].size == n * c * size * size
        d = ConvDesc(h, w, c, n, size, stride, pad, int(activation), int(quant_stop_flag), int(zp_in), int(zp_out),
                     float(s_out), *(a.ctypes.data for a in self._keep), int(saturate))
        self.handle = lib.yq_make_convolutional_layer_quant(C.byref(d))
        if not self.handle:
            raise YqError(_lib.last_error())
        self.h, self.w, self.c, self.n = h, w, c, n
        self.zp_in = int(zp_in) & 255
        self.out_h, self.out_w = lib.yq_conv_out_h(self.handle), lib.yq_conv_out_w(self.handle)
        self.quant_stop_flag = quant_stop_flag
        if kernel >= 0:
            check(lib.yq_conv_set_kernel(self.handle, kernel), "yq_conv_set_kernel")

    @property
    def kernel(self) -> int:
        return _lib.load().yq_conv_get_kernel(self.handle)

    def forward(self, x_nchw: np.ndarray, want_acc: bool = True) -> Dict[str, np.ndarray]:
        """x: uint8 [b,c,h,w] (reference layout). Returns dict(u8=[b,n,oh,ow], acc=int32 ..., f32=...)."""
        lib = _lib.load()
        b = x_nchw.shape[0]
        din = push_nchw_u8(x_nchw)
        cs = channel_stride(self.n)
        dout = DeviceBuffer(b * self.out_h * self.out_w * cs)
        dacc = DeviceBuffer(b * self.out_h * self.out_w * cs * 4) if want_acc else None
        df32 = DeviceBuffer(b * self.n * self.out_h * self.out_w * 4) if self.quant_stop_flag else None
        check(lib.yq_forward_convolutional_layer_quant_gpu(self.handle, din.ptr, dout.ptr, df32.ptr if df32 else None,
                                                           dacc.ptr if dacc else None, b, None),
              "yq_forward_convolutional_layer_quant_gpu")
        check(lib.yq_stream_synchronize(None))
        res = {"u8": pull_nhwc_u8(dout, b, self.n, self.out_h, self.out_w)}
        if dacc:
            res["acc"] = pull_nhwc_i32(dacc, b, self.n, self.out_h, self.out_w)
        if df32:
            res["f32"] = df32.pull((b, self.n, self.out_h, self.out_w), np.float32)
        for d in (din, dout, dacc, df32):
            if d:
                d.free()
        return res

    @property
    def can_fuse_maxpool(self) -> bool:
        return bool(_lib.load().yq_conv_can_fuse_maxpool(self.handle))

    def forward_pooled(self, x_nchw: np.ndarray, want_conv: bool = True) -> Dict[str, np.ndarray]:
        """conv + fused maxpool(2,2): returns dict(pool=[b,n,ceil(oh/2),ceil(ow/2)], u8=... when want_conv)."""
        lib = _lib.load()
        b = x_nchw.shape[0]
        din = push_nchw_u8(x_nchw)
        cs = channel_stride(self.n)
        ph, pw = (self.out_h + 1) // 2, (self.out_w + 1) // 2
        dout = DeviceBuffer(b * self.out_h * self.out_w * cs) if want_conv else None
        dpool = DeviceBuffer(b * ph * pw * cs)
        df32 = DeviceBuffer(b * self.n * self.out_h * self.out_w * 4) if self.quant_stop_flag else None
        check(lib.yq_forward_convolutional_layer_quant_pool_gpu(self.handle, din.ptr, dout.ptr if dout else None, dpool.ptr,
                                                                df32.ptr if df32 else None, None, b, None),
              "yq_forward_convolutional_layer_quant_pool_gpu")
        check(lib.yq_stream_synchronize(None))
        res = {"pool": pull_nhwc_u8(dpool, b, self.n, ph, pw)}
        if dout:
            res["u8"] = pull_nhwc_u8(dout, b, self.n, self.out_h, self.out_w)
        for d in (din, dout, dpool, df32):
            if d:
                d.free()
        return res

    @property
    def rows_supported(self) -> bool:
        return bool(_lib.load().yq_conv_rows_supported(self.handle))

    @property
    def rows_variant(self) -> int:
        """0: no rows flavour; 1: all-ones filter rows + per-output zero-point correction; 2: two signed weight blocks."""
        return int(_lib.load().yq_conv_rows_supported(self.handle))

    @property
    def rows_nchw_supported(self) -> bool:
        """the rows flavour can read the [b,3,h,w] planes directly (c = 3, zp_in = 0, w % 16 = 0, w >= 64)"""
        return bool(_lib.load().yq_conv_rows_nchw_supported(self.handle))

    def forward_rows_pooled(self, x_nchw: np.ndarray, out_pad: int = 0, nchw: bool = False) -> np.ndarray:
        """conv + RELU6 + maxpool(2,2) through the halo-input "rows" flavour: the input is staged in the padded
        geometry the layer asks for (halo = zp_in) -- or, with ``nchw``, handed over as the plain planes; the pooled
        tensor is written into a tensor with an ``out_pad``-pixel halo (as the next rows layer would want) and
        returned as [b,n,h/2,w/2]."""
        lib = _lib.load()
        b = x_nchw.shape[0]
        x = np.ascontiguousarray(x_nchw, np.uint8)
        g = ActGeom()
        check(lib.yq_conv_rows_input_geom(self.handle, C.byref(g)), "yq_conv_rows_input_geom")
        din = DeviceBuffer(lib.yq_act_geom_bytes(C.byref(g), b, self.c), zero=False)
        check(lib.yq_cuda_memset(din.ptr, self.zp_in, din.nbytes, None))
        src = DeviceBuffer.from_numpy(x)
        if not nchw:
            check(lib.yq_nchw_to_nhwc_u8_geom(src.ptr, din.ptr, b, self.c, self.h, self.w, C.byref(g), None))
        ph, pw = self.out_h // 2, self.out_w // 2
        og = ActGeom(out_pad, pw + 2 * out_pad + (3 if out_pad else 0), ph + 2 * out_pad + (1 if out_pad else 0))
        dout = DeviceBuffer(lib.yq_act_geom_bytes(C.byref(og), b, self.n), zero=False)
        check(lib.yq_cuda_memset(dout.ptr, 0xEE, dout.nbytes, None))
        if nchw:
            check(lib.yq_forward_convolutional_layer_quant_rows_pool_nchw_gpu(self.handle, src.ptr, dout.ptr, C.byref(og), b, None),
                  "yq_forward_convolutional_layer_quant_rows_pool_nchw_gpu")
        else:
            check(lib.yq_forward_convolutional_layer_quant_rows_pool_gpu(self.handle, din.ptr, dout.ptr, C.byref(og), b, None),
                  "yq_forward_convolutional_layer_quant_rows_pool_gpu")
        tmp = DeviceBuffer(b * self.n * ph * pw)
        check(lib.yq_nhwc_to_nchw_u8_geom(dout.ptr, tmp.ptr, b, self.n, ph, pw, C.byref(og), None))
        check(lib.yq_stream_synchronize(None))
        out = tmp.pull((b, self.n, ph, pw), np.uint8)
        raw = dout.pull((dout.nbytes,), np.uint8)
        for d in (din, src, dout, tmp):
            d.free()
        self.last_rows_raw = (raw, og.pad, og.pitch_w, og.rows_h)
        return out

    @property
    def flat_supported(self) -> bool:
        return bool(_lib.load().yq_conv_flat_supported(self.handle))

    def forward_flat(self, x_nchw: np.ndarray, halo_fill: int = 0, want_acc: bool = True, yolo_classes: Optional[int] = None
                     ) -> Dict[str, np.ndarray]:
        """The flat-strip flavour: input staged as a flat halo-padded tensor (halo = zp_in), output read back from one.
        Returns dict(u8, acc, f32 as in forward(), halo_ok=True when every halo byte of the output equals halo_fill).
        yolo_classes (quant_stop layers): run the layer as a detection head with the following [yolo] layer fused
        (yq_forward_convolutional_layer_quant_flat_yolo_gpu); adds yolo=[b,n,h,w] float32.  Without side outputs
        (want_acc=False) this is the production launch of the heads: f32 is then not materialised."""
        lib = _lib.load()
        b = x_nchw.shape[0]
        x = np.ascontiguousarray(x_nchw, np.uint8)
        g = ActGeom()
        check(lib.yq_act_geom_flat(self.h, self.w, C.byref(g)), "yq_act_geom_flat")
        cs_in, cs_out = channel_stride(self.c), channel_stride(self.n)
        din = DeviceBuffer(lib.yq_act_geom_bytes(C.byref(g), b, self.c), zero=False)
        check(lib.yq_cuda_memset(din.ptr, self.zp_in, din.nbytes, None))
        src = DeviceBuffer.from_numpy(x)
        check(lib.yq_nchw_to_nhwc_u8_geom(src.ptr, din.ptr, b, self.c, self.h, self.w, C.byref(g), None))
        dout = DeviceBuffer(lib.yq_act_geom_bytes(C.byref(g), b, self.n), zero=False)
        check(lib.yq_cuda_memset(dout.ptr, 0xEE, dout.nbytes, None))
        dacc = DeviceBuffer(b * self.out_h * self.out_w * cs_out * 4) if want_acc else None
        df32 = DeviceBuffer(b * self.n * self.out_h * self.out_w * 4) if self.quant_stop_flag and (yolo_classes is None or want_acc) else None
        dyolo = None
        if yolo_classes is not None:
            dyolo = DeviceBuffer(b * self.n * self.out_h * self.out_w * 4)
            check(lib.yq_forward_convolutional_layer_quant_flat_yolo_gpu(self.handle, din.ptr, dout.ptr, int(halo_fill), df32.ptr if df32 else None,
                                                                         dyolo.ptr, int(yolo_classes), dacc.ptr if dacc else None, b, None),
                  "yq_forward_convolutional_layer_quant_flat_yolo_gpu")
        else:
            check(lib.yq_forward_convolutional_layer_quant_flat_gpu(self.handle, din.ptr, dout.ptr, int(halo_fill), df32.ptr if df32 else None,
                                                                    dacc.ptr if dacc else None, b, None),
                  "yq_forward_convolutional_layer_quant_flat_gpu")
        tmp = DeviceBuffer(b * self.n * self.out_h * self.out_w)
        check(lib.yq_nhwc_to_nchw_u8_geom(dout.ptr, tmp.ptr, b, self.n, self.out_h, self.out_w, C.byref(g), None))
        check(lib.yq_stream_synchronize(None))
        res = {"u8": tmp.pull((b, self.n, self.out_h, self.out_w), np.uint8)}
        if dacc:
            res["acc"] = pull_nhwc_i32(dacc, b, self.n, self.out_h, self.out_w)
        if df32:
            res["f32"] = df32.pull((b, self.n, self.out_h, self.out_w), np.float32)
        if dyolo:
            res["yolo"] = dyolo.pull((b, self.n, self.out_h, self.out_w), np.float32)
        raw = dout.pull((dout.nbytes // cs_out, cs_out), np.uint8)
        rows = raw[: b * g.rows_h * g.pitch_w].reshape(b, g.rows_h, g.pitch_w, cs_out).copy()
        interior = rows[:, 1:1 + self.out_h, 1:1 + self.out_w, :].copy()
        rows[:, 1:1 + self.out_h, 1:1 + self.out_w, :self.n] = halo_fill
        tail = raw[b * g.rows_h * g.pitch_w:]
        res["halo_ok"] = bool((rows[..., :self.n] == halo_fill).all() and (tail[:, :self.n] == halo_fill).all()
                              and not rows[..., self.n:].any() and not tail[:, self.n:].any() and not interior[..., self.n:].any())
        for d in (din, src, dout, dacc, df32, dyolo, tmp):
            if d:
                d.free()
        return res

    @property
    def patch_supported(self) -> bool:
        """the layer has the resident-bank kernel in patch mode (narrow 3x3 stride-2 layers between halo-padded tensors)"""
        return bool(_lib.load().yq_conv_patch_supported(self.handle))

    @property
    def flat_up2_supported(self) -> bool:
        return bool(_lib.load().yq_conv_flat_up2_supported(self.handle))

    def forward_flat_up2(self, x_nchw: np.ndarray) -> Dict[str, np.ndarray]:
        """The flat 1x1 convolution with the following stride-2 upsample fused (yq_forward_convolutional_layer_quant_flat_up2_gpu).
        Returns dict(u8=[b,n,2h,2w], halo_ok=True when the upsampled strip's halo and pad lanes still hold their preset)."""
        lib = _lib.load()
        b = x_nchw.shape[0]
        x = np.ascontiguousarray(x_nchw, np.uint8)
        g, g2 = ActGeom(), ActGeom()
        check(lib.yq_act_geom_flat(self.h, self.w, C.byref(g)), "yq_act_geom_flat")
        check(lib.yq_act_geom_flat(2 * self.out_h, 2 * self.out_w, C.byref(g2)), "yq_act_geom_flat")
        cs_out = channel_stride(self.n)
        din = DeviceBuffer(lib.yq_act_geom_bytes(C.byref(g), b, self.c), zero=False)
        check(lib.yq_cuda_memset(din.ptr, self.zp_in, din.nbytes, None))
        src = DeviceBuffer.from_numpy(x)
        check(lib.yq_nchw_to_nhwc_u8_geom(src.ptr, din.ptr, b, self.c, self.h, self.w, C.byref(g), None))
        dout = DeviceBuffer(lib.yq_act_geom_bytes(C.byref(g2), b, self.n), zero=False)
        check(lib.yq_cuda_memset(dout.ptr, 0xEE, dout.nbytes, None))
        check(lib.yq_forward_convolutional_layer_quant_flat_up2_gpu(self.handle, din.ptr, dout.ptr, b, None),
              "yq_forward_convolutional_layer_quant_flat_up2_gpu")
        tmp = DeviceBuffer(b * self.n * 4 * self.out_h * self.out_w)
        check(lib.yq_nhwc_to_nchw_u8_geom(dout.ptr, tmp.ptr, b, self.n, 2 * self.out_h, 2 * self.out_w, C.byref(g2), None))
        check(lib.yq_stream_synchronize(None))
        res = {"u8": tmp.pull((b, self.n, 2 * self.out_h, 2 * self.out_w), np.uint8)}
        raw = dout.pull((dout.nbytes // cs_out, cs_out), np.uint8)
        rows = raw[: b * g2.rows_h * g2.pitch_w].reshape(b, g2.rows_h, g2.pitch_w, cs_out).copy()
        pads = rows[:, 1:, 1:, self.n:].copy()
        rows[:, 1:, 1:, :] = 0xEE
        res["halo_ok"] = bool((rows == 0xEE).all() and (raw[b * g2.rows_h * g2.pitch_w:] == 0xEE).all() and not pads.any())
        for d in (din, src, dout, tmp):
            d.free()
        return res

    def flat_cat_supported(self, c_first: int) -> bool:
        return bool(_lib.load().yq_conv_flat_cat_supported(self.handle, int(c_first)))

    def forward_flat_cat(self, x_nchw: np.ndarray, c_first: int, halo_fill: int = 0) -> Dict[str, np.ndarray]:
        """forward_flat() with the input handed over as TWO flat tensors, channels [0, c_first) and [c_first, c): the convolution
        behind a route that is never materialised (yq_forward_convolutional_layer_quant_flat_cat_gpu).  Returns dict(u8)."""
        lib = _lib.load()
        b = x_nchw.shape[0]
        x = np.ascontiguousarray(x_nchw, np.uint8)
        g = ActGeom()
        check(lib.yq_act_geom_flat(self.h, self.w, C.byref(g)), "yq_act_geom_flat")
        parts = []
        for lo, hi in ((0, c_first), (c_first, self.c)):
            d = DeviceBuffer(lib.yq_act_geom_bytes(C.byref(g), b, hi - lo), zero=False)
            check(lib.yq_cuda_memset(d.ptr, self.zp_in, d.nbytes, None))
            src = DeviceBuffer.from_numpy(np.ascontiguousarray(x[:, lo:hi]))
            check(lib.yq_nchw_to_nhwc_u8_geom(src.ptr, d.ptr, b, hi - lo, self.h, self.w, C.byref(g), None))
            parts += [d, src]
        dout = DeviceBuffer(lib.yq_act_geom_bytes(C.byref(g), b, self.n), zero=False)
        check(lib.yq_cuda_memset(dout.ptr, 0xEE, dout.nbytes, None))
        check(lib.yq_forward_convolutional_layer_quant_flat_cat_gpu(self.handle, parts[0].ptr, int(c_first), parts[2].ptr, dout.ptr, int(halo_fill), b, None),
              "yq_forward_convolutional_layer_quant_flat_cat_gpu")
        tmp = DeviceBuffer(b * self.n * self.out_h * self.out_w)
        check(lib.yq_nhwc_to_nchw_u8_geom(dout.ptr, tmp.ptr, b, self.n, self.out_h, self.out_w, C.byref(g), None))
        check(lib.yq_stream_synchronize(None))
        res = {"u8": tmp.pull((b, self.n, self.out_h, self.out_w), np.uint8)}
        for d in parts + [dout, tmp]:
            d.free()
        return res

    @property
    def geom_supported(self) -> bool:
        """the layer's flavour reads / writes halo-padded tensors of any geometry (the per-tap TMA flavour)"""
        return bool(_lib.load().yq_conv_geom_supported(self.handle))

    def forward_geom(self, x_nchw: np.ndarray, in_geom=None, out_geom=None, in_fill: Optional[int] = None, want_acc: bool = True
                     ) -> Dict[str, np.ndarray]:
        """forward() between halo-padded tensors: in_geom / out_geom = (pad, pitch_w, rows_h) or "flat" or None (plain).
        The input's halo is filled with ``in_fill`` (default zp_in) and announced as such; in_fill != zp_in exercises the
        out-of-bounds + border-correction form on a padded tensor.  Returns dict(u8, acc, f32, halo_ok): halo_ok = the
        output's halo still holds the 0xEE it was preset to (only the interior may be written)."""
        lib = _lib.load()
        b = x_nchw.shape[0]
        x = np.ascontiguousarray(x_nchw, np.uint8)
        cs_out = channel_stride(self.n)

        def geom(g, h, w):
            if g is None:
                return ActGeom(0, w, h)
            if g == "flat":
                r = ActGeom()
                check(lib.yq_act_geom_flat(h, w, C.byref(r)))
                return r
            return ActGeom(*g)
        gi, go = geom(in_geom, self.h, self.w), geom(out_geom, self.out_h, self.out_w)
        fill = self.zp_in if in_fill is None else int(in_fill)
        din = DeviceBuffer(lib.yq_act_geom_bytes(C.byref(gi), b, self.c), zero=False)
        check(lib.yq_cuda_memset(din.ptr, fill, din.nbytes, None))
        src = DeviceBuffer.from_numpy(x)
        check(lib.yq_nchw_to_nhwc_u8_geom(src.ptr, din.ptr, b, self.c, self.h, self.w, C.byref(gi), None))
        dout = DeviceBuffer(lib.yq_act_geom_bytes(C.byref(go), b, self.n), zero=False)
        check(lib.yq_cuda_memset(dout.ptr, 0xEE, dout.nbytes, None))
        dacc = DeviceBuffer(b * self.out_h * self.out_w * cs_out * 4) if want_acc else None
        df32 = DeviceBuffer(b * self.n * self.out_h * self.out_w * 4) if self.quant_stop_flag else None
        check(lib.yq_forward_convolutional_layer_quant_geom_gpu(self.handle, din.ptr, C.byref(gi), fill if gi.pad else -1, dout.ptr, C.byref(go),
                                                                df32.ptr if df32 else None, dacc.ptr if dacc else None, b, None),
              "yq_forward_convolutional_layer_quant_geom_gpu")
        tmp = DeviceBuffer(b * self.n * self.out_h * self.out_w)
        check(lib.yq_nhwc_to_nchw_u8_geom(dout.ptr, tmp.ptr, b, self.n, self.out_h, self.out_w, C.byref(go), None))
        check(lib.yq_stream_synchronize(None))
        res = {"u8": tmp.pull((b, self.n, self.out_h, self.out_w), np.uint8)}
        if dacc:
            res["acc"] = pull_nhwc_i32(dacc, b, self.n, self.out_h, self.out_w)
        if df32:
            res["f32"] = df32.pull((b, self.n, self.out_h, self.out_w), np.float32)
        raw = dout.pull((dout.nbytes // cs_out, cs_out), np.uint8)
        rows = raw[: b * go.rows_h * go.pitch_w].reshape(b, go.rows_h, go.pitch_w, cs_out).copy()
        rows[:, go.pad:go.pad + self.out_h, go.pad:go.pad + self.out_w, :] = 0xEE
        res["halo_ok"] = bool((rows == 0xEE).all() and (raw[b * go.rows_h * go.pitch_w:] == 0xEE).all())
        for d in (din, src, dout, dacc, df32, tmp):
            if d:
                d.free()
        return res

    def free(self) -> None:
        if self.handle:
            _lib.load().yq_free_convolutional_layer_quant(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def forward_maxpool_layer_quant_gpu(x: np.ndarray, size: int, stride: int, pad: Optional[int] = None) -> np.ndarray:
    b, c, h, w = x.shape
    pad = size - 1 if pad is None else pad
    oh, ow = (h + pad - size) // stride + 1, (w + pad - size) // stride + 1
    din = push_nchw_u8(x)
    dout = DeviceBuffer(b * oh * ow * channel_stride(c))
    check(_lib.load().yq_forward_maxpool_layer_quant_gpu(din.ptr, dout.ptr, b, h, w, c, size, stride, pad, None))
    out = pull_nhwc_u8(dout, b, c, oh, ow)
    din.free(); dout.free()
    return out


def forward_upsample_layer_quant_gpu(x: np.ndarray, stride: int) -> np.ndarray:
    b, c, h, w = x.shape
    din = push_nchw_u8(x)
    dout = DeviceBuffer(b * h * stride * w * stride * channel_stride(c))
    check(_lib.load().yq_forward_upsample_layer_quant_gpu(din.ptr, dout.ptr, b, h, w, c, stride, None))
    out = pull_nhwc_u8(dout, b, c, h * stride, w * stride)
    din.free(); dout.free()
    return out


def forward_route_layer_quant_gpu(xs: Sequence[np.ndarray], ups: Optional[Sequence[int]] = None,
                                 parts: Optional[Sequence[int]] = None, fill: int = 0) -> np.ndarray:
    """Channel concat (route_layer.c:107-130).  ups[k] > 1: input k is stored at 1/ups[k] of the route's size and is
    read through the nearest-neighbour upsample (upsample_layer.c:96-113 folded into the route).  parts: input masks,
    one launch each, into an output pre-set to `fill` (the inputs of one route written as they become ready)."""
    ups = [1] * len(xs) if ups is None else [int(u) for u in ups]
    b, _, h, w = xs[0].shape
    h, w = h * ups[0], w * ups[0]
    dins = [push_nchw_u8(x) for x in xs]
    cs = [int(x.shape[1]) for x in xs]
    ctot = sum(cs)
    dout = DeviceBuffer(b * h * w * channel_stride(ctot))
    ptrs = (C.c_void_p * len(xs))(*[d.ptr for d in dins])
    carr = (C.c_int * len(xs))(*cs)
    uarr = (C.c_int * len(xs))(*ups)
    if parts is not None:
        check(_lib.load().yq_cuda_memset(dout.ptr, int(fill), dout.nbytes, None))
        for m in parts:
            check(_lib.load().yq_forward_route_layer_quant_part_gpu(ptrs, None, carr, uarr, len(xs), int(m), dout.ptr, None, b, h, w, None))
    elif any(u != 1 for u in ups):
        check(_lib.load().yq_forward_route_layer_quant_up_gpu(ptrs, None, carr, uarr, len(xs), dout.ptr, None, b, h, w, None))
    else:
        check(_lib.load().yq_forward_route_layer_quant_gpu(ptrs, carr, len(xs), dout.ptr, b, h, w, None))
    out = pull_nhwc_u8(dout, b, ctot, h, w)
    for d in dins:
        d.free()
    dout.free()
    return out


def shortcut_multiplier(s_x: float, s_out: float) -> int:
    k = C.c_int32()
    check(_lib.load().yq_shortcut_multiplier(float(s_x), float(s_out), C.byref(k)), "yq_shortcut_multiplier")
    return int(k.value)


def forward_shortcut_layer_quant_gpu(a: np.ndarray, b: np.ndarray, q_a, q_b, q_out, geoms=None) -> np.ndarray:
    """The quantized shortcut (extension: include/yq_b200.h).  a, b uint8 [n,c,h,w]; q_* = (scale, zero point).
    geoms: optional ((pad, pitch_w, rows_h),) * 3 for a, b and the output (halo-padded tensors)."""
    lib = _lib.load()
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    n, c, h, w = a.shape
    ka, kb = shortcut_multiplier(q_a[0], q_out[0]), shortcut_multiplier(q_b[0], q_out[0])
    if geoms is None:
        da, db = push_nchw_u8(a), push_nchw_u8(b)
        dout = DeviceBuffer(n * h * w * channel_stride(c))
        check(lib.yq_forward_shortcut_layer_quant_gpu(da.ptr, db.ptr, dout.ptr, n, h, w, c, int(q_a[1]), int(q_b[1]), ka, kb, int(q_out[1]), None),
              "yq_forward_shortcut_layer_quant_gpu")
        out = pull_nhwc_u8(dout, n, c, h, w)
        for d in (da, db, dout):
            d.free()
        return out
    gs = [ActGeom(*g) for g in geoms]
    bufs = []
    for x, g in zip((a, b), gs[:2]):
        d = DeviceBuffer(lib.yq_act_geom_bytes(C.byref(g), n, c), zero=False)
        check(lib.yq_cuda_memset(d.ptr, 0x5A, d.nbytes, None))
        src = DeviceBuffer.from_numpy(x)
        check(lib.yq_nchw_to_nhwc_u8_geom(src.ptr, d.ptr, n, c, h, w, C.byref(g), None))
        check(lib.yq_stream_synchronize(None))
        src.free()
        bufs.append(d)
    dout = DeviceBuffer(lib.yq_act_geom_bytes(C.byref(gs[2]), n, c), zero=False)
    check(lib.yq_cuda_memset(dout.ptr, 0xEE, dout.nbytes, None))
    check(lib.yq_forward_shortcut_layer_quant_geom_gpu(bufs[0].ptr, C.byref(gs[0]), bufs[1].ptr, C.byref(gs[1]), dout.ptr, C.byref(gs[2]), n, h, w, c,
                                                       int(q_a[1]), int(q_b[1]), ka, kb, int(q_out[1]), None), "yq_forward_shortcut_layer_quant_geom_gpu")
    tmp = DeviceBuffer(n * c * h * w)
    check(lib.yq_nhwc_to_nchw_u8_geom(dout.ptr, tmp.ptr, n, c, h, w, C.byref(gs[2]), None))
    check(lib.yq_stream_synchronize(None))
    out = tmp.pull((n, c, h, w), np.uint8)
    for d in bufs + [dout, tmp]:
        d.free()
    return out


def forward_yolo_layer_gpu(x: np.ndarray, n_anchors: int, classes: int) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    b, c, h, w = x.shape
    din = DeviceBuffer.from_numpy(x)
    dout = DeviceBuffer(x.nbytes)
    check(_lib.load().yq_forward_yolo_layer_gpu(din.ptr, dout.ptr, b, n_anchors, classes, h, w, None))
    out = dout.pull(x.shape, np.float32)
    din.free(); dout.free()
    return out


# ----------------------------------------------------------------------------------------------
# network level
# ----------------------------------------------------------------------------------------------

class Network:
    """``network *`` of the reference for the quantized inference path."""

    def __init__(self, handle):
        self._h = handle
        lib = _lib.load()
        self.batch = lib.yq_network_batch(handle)
        c, h, w = C.c_int(), C.c_int(), C.c_int()
        lib.yq_network_input_dims(handle, C.byref(c), C.byref(h), C.byref(w))
        self.c, self.h, self.w = c.value, h.value, w.value
        self.n = lib.yq_network_num_layers(handle)
        self.output_floats = lib.yq_network_output_floats(handle)

    @property
    def handle(self):
        return self._h

    @property
    def stream(self) -> int:
        return _lib.load().yq_network_stream(self._h) or 0

    @property
    def launches_per_forward(self) -> int:
        return _lib.load().yq_network_launches_per_forward(self._h)

    def launch_order(self) -> List[int]:
        """the layer behind every launch of one forward, in issue order (-1 = the input layout transform)"""
        lib = _lib.load()
        per = [lib.yq_network_layer_launches(self._h, i) for i in range(self.n)]
        return [-1] * (self.launches_per_forward - sum(per)) + [i for i, k in enumerate(per) for _ in range(k)]

    def layer_info(self, i: int) -> LayerInfo:
        info = LayerInfo()
        check(_lib.load().yq_network_layer_info(self._h, i, C.byref(info)))
        return info

    def layers(self) -> List[LayerInfo]:
        return [self.layer_info(i) for i in range(self.n)]

    def set_input_quant(self, s_in: float, zp_in: int) -> None:
        check(_lib.load().yq_network_set_input_quant(self._h, float(s_in), int(zp_in)))

    def set_debug(self, keep_acc: bool = True) -> None:
        check(_lib.load().yq_network_set_debug(self._h, int(keep_acc)))

    def set_fusion(self, enable: bool) -> None:
        check(_lib.load().yq_network_set_fusion(self._h, int(enable)))

    def set_conv_kernel(self, kind: int) -> None:
        check(_lib.load().yq_network_set_conv_kernel(self._h, int(kind)))

    def use_graph(self, enable: bool = True) -> None:
        check(_lib.load().yq_network_use_graph(self._h, int(enable)))

    def forward_device(self, dev_ptr: int) -> None:
        """forward_network on a device-resident uint8 NCHW batch (asynchronous on self.stream)."""
        check(_lib.load().yq_forward_network_device(self._h, dev_ptr), "yq_forward_network_device")

    def profile_forward(self, dev_ptr: int) -> np.ndarray:
        """Per-layer milliseconds of one forward (CUDA events on the network's stream); [0] = input transform."""
        ms = np.zeros(self.n + 1, np.float32)
        check(_lib.load().yq_network_profile_forward(self._h, dev_ptr, ms.ctypes.data), "yq_network_profile_forward")
        return ms

    def synchronize(self) -> None:
        check(_lib.load().yq_network_synchronize(self._h))

    def predict_u8(self, x: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """network_predict with host buffers: x uint8 [batch,c,h,w]; returns the yolo heads, flat float32."""
        x = np.ascontiguousarray(x, np.uint8)
        assert x.shape == (self.batch, self.c, self.h, self.w), x.shape
        if out is None:
            out = np.empty(self.output_floats, np.float32)
        check(_lib.load().yq_network_predict_u8(self._h, x.ctypes.data, out.ctypes.data), "yq_network_predict_u8")
        return out

    def predict_f32(self, x: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """network_predict with host FLOAT images [batch,c,h,w] (the reference's net->input): the layer-0 input
        quantiser (src/blas.c:108-168 via :279) runs on the device, layer 0 is re-prepared when (s_in, zp_in) changed."""
        x = np.ascontiguousarray(x, np.float32)
        assert x.shape == (self.batch, self.c, self.h, self.w), x.shape
        if out is None:
            out = np.empty(self.output_floats, np.float32)
        check(_lib.load().yq_network_predict_f32(self._h, x.ctypes.data, out.ctypes.data), "yq_network_predict_f32")
        return out

    def predict_image(self, images: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """test_detector's input path: host float images [batch,c,ih,iw] (load_image_color's layout, any size) ->
        letterbox_image (src/image.c:812-831) + dynamic input quantiser (src/blas.c:108-168) on the device -> forward."""
        x = np.ascontiguousarray(images, np.float32)
        assert x.ndim == 4 and x.shape[:2] == (self.batch, self.c), x.shape
        if out is None:
            out = np.empty(self.output_floats, np.float32)
        check(_lib.load().yq_network_predict_image_f32(self._h, x.ctypes.data, int(x.shape[2]), int(x.shape[3]), out.ctypes.data),
              "yq_network_predict_image_f32")
        return out

    def predict_raw(self, in_ptr: int, out_ptr: int) -> None:
        """Same as predict_u8 on raw host addresses (e.g. pinned torch tensors)."""
        check(_lib.load().yq_network_predict_u8(self._h, in_ptr, out_ptr), "yq_network_predict_u8")

    def submit_raw(self, in_ptr: int) -> int:
        """Pipelined network_predict: enqueue H2D + forward of one batch; returns the slot for collect_raw."""
        slot = _lib.load().yq_network_submit_u8(self._h, in_ptr)
        if slot < 0:
            raise YqError(_lib.last_error())
        return slot

    def collect_raw(self, slot: int, out_ptr: int) -> None:
        check(_lib.load().yq_network_collect(self._h, slot, out_ptr), "yq_network_collect")

    def split_heads(self, flat: np.ndarray) -> List[np.ndarray]:
        outs, off = [], 0
        for li in self.layers():
            if li.type == 4:
                cnt = self.batch * li.out_c * li.out_h * li.out_w
                outs.append(flat[off:off + cnt].reshape(self.batch, li.out_c, li.out_h, li.out_w))
                off += cnt
        return outs

    def get_boxes(self, w: int, h: int, thresh: float = 0.5, nms: float = 0.45, relative: int = 1) -> List[np.ndarray]:
        """get_network_boxes + do_nms_sort for every image of the last forward, on the device.
        Returns one [count, 5 + classes] array per image: x, y, w, h, objectness, prob[classes]."""
        lib = _lib.load()
        cap, classes = lib.yq_network_box_capacity(self._h), lib.yq_network_classes(self._h)
        counts = np.zeros(self.batch, np.int32)
        dets = np.empty((self.batch, cap, 5 + classes), np.float32)
        check(lib.yq_network_get_boxes(self._h, int(w), int(h), float(thresh), float(nms), int(relative), counts.ctypes.data,
                                       dets.ctypes.data), "yq_network_get_boxes")
        return [dets[b, :counts[b]].copy() for b in range(self.batch)]

    def pull_layer(self, i: int, what: str = "u8") -> np.ndarray:
        li = self.layer_info(i)
        code, dt = {"u8": (0, np.uint8), "acc": (1, np.int32), "f32": (2, np.float32)}[what]
        out = np.empty((self.batch, li.out_c, li.out_h, li.out_w), dt)
        check(_lib.load().yq_network_pull_layer(self._h, i, code, out.ctypes.data, out.nbytes), "yq_network_pull_layer")
        return out

    def conv_params(self, i: int) -> Dict[str, np.ndarray]:
        n = self.layer_info(i).n
        r = {"M0": np.empty(n, np.int32), "M0_right_shift": np.empty(n, np.int32), "M_value": np.empty(n, np.float64),
             "M0_right_shift_value": np.empty(n, np.float64), "biases_int32": np.empty(n, np.int32)}
        check(_lib.load().yq_network_conv_params(self._h, i, *(a.ctypes.data for a in r.values())))
        return r

    def free(self) -> None:
        if self._h:
            _lib.load().yq_free_network(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def pack_arena_stats() -> Dict[str, int]:
    """entries / hits / misses / dirty of the packed-weight arena (yq_pack.cu)"""
    v = [C.c_int(0) for _ in range(4)]
    check(_lib.load().yq_pack_arena_stats(*[C.byref(x) for x in v]))
    return dict(zip(("entries", "hits", "misses", "dirty"), (int(x.value) for x in v)))


def load_network(cfg: str, weights: str, batch: int = 0, device: int = 0, packed: Optional[str] = None) -> Network:
    """load_network(cfg, weights, clear) (src/network.c:49-57) + the one-time
    quantization_weights_and_activations host prep (src/blas.c:259-346).  ``batch`` overrides the
    cfg's ``[net] batch`` (the reference sizes its buffers from the cfg at parse time).
    ``packed``: path of a packed-weight arena: read first if it exists, so that filter images built for the
    same weights are uploaded as they are; (re)written afterwards when the load had to build any."""
    lib = _lib.load()
    if packed is not None:
        check(lib.yq_pack_arena_clear())
        check(lib.yq_pack_arena_enable(1))
        if os.path.exists(packed) and lib.yq_pack_arena_load(packed.encode()) < 0:
            raise YqError(_lib.last_error())
    h = lib.yq_load_network(cfg.encode(), weights.encode(), int(batch), int(device))
    if not h:
        raise YqError(_lib.last_error())
    if packed is not None:
        if pack_arena_stats()["dirty"] and lib.yq_pack_arena_save(packed.encode()) < 0:
            raise YqError(_lib.last_error())
        check(lib.yq_pack_arena_enable(0))       # later loads / re-preps keep no second copy of their filter images
    return Network(h)


def network_predict(net: Network, x_u8: np.ndarray) -> np.ndarray:
    return net.predict_u8(x_u8)


def free_network(net: Network) -> None:
    net.free()
