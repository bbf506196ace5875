"""ctypes binding of libyq_b200.so (the C ABI declared in include/yq_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C yolo_quantization_b200/csrc``.
There is no fallback of any kind: if the shared object is missing this module raises, and every
compute entry point fails loudly when no CUDA device is visible.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libyq_b200.so")


class YqError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("h", C.c_int), ("w", C.c_int), ("c", C.c_int),
        ("n", C.c_int), ("size", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
        ("activation", C.c_int), ("quant_stop_flag", C.c_int),
        ("zp_in", C.c_int), ("zp_out", C.c_int), ("s_out", C.c_float),
        ("weights_uint8", C.c_void_p), ("weight_zero_point", C.c_void_p), ("biases_int32", C.c_void_p),
        ("M_value", C.c_void_p), ("M0_right_shift_value", C.c_void_p),
        ("saturate", C.c_int),
    ]


class ActGeom(C.Structure):
    _fields_ = [("pad", C.c_int), ("pitch_w", C.c_int), ("rows_h", C.c_int)]


class LayerInfo(C.Structure):
    _fields_ = [
        ("type", C.c_int), ("c", C.c_int), ("h", C.c_int), ("w", C.c_int),
        ("out_c", C.c_int), ("out_h", C.c_int), ("out_w", C.c_int),
        ("n", C.c_int), ("size", C.c_int), ("stride", C.c_int), ("pad", C.c_int), ("activation", C.c_int),
        ("batch_normalize", C.c_int), ("quant_stop_flag", C.c_int),
        ("s_in", C.c_float), ("s_out", C.c_float), ("zp_in", C.c_int), ("zp_out", C.c_int),
        ("kernel", C.c_int), ("classes", C.c_int), ("n_anchors", C.c_int), ("fused", C.c_int),
    ]


_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/yq_b200.h declares (tests check this)
SIGNATURES = {
    "yq_last_error": (C.c_char_p, []),
    "yq_set_abort_on_error": (None, [_i]),
    "yq_device_count": (_i, []),
    "yq_set_device": (_i, [_i]),
    "yq_version": (C.c_char_p, []),
    "yq_cuda_malloc": (_vp, [_sz]),
    "yq_cuda_free": (_i, [_vp]),
    "yq_cuda_push": (_i, [_vp, _vp, _sz, _vp]),
    "yq_cuda_pull": (_i, [_vp, _vp, _sz, _vp]),
    "yq_cuda_memset": (_i, [_vp, _i, _sz, _vp]),
    "yq_stream_synchronize": (_i, [_vp]),
    "yq_host_alloc": (_vp, [_sz, _i]),
    "yq_host_free": (_i, [_vp]),
    "yq_channel_stride": (_i, [_i]),
    "yq_make_convolutional_layer_quant": (_vp, [C.POINTER(ConvDesc)]),
    "yq_free_convolutional_layer_quant": (None, [_vp]),
    "yq_conv_out_h": (_i, [_vp]),
    "yq_conv_out_w": (_i, [_vp]),
    "yq_conv_set_kernel": (_i, [_vp, _i]),
    "yq_conv_get_kernel": (_i, [_vp]),
    "yq_forward_convolutional_layer_quant_gpu": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "yq_forward_convolutional_layer_quant_pool_gpu": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "yq_conv_can_fuse_maxpool": (_i, [_vp]),
    "yq_conv_geom_supported": (_i, [_vp]),
    "yq_conv_out_geom_supported": (_i, [_vp]),
    "yq_forward_convolutional_layer_quant_geom_gpu": (_i, [_vp, _vp, C.POINTER(ActGeom), _i, _vp, C.POINTER(ActGeom), _vp, _vp, _i, _vp]),
    "yq_shortcut_multiplier": (_i, [C.c_float, C.c_float, C.POINTER(C.c_int32)]),
    "yq_forward_shortcut_layer_quant_gpu": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "yq_forward_shortcut_layer_quant_geom_gpu": (_i, [_vp, C.POINTER(ActGeom), _vp, C.POINTER(ActGeom), _vp, C.POINTER(ActGeom), _i, _i, _i, _i, _i, _i, _i, _i,
                                                      _i, _vp]),
    "yq_conv_flat_supported": (_i, [_vp]),
    "yq_act_geom_flat": (_i, [_i, _i, C.POINTER(ActGeom)]),
    "yq_forward_convolutional_layer_quant_flat_gpu": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _vp]),
    "yq_conv_flat_shortcut_supported": (_i, [_vp]),
    "yq_forward_convolutional_layer_quant_flat_shortcut_gpu": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "yq_conv_patch_supported": (_i, [_vp]),
    "yq_conv_flat_up2_supported": (_i, [_vp]),
    "yq_forward_convolutional_layer_quant_flat_up2_gpu": (_i, [_vp, _vp, _vp, _i, _vp]),
    "yq_conv_flat_cat_supported": (_i, [_vp, _i]),
    "yq_forward_convolutional_layer_quant_flat_cat_gpu": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _vp]),
    "yq_forward_convolutional_layer_quant_flat_yolo_gpu": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _i, _vp]),
    "yq_forward_maxpool_layer_quant_geom_gpu": (_i, [_vp, C.POINTER(ActGeom), _vp, C.POINTER(ActGeom), _i, _i, _i, _i, _i, _i, _i, _vp]),
    "yq_forward_upsample_layer_quant_geom_gpu": (_i, [_vp, C.POINTER(ActGeom), _vp, C.POINTER(ActGeom), _i, _i, _i, _i, _i, _vp]),
    "yq_forward_route_layer_quant_geom_gpu": (_i, [C.POINTER(_vp), C.POINTER(ActGeom), C.POINTER(_i), _i, _vp, C.POINTER(ActGeom), _i, _i, _i, _vp]),
    "yq_forward_route_layer_quant_part_gpu": (_i, [C.POINTER(_vp), C.POINTER(ActGeom), C.POINTER(_i), C.POINTER(_i), _i, C.c_uint, _vp, C.POINTER(ActGeom), _i, _i, _i, _vp]),
    "yq_forward_route_layer_quant_up_gpu": (_i, [C.POINTER(_vp), C.POINTER(ActGeom), C.POINTER(_i), C.POINTER(_i), _i, _vp, C.POINTER(ActGeom), _i, _i, _i, _vp]),
    "yq_conv_rows_supported": (_i, [_vp]),
    "yq_conv_rows_input_geom": (_i, [_vp, C.POINTER(ActGeom)]),
    "yq_act_geom_bytes": (_sz, [C.POINTER(ActGeom), _i, _i]),
    "yq_forward_convolutional_layer_quant_rows_pool_gpu": (_i, [_vp, _vp, _vp, C.POINTER(ActGeom), _i, _vp]),
    "yq_conv_rows_nchw_supported": (_i, [_vp]),
    "yq_forward_convolutional_layer_quant_rows_pool_nchw_gpu": (_i, [_vp, _vp, _vp, C.POINTER(ActGeom), _i, _vp]),
    "yq_nchw_to_nhwc_u8_geom": (_i, [_vp, _vp, _i, _i, _i, _i, C.POINTER(ActGeom), _vp]),
    "yq_nhwc_to_nchw_u8_geom": (_i, [_vp, _vp, _i, _i, _i, _i, C.POINTER(ActGeom), _vp]),
    "yq_forward_maxpool_layer_quant_gpu": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "yq_forward_upsample_layer_quant_gpu": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "yq_forward_route_layer_quant_gpu": (_i, [C.POINTER(_vp), C.POINTER(_i), _i, _vp, _i, _i, _i, _vp]),
    "yq_forward_yolo_layer_gpu": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "yq_quantize_input_gpu": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "yq_network_predict_f32": (_i, [_vp, _vp, _vp]),
    "yq_network_predict_image_f32": (_i, [_vp, _vp, _i, _i, _vp]),
    "yq_letterbox_image_gpu": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _vp]),
    "yq_forward_convolutional_layer_quant_per_image_gpu": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp]),
    "yq_nchw_to_nhwc_u8": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "yq_nhwc_to_nchw_u8": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "yq_nhwc_to_nchw_i32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "yq_load_network": (_vp, [C.c_char_p, C.c_char_p, _i, _i]),
    "yq_free_network": (None, [_vp]),
    "yq_network_num_layers": (_i, [_vp]),
    "yq_network_batch": (_i, [_vp]),
    "yq_network_input_dims": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "yq_network_layer_info": (_i, [_vp, _i, C.POINTER(LayerInfo)]),
    "yq_network_set_input_quant": (_i, [_vp, C.c_float, _i]),
    "yq_network_set_debug": (_i, [_vp, _i]),
    "yq_network_set_conv_kernel": (_i, [_vp, _i]),
    "yq_network_set_fusion": (_i, [_vp, _i]),
    "yq_forward_network_device": (_i, [_vp, _vp]),
    "yq_network_predict_u8": (_i, [_vp, _vp, _vp]),
    "yq_network_submit_u8": (_i, [_vp, _vp]),
    "yq_network_collect": (_i, [_vp, _i, _vp]),
    "yq_network_output_floats": (_sz, [_vp]),
    "yq_network_synchronize": (_i, [_vp]),
    "yq_network_stream": (_vp, [_vp]),
    "yq_network_use_graph": (_i, [_vp, _i]),
    "yq_network_launches_per_forward": (_i, [_vp]),
    "yq_network_layer_launches": (_i, [_vp, _i]),
    "yq_dp_load_network": (_vp, [C.c_char_p, C.c_char_p, _i, C.POINTER(_i), _i]),
    "yq_dp_free_network": (None, [_vp]),
    "yq_dp_num_devices": (_i, [_vp]),
    "yq_dp_replica": (_vp, [_vp, _i]),
    "yq_dp_arena_bytes": (_sz, [_vp]),
    "yq_dp_images_from_arena": (_i, [_vp]),
    "yq_dp_network_predict_u8": (_i, [_vp, _vp, _vp]),
    "yq_dp_network_submit_u8": (_i, [_vp, _vp]),
    "yq_dp_network_collect": (_i, [_vp, _i, _vp]),
    "yq_pack_arena_load": (_i, [C.c_char_p]),
    "yq_pack_arena_save": (_i, [C.c_char_p]),
    "yq_pack_arena_clear": (_i, []),
    "yq_pack_arena_enable": (_i, [_i]),
    "yq_pack_arena_stats": (_i, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "yq_network_profile_forward": (_i, [_vp, _vp, _vp]),
    "yq_network_box_capacity": (_i, [_vp]),
    "yq_network_classes": (_i, [_vp]),
    "yq_network_get_boxes": (_i, [_vp, _i, _i, C.c_float, C.c_float, _i, _vp, _vp]),
    "yq_network_pull_layer": (_i, [_vp, _i, _i, _vp, _sz]),
    "yq_network_layer_output_f32_device": (_vp, [_vp, _i]),
    "yq_network_conv_params": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load libyq_b200.so; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise YqError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().yq_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise YqError(f"{what}: {last_error()}" if what else last_error())


def require_gpu() -> None:
    if load().yq_device_count() <= 0:
        raise YqError("no CUDA device visible: the yq_b200 hot path is CUDA-only (sm_100a); there is no CPU fallback")
