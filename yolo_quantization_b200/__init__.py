"""yolo_quantization_b200 -- B200-native (sm_100a) INT8 inference hot path of ArtyZe/yolo_quantization.

Only what the hot path needs lives here:
  csrc/        hand-written CUDA kernels + the C ABI (include/yq_b200.h) -> libyq_b200.so
  _lib.py      ctypes binding (fails loudly when the library or a GPU is missing; no fallback)
  darknet.py   host-side mirror of the reference's API for this path (load_network, network_predict, ...)
  synth.py     seeded synthetic .cfg / .weights / images in the reference's on-disk formats
"""
from . import synth  # noqa: F401

__all__ = ["synth"]
